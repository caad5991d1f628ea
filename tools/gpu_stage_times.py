#!/usr/bin/env python3
"""Development aid: wall time of each pipeline stage with a device sync after each (PAL/NTSC 1 s)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lddecode_b200 import _lib, field as F, pipeline, rfdecode

system = os.environ.get("SYSTEM", "PAL")
audio = os.environ.get("AUDIO", "0") == "1"
ncap = bench.one_second(system) + bench.TAIL
cap = bench.synth_capture(system, ncap, 1)
rf = rfdecode.RFDecode(bench.FS[system], system, 16384, decode_analog_audio=audio)
cd = pipeline.CaptureDecoder(rf, max_fields=256)
be = rf._be
cap_dev = torch.from_numpy(cap).cuda()
sync = torch.cuda.synchronize

def timed(label, fn, acc):
    sync(); t = time.perf_counter(); r = fn(); sync(); acc.setdefault(label, []).append((time.perf_counter() - t) * 1e3); return r

for it in range(4):
    acc = {}
    t_all = time.perf_counter()
    S, N = cd.stride, rf.blocklen
    first_block, nblocks, walk_start = cd.plan_range(ncap, 0, ncap + 1)
    while first_block + (nblocks - 1) * S + N > ncap: nblocks -= 1
    total = nblocks * S
    rf._set_mtf(1)
    planes, parr = timed("alloc planes", lambda: rf._alloc_planes(total), acc)
    a1 = None; alen = 0
    if audio:
        ds = N // len(rf.Filters['audio_lfilt']); alen = total // ds
        a1 = (be.empty(alen, np.float64), be.empty(alen, np.float64))
    timed("demod", lambda: rf._check(be.lib.ldd_demod_blocks(rf._h, be.ptr(cap_dev), 0, 0, ncap, 0, nblocks, total, parr,
          be.ptr(a1[0]) if a1 else None, be.ptr(a1[1]) if a1 else None, alen, be.stream())), acc)
    if audio:
        timed("audio phase2", lambda: rf._audio_phase2_device(a1[0], a1[1], alen), acc)
    gpk, gvl = timed("peaks (+D2H)", lambda: F.sync_peaks_device(rf, planes['demod_sync'], total, 0), acc)
    batch, infos, rs = timed("host walk", lambda: cd._walk(rf, planes, total, 0, ncap, 0, ncap + 1, False, gpk, gvl), acc)
    located = [i for i, f in enumerate(infos) if f.stage == _lib.FIELD_LOCATED]
    def mk():
        sub = F.FieldBatch(rf, len(located))
        for k, i in enumerate(located):
            sub.base[k], sub.winlen[k], sub.linecount[k] = batch.base[i], batch.winlen[i], infos[i].linecount
            sub.linelocs1[k], sub.linebad[k] = batch.linelocs1[i], batch.linebad[i]
        return sub
    sub = timed("build batch (host)", mk, acc)
    ref = timed("refine + tbc", lambda: F.refine_and_tbc(rf, planes, total, sub, want_intermediates=False), acc)
    tot = (time.perf_counter() - t_all) * 1e3
    if it >= 1:
        print("iter", it, "total %.2f ms |" % tot, " | ".join("%s %.2f" % (k, v[0]) for k, v in acc.items()), "| fields", len(located), flush=True)
# a second view: nothing synced except what the pipeline itself syncs
for it in range(3):
    sync(); t = time.perf_counter(); res = cd.decode(cap_dev, 0, ncap); sync(); print("decode() %.2f ms" % ((time.perf_counter() - t) * 1e3))

#!/usr/bin/env python3
"""Development aid: where the host walk's time goes (PAL 1 s)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lddecode_b200 import _lib, field as F, pipeline, rfdecode
system = os.environ.get("SYSTEM", "PAL")
ncap = bench.one_second(system) + bench.TAIL
cap = bench.synth_capture(system, ncap, 1)
rf = rfdecode.RFDecode(bench.FS[system], system, 16384, decode_analog_audio=False)
cd = pipeline.CaptureDecoder(rf, max_fields=256)
be = rf._be
cap_dev = torch.from_numpy(cap).cuda()
res = cd.decode(cap_dev, 0, ncap)
torch.cuda.synchronize()
planes, total = res.planes, res.plane_len
orig = F.sync_peaks_device
tcb = []
def timed_spd(*a, **k):
    t = time.perf_counter(); r = orig(*a, **k); tcb.append((time.perf_counter() - t) * 1e3); return r
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); gpk, gvl = orig(rf, planes['demod_sync'], total, 0); t1 = time.perf_counter()
    F.sync_peaks_device = timed_spd; tcb.clear()
    batch, infos, rs = cd._walk(rf, planes, total, 0, ncap, 0, ncap + 1, False, gpk, gvl)
    F.sync_peaks_device = orig
    t2 = time.perf_counter()
    print("peaks+D2H %.2f ms | walk %.2f ms (of which device chase in callback %.2f ms, calls %d) | fields %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, sum(tcb), len(tcb), len(infos)))

#!/usr/bin/env python3
"""Where the time of an end-to-end step goes (host buffers in, uint16 fields out): host time of every call of the
HostStreamDecoder loop, and the device-side duration of the H2D / D2H copies (CUDA events on their streams)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lddecode_b200 import _lib, pipeline, rfdecode

system = sys.argv[1] if len(sys.argv) > 1 else "PAL"
ncap = bench.one_second(system) + bench.TAIL
cap = bench.synth_capture(system, ncap, 1)
rf = rfdecode.RFDecode(bench.FS[system], system, bench.BLOCKLEN, decode_analog_audio=False, device=0)
cd = pipeline.CaptureDecoder(rf, max_fields=256)
cap_pin = torch.from_numpy(cap).pin_memory()
sd = pipeline.HostStreamDecoder(cd, _lib.FMT_U8, ncap, 64)
acc = {}


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[label] = acc.get(label, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, w)


def run(nsteps):
    pend = sd.launch(sd.upload(cap_pin, ncap))
    t = sd.upload(cap_pin, ncap) if nsteps > 1 else None
    prev = None
    for i in range(nsteps):
        nxt = sd.launch(t) if t is not None else None
        t = sd.upload(cap_pin, ncap) if i + 2 < nsteps else None
        job = sd.finish(pend)
        pend = nxt
        if prev is not None:
            sd.fetch(prev)
        prev = job
    sd.fetch(prev)


run(4)
torch.cuda.synchronize()
# copy durations alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d_in = torch.empty(ncap, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D %d MB" % (ncap >> 20), lambda: d_in.copy_(cap_pin, non_blocking=True)),):
    fn(); torch.cuda.synchronize()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print("%s alone: %.3f ms (%.1f GB/s)" % (name, e0.elapsed_time(e1), ncap / e0.elapsed_time(e1) / 1e6))
h_out = torch.empty(36 << 20, dtype=torch.uint8).pin_memory()
d_out = torch.empty(36 << 20, dtype=torch.uint8, device="cuda")
h_out.copy_(d_out, non_blocking=True); torch.cuda.synchronize()
e0.record(); h_out.copy_(d_out, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H 36 MB alone: %.3f ms (%.1f GB/s)" % (e0.elapsed_time(e1), (36 << 20) / e0.elapsed_time(e1) / 1e6))

for obj, names in ((sd, ("upload", "launch", "finish", "fetch")), (cd, ("_launch", "_finish"))):
    for n in names:
        wrap(obj, n)
# inside _finish: the library call (peak wait + walk + launches) against the Python around it
_orig = rf._be.lib.ldd_pipe_finish


def _timed_finish(*a):
    t = time.perf_counter()
    r = _orig(*a)
    acc["ldd_pipe_finish"] = acc.get("ldd_pipe_finish", 0.0) + time.perf_counter() - t
    return r


rf._be.lib.ldd_pipe_finish = _timed_finish
N = 30
torch.cuda.synchronize()
t0 = time.perf_counter()
run(N)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
print("e2e wall %.3f ms/step" % wall)
for k, v in acc.items():
    print("  %-16s %.3f ms/step" % (k, v / N * 1e3))
# resident for comparison
cap_dev = torch.from_numpy(cap).cuda()
for res in cd.decode_stream((cap_dev, _lib.FMT_U8, ncap) for _ in range(5)):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
for res in cd.decode_stream((cap_dev, _lib.FMT_U8, ncap) for _ in range(N)):
    pass
torch.cuda.synchronize()
print("resident wall %.3f ms/step" % ((time.perf_counter() - t0) / N * 1e3))
# e2e without the download, without the upload

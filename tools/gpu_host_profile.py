"""Host-side profile of the streaming decode (where the Python / ctypes time of a step goes)."""
import cProfile
import os
import pstats
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
cap_dev = torch.from_numpy(cap).cuda()


def run(n):
    for res in cd.decode_stream((cap_dev, _lib.FMT_U8, ncap) for _ in range(n)):
        pass
    torch.cuda.synchronize()


run(5)
t0 = time.perf_counter()
run(20)
print("ms/step", (time.perf_counter() - t0) / 20 * 1e3)
# stage stamps
orig_launch, orig_finish, orig_walk, orig_ref = cd._launch_demod, cd._finish_range, cd._walk, pipeline.F.refine_and_tbc
acc = {}


def wrap(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    return w


cd._launch_demod = wrap("launch_demod", orig_launch)
cd._finish_range = wrap("finish_range(total)", orig_finish)
cd._walk = wrap("walk", orig_walk)
pipeline.F.refine_and_tbc = wrap("refine_and_tbc", orig_ref)
pp = pipeline.F.PendingPeaks.result
pipeline.F.PendingPeaks.result = wrap("peaks_wait", pp)
run(20)
for k, v in acc.items():
    print("%-22s %.3f ms/step" % (k, v / 20 * 1e3))
cd._launch_demod, cd._finish_range, cd._walk, pipeline.F.refine_and_tbc = orig_launch, orig_finish, orig_walk, orig_ref
pipeline.F.PendingPeaks.result = pp
pr = cProfile.Profile()
pr.enable()
run(20)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

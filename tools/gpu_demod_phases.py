#!/usr/bin/env python3
"""Where a block's time goes inside the fused demodulation kernel: runs the kernel from a profiling build of the library
(-DLDD_PHASE_TIMING: thread 0 of every CTA accumulates the clock cycles between the barriers that end each step) and
prints microseconds per block and step.  Build the profiling library HERE first (it travels with the snapshot):
    python tools/gpu_demod_phases.py --build
then on the GPU box:  python tools/gpu_demod_phases.py [mixed|f32] [PAL|NTSC] [audio]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_phase", "libldd_b200_phase.so")

if "--build" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "lddecode_b200", "csrc"))
    import build as B
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    for s in B.SOURCES:
        o = os.path.join(os.path.dirname(OUT), s.replace(".cu", ".o"))
        subprocess.check_call([B.NVCC] + [f for f in B.FLAGS if f not in ("-Xptxas", "-v")] + ["-DLDD_PHASE_TIMING", "-c", os.path.join(B.HERE, s), "-o", o])
        objs.append(o)
    subprocess.check_call([B.NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    print(OUT)
    sys.exit(0)

import numpy as np
import torch

from lddecode_b200 import _lib
_lib.DEFAULT_PATH = OUT
import bench
from lddecode_b200 import pipeline, rfdecode

lane = next((a for a in sys.argv[1:] if a in ("f32", "mixed")), "mixed")
system = "NTSC" if "NTSC" in sys.argv else "PAL"
audio = "audio" in sys.argv
ncap = bench.one_second(system) + bench.TAIL
cap_dev = torch.from_numpy(bench.synth_capture(system, ncap, 1)).cuda()
rf = rfdecode.RFDecode(bench.FS[system], system, bench.BLOCKLEN, decode_analog_audio=audio, device=0, precision=lane)
cd = pipeline.CaptureDecoder(rf)
lib = rf._be.lib
lib.ldd_debug_phases.restype = C.c_int
lib.ldd_debug_phases.argtypes = [C.c_void_p]
buf = (C.c_ulonglong * 64)()
for _ in range(3):
    total = bench.demod_only(cd, cap_dev, _lib.FMT_U8, ncap)
lib.ldd_debug_phases(buf)
R = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(R):
    bench.demod_only(cd, cap_dev, _lib.FMT_U8, ncap)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / R
lib.ldd_debug_phases(buf)
nblocks = total // cd.stride
names = ["A load samples", "B FFT x", "C untangle X", "D audio phase 1", "E RF filter + split", "F 2 FFTs (analytic signal)", "G atan2",
         "G difference", "H FFT demod", "H untangle D", "I tangle x filter (all filters)", "I FFTs (all filters)", "I plane stores (all filters)",
         "J decisions", "J scan rest", "J end", "J2 neighbour decision", "J2 chunk sums", "J2 warp scan", "J2 carry + total", "J2 outputs to staging",
         "J2 plane store"]
mhz = 1965.0
print("%s %s%s: %.3f ms per launch (instrumented), %d blocks" % (system, lane, " +audio" if audio else "", ms, nblocks))
for off, lab in ((0, "float32 pass"), (32, "float64 re-run")):
    v = np.array([buf[off + i] for i in range(len(names))], dtype=np.float64) / R
    if v.sum() == 0:
        continue
    print(" %s: %.1f us per block of this lane's total %.1f SM-ms" % (lab, v.sum() / mhz / nblocks, v.sum() / mhz / 1e3))
    for n, c in zip(names, v):
        if c:
            print("   %-34s %7.2f us/block  %5.1f%%" % (n, c / mhz / nblocks, 100 * c / v.sum()))

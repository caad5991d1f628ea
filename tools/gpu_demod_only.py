#!/usr/bin/env python3
"""Times the block demodulation kernel alone (PAL or NTSC 1 s), for every lane given on the command line.
The command the ncu source-level captures wrap:  python tools/gpu_demod_only.py f32 [PAL|NTSC]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lddecode_b200 import _lib, pipeline, rfdecode

lanes = [a for a in sys.argv[1:] if a in ("f32", "f64", "mixed")] or ["mixed"]
system = "NTSC" if "NTSC" in sys.argv else "PAL"
audio = "audio" in sys.argv
ncap = bench.one_second(system) + bench.TAIL
cap_dev = torch.from_numpy(bench.synth_capture(system, ncap, 1)).cuda()
for lane in lanes:
    rf = rfdecode.RFDecode(bench.FS[system], system, bench.BLOCKLEN, decode_analog_audio=audio, device=0, precision=lane)
    cd = pipeline.CaptureDecoder(rf)
    for _ in range(3):
        total = bench.demod_only(cd, cap_dev, _lib.FMT_U8, ncap)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record()
        bench.demod_only(cd, cap_dev, _lib.FMT_U8, ncap)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%s %s%s: demod %.3f ms median, %.3f min  (%.0f Msamples/s)" % (system, lane, " +audio" if audio else "", float(np.median(ts)),
                                                                        min(ts), total / np.median(ts) / 1e3), flush=True)

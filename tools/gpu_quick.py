#!/usr/bin/env python3
"""Quick GPU probe of the demod kernel variants (development aid): Msamples/s of ldd_demod_blocks
over a synthetic capture resident in HBM, for precision x CTA size x max radix."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lddecode_b200 import _lib, rfdecode, synth  # noqa: E402


def run(system, fs, N, prec, cap_dev, ncap, audio, threads, radix, ctas=1, reps=5, kernel=2):
    os.environ["LDD_THREADS"] = str(threads)
    os.environ["LDD_RADIX_MAX"] = str(radix)
    os.environ["LDD_CTAS_PER_SM"] = str(ctas)
    rf = rfdecode.RFDecode(fs, system, N, decode_analog_audio=audio, precision=prec)
    length = ncap - 2 * N - 2048
    for i in range(2):
        out = rf.demod_device(cap_dev, _lib.FMT_U8, 0, ncap, 0, length, 0, phase2=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(reps):
        e0.record()
        out = rf.demod_device(cap_dev, _lib.FMT_U8, 0, ncap, 0, length, 0, phase2=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    extra = rf.mixed_stats() if prec == "mixed" else None
    del out
    return dict(flagged=extra, system=system, N=N, prec=prec, audio=audio, kernel=kernel, threads=threads, radix=radix, ctas=ctas, ms=round(ms, 3),
                msps=round(length / ms / 1e3, 1))


def main():
    system = os.environ.get("SYSTEM", "PAL")
    fs = 8 * 315 / 88 if system == "NTSC" else 35.46895
    n = int(os.environ.get("NSAMP", 36000000))
    one = synth.SynthRF(system, fs, seed=0).generate(2000000)
    cap = np.tile(one, n // len(one))
    cap_dev = torch.from_numpy(cap).cuda()
    res = []
    variants = [("f64", 512, 16, 1, 1), ("mixed", 512, 16, 1, 1), ("f32", 512, 16, 1, 1)]
    for N in (16384,):
        for prec, thr, rad, ctas, kern in variants:
            try:
                r = run(system, fs, N, prec, cap_dev, len(cap), False, thr, rad, ctas, kernel=kern)
            except Exception as e:
                r = dict(N=N, prec=prec, threads=thr, radix=rad, error=str(e)[:200])
            print(json.dumps(r), flush=True)
            res.append(r)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()

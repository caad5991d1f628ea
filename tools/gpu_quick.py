#!/usr/bin/env python3
"""Quick GPU probe of the demod kernel (development aid): Msamples/s of ldd_demod_blocks for both
precision lanes over a synthetic capture resident in HBM."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lddecode_b200 import _lib, rfdecode, synth  # noqa: E402


def run(system, fs, N, prec, cap_dev, ncap, audio, reps=5):
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
    return dict(system=system, N=N, prec=prec, audio=audio, ms=ms, msps=length / ms / 1e3, ctas=os.environ.get("LDD_CTAS_PER_SM", "2"))


def main():
    fs = 8 * 315 / 88
    n = int(os.environ.get("NSAMP", 12000000))
    t = time.time()
    one = synth.SynthRF("NTSC", fs, seed=0).generate(2000000)
    cap = np.tile(one, n // len(one))
    print("gen", time.time() - t, len(cap))
    cap_dev = torch.from_numpy(cap).cuda()
    res = []
    for prec in ("f64", "f32"):
        for N in (16384, 32768, 65536):
            for audio in (True, False):
                try:
                    r = run("NTSC", fs, N, prec, cap_dev, len(cap), audio)
                except Exception as e:
                    r = dict(N=N, prec=prec, audio=audio, error=str(e))
                print(json.dumps(r), flush=True)
                res.append(r)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "quick.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()

#!/usr/bin/env python3
"""BASELINE.json configs[4]: block-length / overlap sweep of the demodulation kernel on NTSC 8fsc RF:
Msamples/s kept and algorithmic HBM GB/s per point, for the float64 and the mixed lane.
Writes profiles/r01_sweep.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lddecode_b200 import _lib, rfdecode, synth  # noqa: E402


def main():
    fs = 8 * 315 / 88
    one = synth.SynthRF("NTSC", fs, seed=0).generate(2000000)
    cap = np.tile(one, 14)
    ncap = len(cap)
    cap_dev = torch.from_numpy(cap).cuda()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for prec in ("f64", "mixed"):
        for N in (16384, 32768, 65536, 131072, 262144):
            audio = N < 131072                      # audio_phase2 needs len >= blocklen
            for bc in (512, 1024, 2048, 4096):
                rf = rfdecode.RFDecode(fs, "NTSC", N, decode_analog_audio=audio, precision=prec)
                if bc != 1024:
                    rf.set_blockcut(bc)
                S = N - bc - 32
                length = ncap - 2 * N - 2 * bc
                for _ in range(2):
                    o = rf.demod_device(cap_dev, _lib.FMT_U8, 0, ncap, 0, length, 0, phase2=False)
                torch.cuda.synchronize()
                ts = []
                for _ in range(4):
                    e0.record()
                    o = rf.demod_device(cap_dev, _lib.FMT_U8, 0, ncap, 0, length, 0, phase2=False)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = min(ts)
                bps = N / S + 4 * 3 + 8 + (2.0 if audio else 0.0)
                msps = length / ms / 1e3
                r = dict(precision=prec, blocklen=N, blockcut=bc, audio=audio, ms=round(ms, 3), msamples_per_s=round(msps, 1),
                         bytes_per_sample=round(bps, 3), hbm_gbs=round(msps * bps / 1e3, 1), frac_of_measured_hbm=round(msps * bps / 1e3 / peak, 4))
                print(json.dumps(r), flush=True)
                out.append(r)
                del o, rf
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()

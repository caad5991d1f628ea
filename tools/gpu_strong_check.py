#!/usr/bin/env python3
"""world_size-N NCCL check of the strong-scaling path on real GPUs (torchrun): ONE tiled NTSC CLV capture (seconds given
on the command line, default 2 per rank) sharded by read-position ranges; every rank generates only its window, decodes
it in chunks through decode_stream with the gather's send buffers as the TBC kernel's destination; rank 0 collects every
chunk's fields and compares (a) the union with its own single-range decode of the whole capture (bit-identical), and
(b) the first and last field of every shard -- both sides of every seam -- with the oracle (+-1 LSB)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lddecode_b200 import _lib, parallel, pipeline, rfdecode, synth  # noqa: E402


def _oracle_field(args):
    from oracle import ldd_oracle as O
    cap, rs = args
    dec = O.Decoder(8 * 315 / 88, "NTSC", 16384, analog_audio=False)
    d = O.demod(dec, lambda s, n: cap[s:s + n] if s + n <= len(cap) else None, rs, 1000000, 1)
    f = O.decode_field(dec, d[0], 0)
    return f.valid, int(getattr(f, "istop", -1)), f.dspicture


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    fs = 8 * 315 / 88
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0 * world
    ncap = int(round(seconds * fs * 1e6)) + 1100000
    tc = synth.TiledCapture(seed=2, device="cuda")
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, device=local)
    cd = pipeline.CaptureDecoder(rf, max_fields=4096)
    R0, R1 = parallel.shard_bounds(ncap, world)[rank]
    lo, hi = parallel.needed_window(cd, ncap, R0, R1)
    win = tc.generate(lo, hi - lo)
    nch = 2
    step = ((ncap // world) + nch) // nch
    edges = [R0 + i * step for i in range(nch)] + [R1]
    g = parallel.make_gatherer(cd, rank, world, 96, dist)
    got = []
    for res in cd.decode_stream(iter([(win, _lib.FMT_U8, lo, hi - lo, ncap, edges[i], edges[i + 1]) for i in range(nch)]), sink=g):
        part = g.to_host()
        if rank == 0:
            got += part
    if rank == 0:
        whole = tc.generate(0, ncap)
        one = cd.pictures(cd.decode(whole, _lib.FMT_U8, ncap))
        got.sort(key=lambda t: t[0])
        same = len(got) == len(one) and all(a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) for a, b in zip(got, one))
        # seams: first and last owned field of every shard
        bounds = parallel.shard_bounds(ncap, world)
        rs_all = [p[0] for p in one]
        checks = set()
        for r0, r1 in bounds:
            own = [r for r in rs_all if r0 <= r < r1]
            checks |= {own[0], own[-1]}
        checks = sorted(checks)
        host = whole.cpu().numpy()
        with mp.get_context("fork").Pool(min(8, len(checks))) as pool:
            ref = pool.map(_oracle_field, [(host, rs) for rs in checks])
        by_rs = {p[0]: p for p in got}
        worst = 0
        for rs, (valid, istop, pic) in zip(checks, ref):
            d = np.abs(by_rs[rs][2].astype(np.int64) - pic.astype(np.int64))
            worst = max(worst, int(d.max()))
            same = same and valid and istop == by_rs[rs][1]
        print("STRONG_OK" if same and worst <= 1 else "STRONG_MISMATCH", "fields", len(got), "of", len(one), "seam checks", len(checks),
              "worst LSB", worst, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""world_size-N NCCL check of parallel.FieldGatherer on real GPUs (torchrun): shard a short capture, decode, gather
on rank 0 and compare with the single-GPU decode.  The CPU suite covers the same logic over gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lddecode_b200 import _lib, parallel, pipeline, rfdecode, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    fs = 8 * 315 / 88
    ncap = 2600000
    cap = synth.SynthRF("NTSC", fs, seed=9).generate(ncap)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, device=local)
    cd = pipeline.CaptureDecoder(rf)
    r0, r1 = parallel.shard_bounds(ncap, world)[rank]
    lo, hi = parallel.needed_window(cd, ncap, r0, r1)
    g = parallel.make_gatherer(cd, rank, world, 8, dist)
    for _ in range(3):                                   # buffers are reused across gathers
        res = cd.decode_range(torch.from_numpy(cap[lo:hi]).cuda(), _lib.FMT_U8, lo, hi - lo, ncap, r0, r1)
        g.gather(res)
    got = g.to_host()
    if rank == 0:
        one = cd.pictures(cd.decode(torch.from_numpy(cap).cuda(), _lib.FMT_U8, ncap))
        ok = len(got) == len(one) and all(a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) for a, b in zip(got, one))
        print("GATHER_OK" if ok else "GATHER_MISMATCH", len(got), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Worker of tests/test_pipeline.py::test_two_rank_strong_scaling_gloo (launched by torch.distributed.run): ONE tiled
capture sharded by read-position ranges over two ranks, each rank generating only the window it needs and working
through it in two chunks (decode_stream over ranges, pictures written straight into the gather's send buffers), NCCL's
place taken by gloo.  Rank 0 checks the gathered fields of the last chunk pair against the single-range decode."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from emu_util import emu_backend  # noqa: E402
from lddecode_b200 import _lib, parallel, pipeline, rfdecode, synth  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    be = emu_backend()
    fs = 8 * 315 / 88
    ncap = 3400000
    tc = synth.TiledCapture(seed=2, device="cpu")
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=be, precision="f64")
    cd = pipeline.CaptureDecoder(rf)
    r0, r1 = parallel.shard_bounds(ncap, world)[rank]
    lo, hi = parallel.needed_window(cd, ncap, r0, r1)
    win = tc.generate(lo, hi - lo).numpy()
    mid = (r0 + min(r1, ncap)) // 2
    g = parallel.FieldGatherer(cd, rank, world, 8, dist)
    got = []
    for res in cd.decode_stream(iter([(win, _lib.FMT_U8, lo, hi - lo, ncap, r0, mid), (win, _lib.FMT_U8, lo, hi - lo, ncap, mid, r1)]), sink=g):
        part = g.to_host()
        if rank == 0:
            got += part
    if rank == 0:
        whole = tc.generate(0, ncap).numpy()
        one = cd.pictures(cd.decode(whole, _lib.FMT_U8, ncap))
        got.sort(key=lambda t: t[0])
        ok = len(got) == len(one) and all(a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) for a, b in zip(got, one))
        print("STRONG_OK" if ok else "STRONG_MISMATCH", len(got), len(one), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

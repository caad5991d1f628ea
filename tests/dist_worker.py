"""Worker of tests/test_pipeline.py::test_two_rank_gloo_gather (launched by torch.distributed.run)."""
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
    ncap = 2600000
    cap = synth.SynthRF("NTSC", fs, seed=9).generate(ncap)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=be, precision="f64")
    cd = pipeline.CaptureDecoder(rf)
    r0, r1 = parallel.shard_bounds(ncap, world)[rank]
    lo, hi = parallel.needed_window(cd, ncap, r0, r1)
    res = cd.decode_range(be.to_device(cap[lo:hi]), _lib.FMT_U8, lo, hi - lo, ncap, r0, r1)
    got = parallel.gather_fields(cd, res, rank, world, 8, dist)
    if rank == 0:
        one = cd.pictures(cd.decode(be.to_device(cap), _lib.FMT_U8, ncap))
        ok = len(got) == len(one) and all(a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) for a, b in zip(got, one))
        print("GATHER_OK" if ok else "GATHER_MISMATCH", len(got), flush=True)
    # 48 kHz PCM of the sharded capture: every rank chains the offsets over the fields of the ranks before it and
    # resamples its own fields; rank 0 gets all of them.  Same counts / dropped fields as one process working through the
    # same ranges in sequence, identical samples (same ranges, same audio).
    rfa = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=be)
    cda = pipeline.CaptureDecoder(rfa)
    resa = cda.decode_range(be.to_device(cap[lo:hi]), _lib.FMT_U8, lo, hi - lo, ncap, r0, r1)
    pcm = parallel.sharded_pcm(cda, resa, rank, world, dist)
    if rank == 0:
        want, offset, state = [], 0.0, 2
        for q0, q1 in parallel.shard_bounds(ncap, world):
            a, b = parallel.needed_window(cda, ncap, q0, q1)
            r = cda.decode_range(be.to_device(cap[a:b]), _lib.FMT_U8, a, b - a, ncap, q0, q1)
            p, offset, state = cda.pcm(r, offset, state)
            want += [(int(r.readsamples[j]), x) for j, x in zip(r.located, p)]
        same = len(pcm) == len(want) == len(got) and all(
            a[0] == b[0] and ((a[1] is None and b[1] is None) or (a[1] is not None and b[1] is not None and np.array_equal(a[1], b[1])))
            for a, b in zip(pcm, want))
        print("PCM_OK" if same else "PCM_MISMATCH", sum(x[1] is not None for x in pcm), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

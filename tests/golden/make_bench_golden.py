#!/usr/bin/env python3
"""Golden fields for bench.py's self-check (tests/golden/bench_fields.npz).

bench.py verifies, outside its timed regions, that the first field its GPU path decodes from each benchmark capture
is the field the REFERENCE decodes from the same bytes (+-1 LSB of uint16).  This script produces those fields by
running the unmodified reference (/root/reference, through tools/refshim.py) on the first read window of the same
seeded synthetic captures bench.py generates: RFDecode.demod(0, 1e6, mtf_level=1) -> FieldPAL / FieldNTSC, exactly
what Framer.readfield does first (lddecode_core.py:1194-1203).  Build container only; the vectors are committed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refshim  # noqa: E402
from lddecode_b200 import synth  # noqa: E402

FS = {"NTSC": 8 * 315 / 88, "PAL": 35.46895}
# (key, system, seed, bits, audio): the rank-0 captures of bench.py's workloads
CASES = [("PAL_u8_seed1", "PAL", 1, 8, False), ("NTSC_10bit_seed0", "NTSC", 0, 10, True), ("NTSC_u8_seed0", "NTSC", 0, 8, True)]


def main():
    core = refshim.load_reference()
    out = {}
    for key, system, seed, bits, audio in CASES:
        cap = synth.SynthRF(system, FS[system], seed=seed, bits=bits).generate(1100000)
        rf = core.RFDecode(inputfreq=FS[system], system=system, blocklen_=16384, decode_analog_audio=audio)
        core.loader = refshim.make_array_loader(cap)
        data = rf.demod(refshim.MemFile(b""), 0, 1000000, 1)
        f = (core.FieldPAL if system == "PAL" else core.FieldNTSC)(rf, data, 0)
        assert f.valid, key
        out[key + "_pic"] = np.asarray(f.dspicture, dtype=np.uint16)
        out[key + "_next"] = np.int64(f.nextfieldoffset)
        out[key + "_istop"] = np.int64(f.istop)
        print(key, "field", len(f.dspicture), "next", f.nextfieldoffset, "istop", f.istop, flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bench_fields.npz"), **out)


if __name__ == "__main__":
    main()

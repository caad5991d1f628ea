"""Full-size checks on the GPU (BASELINE.json configs at 1 s of RF): the oracle is too slow to replay
a whole second, so these use size-independent properties of the path plus oracle spot checks."""
import numpy as np
import pytest

from lddecode_b200 import _lib, parallel, pipeline, rfdecode, synth
from oracle import ldd_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ntsc_second():
    fs = 8 * 315 / 88
    n = int(round(fs * 1e6)) + 1100000
    return fs, synth.SynthRF("NTSC", fs, seed=0, bits=10).generate(n)


def test_one_second_ntsc_properties(cuda_backend, ntsc_second):
    """config 1/3 at full size: field cadence, parity alternation, read positions, and bit-identity of
    (a) a re-run, (b) a 3-range sharded decode, (c) decodes straight from .r30 / .lds packed bytes."""
    be = cuda_backend
    fs, s10 = ntsc_second
    ncap = len(s10) // 12 * 12
    s10 = s10[:ncap]
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, precision="f64", _backend=be)
    cd = pipeline.CaptureDecoder(rf, max_fields=256)
    dev = be.to_device(s10)
    res = cd.decode(dev, _lib.FMT_U16, ncap)
    pics = cd.pictures(res)
    assert 58 <= len(pics) <= 61                                   # 59.94 fields per second, minus the tail
    assert all(p[2] is not None for p in pics)
    tops = [p[1] for p in pics]
    assert all(a != b for a, b in zip(tops, tops[1:]))             # top / bottom alternate
    gaps = np.diff([p[0] for p in pics])
    assert np.all(np.abs(gaps[1:] - 477750) < 3000)                # one field = 262.5 lines of 1820 samples (the first read starts mid-field)
    W = rf.SysParams["outlinelen"]
    for rs, istop, pic in pics:
        assert len(pic) == (263 if istop else 262) * W
    # (a) idempotence
    again = cd.pictures(cd.decode(dev, _lib.FMT_U16, ncap))
    assert all(np.array_equal(a[2], b[2]) and a[0] == b[0] for a, b in zip(pics, again))
    # (b) sharding by read-position ranges is bit-identical
    parts = []
    for r0, r1 in parallel.shard_bounds(ncap, 3):
        lo, hi = parallel.needed_window(cd, ncap, r0, r1)
        parts += cd.pictures(cd.decode_range(be.to_device(s10[lo:hi]), _lib.FMT_U16, lo, hi - lo, ncap, r0, r1))
    assert len(parts) == len(pics)
    assert all(a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) for a, b in zip(parts, pics))
    # (c) packed captures, unpack fused into the block load
    for fmt, packed in ((_lib.FMT_R30, synth.pack_r30(s10)), (_lib.FMT_LDS40, synth.pack_lds(s10))):
        got = cd.pictures(cd.decode(be.to_device(packed), fmt, ncap))
        assert len(got) == len(pics)
        assert all(a[0] == b[0] and np.array_equal(a[2], b[2]) for a, b in zip(got, pics))


@pytest.mark.parametrize("precision", ["f64", "mixed"])
def test_one_second_oracle_spot_checks(cuda_backend, ntsc_second, precision):
    """Three fields of the 1-s decode (first, middle, last) against the oracle's own decode of the same
    read window: peak list identical, TBC within +-1 LSB -- for the exact and for the mixed lane."""
    be = cuda_backend
    fs, s10 = ntsc_second
    ncap = len(s10)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, precision=precision, _backend=be)
    dec = O.Decoder(fs, "NTSC", 16384, analog_audio=False)
    cd = pipeline.CaptureDecoder(rf, max_fields=256)
    res = cd.decode(be.to_device(s10), _lib.FMT_U16, ncap)
    pics = cd.pictures(res)
    ld = lambda s, n: s10[s:s + n] if s + n <= ncap else None
    for k in (0, len(pics) // 2, len(pics) - 1):
        rs = pics[k][0]
        ov, _ = O.demod(dec, ld, rs, 1000000, 1)
        of = O.decode_field(dec, ov, 0)
        assert of.valid
        j = res.located[k]
        assert res.infos[j].npeaks == len(of.peaklist) and res.infos[j].nextfieldoffset == of.nextfieldoffset
        d = pics[k][2].astype(np.int64) - of.dspicture.astype(np.int64)
        print("spot", precision, k, "maxdiff", np.abs(d).max(), "differing", np.count_nonzero(d), "of", d.size)
        assert np.abs(d).max() <= 1


def test_one_second_mixed_lane_equals_f64_sync(cuda_backend, ntsc_second):
    """config 5 flavour at full size: the mixed lane's demod_sync plane and peak list are bit-identical
    to the float64 lane's over the whole second, for two block lengths."""
    be = cuda_backend
    fs, s10 = ntsc_second
    ncap = len(s10)
    dev = be.to_device(s10)
    for N in (16384, 32768):
        out = {}
        for prec in ("f64", "mixed"):
            rf = rfdecode.RFDecode(fs, "NTSC", N, decode_analog_audio=False, precision=prec, _backend=be)
            cd = pipeline.CaptureDecoder(rf, max_fields=256)
            res = cd.decode(dev, _lib.FMT_U16, ncap)
            out[prec] = (res.planes["demod_sync"].clone(), np.array(res.gpeaks))
            if prec == "mixed":
                flagged, total = rf.mixed_stats()
                assert flagged < total // 5
        import torch
        assert torch.equal(out["f64"][0], out["mixed"][0])
        assert np.array_equal(out["f64"][1], out["mixed"][1])

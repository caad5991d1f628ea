"""Frame assembly and 48 kHz audio: the Framer mirror on top of the
device-backed classes against a whole frame recorded from the reference's Framer.readframe
(lddecode.py:88-98 loop), and downscale_audio against the reference's per-field PCM."""
import numpy as np
import pytest

from lddecode_b200 import field, framer, rfdecode
from oracle import ldd_oracle as O


def _mem_loader(cap):
    def ld(infile, sample, n):
        if sample + n > len(cap):
            return None
        return cap[sample:sample + n]
    return ld


@pytest.mark.parametrize("precision", ["f64", None])
def test_readframe_matches_reference(backend, golden, precision):
    """precision None = the library default lane ('mixed')."""
    g = golden("ntsc_frame")
    cap = g["capture"]
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]), _backend=backend, precision=precision)
    rfdecode.loader = _mem_loader(cap)
    fr = framer.Framer(rf)
    combined, conaudio, nextsample, fields = fr.readframe(None, 0, True)
    assert nextsample == int(g["nextsample"])
    assert fr.vbi["framenr"] == int(g["framenr"])
    np.testing.assert_allclose(fr.mtf_level, float(g["mtf_level"]), rtol=1e-12)
    assert [f.linecount for f in fields] == list(g["field_readlens"])
    d = combined.astype(np.int64) - g["combined"].astype(np.int64)
    assert np.abs(d).max() <= 1                                                   # +-1 LSB of uint16
    if precision == "f64":
        assert np.count_nonzero(d) < 0.002 * d.size
    assert len(conaudio) == len(g["conaudio"])
    da = conaudio.astype(np.int64) - g["conaudio"].astype(np.int64)
    assert np.abs(da).max() <= 1                                                  # int16 PCM, +-1 LSB
    np.testing.assert_allclose(fr.audio_offset, float(g["audio_offset"]), rtol=0, atol=1e-12)


def test_downscale_audio_golden(backend, golden):
    """Field-level PCM (lddecode_core.py:431-484) on the reference's own line table."""
    g = golden("ntsc")
    cap = g["capture"]
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]), _backend=backend, precision="f64")
    rfdecode.loader = _mem_loader(cap)
    raw = rf.demod_raw(None, 0, int(g["demod_length"]), 1)
    out16, nxt = framer.downscale_audio(raw.audio_recarray(), g["field_linelocs"], rf, int(g["field_linecount"]), 0)
    assert len(out16) == len(g["field_dsaudio"])
    assert np.abs(out16.astype(np.int64) - g["field_dsaudio"].astype(np.int64)).max() <= 1
    np.testing.assert_allclose(nxt, float(g["field_audio_next_offset"]), rtol=0, atol=1e-12)

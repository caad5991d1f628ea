"""Callers either side of the hot path: the REFERENCE's own Framer (lddecode_core.py:1193-1334,
unmodified, imported through tools/refshim.py) runs on top of the drop-in RFDecode / FieldNTSC classes and must
reproduce the frame recorded from the all-reference run (tests/golden/ntsc_frame.npz: picture +-1 LSB, 48 kHz PCM
+-1 LSB, VBI frame number, MTF level, next read position).  The package itself ships no Framer: frame pairing and
seek are control logic the reference keeps.  downscale_audio (the PCM resample along the line positions) is a
device kernel and is checked against the reference's per-field PCM as well."""
import os
import sys

import numpy as np
import pytest

from lddecode_b200 import field, rfdecode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import refshim  # noqa: E402


def _mem_loader(cap):
    def ld(infile, sample, n):
        if sample + n > len(cap):
            return None
        return cap[sample:sample + n]
    return ld


@pytest.mark.skipif(not refshim.available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("precision,device_resident", [("f64", False), (None, True)])
def test_reference_framer_over_dropin_classes(backend, golden, precision, device_resident, monkeypatch):
    """precision None = the library default lane ('mixed').  device_resident: rf.demod hands the Field classes the
    device-resident result (demod_raw) instead of host record arrays -- both are valid `rawdecode` arguments."""
    core = refshim.load_reference()
    g = golden("ntsc_frame")
    cap = g["capture"]
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]), _backend=backend, precision=precision)
    rfdecode.loader = _mem_loader(cap)
    if device_resident:
        rf.demod = rf.demod_raw
    monkeypatch.setattr(core, "FieldNTSC", field.FieldNTSC)
    monkeypatch.setattr(core, "FieldPAL", field.FieldPAL)
    monkeypatch.setattr(core, "Field", field.Field)
    fr = core.Framer(rf)                                        # the reference's class, our rf
    assert fr.FieldClass is field.FieldNTSC
    combined, conaudio, nextsample, fields = fr.readframe(refshim.MemFile(b""), 0, True)      # Framer checks isinstance(io.IOBase)
    assert nextsample == int(g["nextsample"])
    assert fr.vbi["framenr"] == int(g["framenr"])
    np.testing.assert_allclose(fr.mtf_level, float(g["mtf_level"]), rtol=1e-12)
    assert [f.linecount for f in fields] == list(g["field_readlens"])
    d = combined.astype(np.int64) - g["combined"].astype(np.int64)
    assert np.abs(d).max() <= 1                                                   # +-1 LSB of uint16
    if precision == "f64":
        assert np.count_nonzero(d) < 0.002 * d.size
    assert conaudio is not None and len(conaudio) == len(g["conaudio"])
    da = conaudio.astype(np.int64) - g["conaudio"].astype(np.int64)
    assert np.abs(da).max() <= 1                                                  # int16 PCM, +-1 LSB
    np.testing.assert_allclose(fr.audio_offset, float(g["audio_offset"]), rtol=0, atol=1e-12)


def test_downscale_audio_golden(backend, golden):
    """Field-level PCM (lddecode_core.py:431-484) on the reference's own line table, host record array and device
    buffers as input."""
    g = golden("ntsc")
    cap = g["capture"]
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]), _backend=backend, precision="f64")
    rfdecode.loader = _mem_loader(cap)
    raw = rf.demod_raw(None, 0, int(g["demod_length"]), 1)
    for audio in (raw.audio_recarray(), raw.audio):
        out16, nxt = field.downscale_audio(audio, g["field_linelocs"], rf, int(g["field_linecount"]), 0)
        assert out16.dtype == np.int16 and len(out16) == len(g["field_dsaudio"])
        assert np.abs(out16.astype(np.int64) - g["field_dsaudio"].astype(np.int64)).max() <= 1
        np.testing.assert_allclose(nxt, float(g["field_audio_next_offset"]), rtol=0, atol=1e-12)
    # a non-zero carried offset: same arange as numpy's
    out16b, nxtb = field.downscale_audio(raw.audio, g["field_linelocs"], rf, int(g["field_linecount"]), float(nxt))
    fr = (rf.SysParams['line_period'] * int(g["field_linecount"])) / 1000000
    ar = np.arange(float(nxt), fr + 1 / 48000.0, 1 / 48000.0)
    assert len(out16b) == 2 * (len(ar) - 1) and nxtb == ar[-1] - fr

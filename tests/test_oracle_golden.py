"""Pins oracle/ldd_oracle.py to the reference: every function is checked against the vectors that
tests/golden/make_golden.py recorded from the unmodified reference (float64, so tolerances are
rounding-level).  CPU only."""
import numpy as np
import pytest

from oracle import ldd_oracle as O

CASES = ["ntsc", "pal", "ntsc10"]


def _decoder(g):
    system = "PAL" if "demod_pilot" in list(g["planes"]) else "NTSC"
    return O.Decoder(float(g["fs_mhz"]), system, int(g["blocklen"]))


def _loader(cap):
    def ld(sample, n):
        if sample + n > len(cap):
            return None
        return cap[sample:sample + n]
    return ld


@pytest.mark.parametrize("name", CASES)
def test_demodblock_matches_reference(golden, name):
    g = golden(name)
    dec = _decoder(g)
    cap = g["capture"]
    for bi in range(3):
        pos, mtf = int(g["blk%d_pos" % bi]), float(g["blk%d_mtf" % bi])
        v, a = O.demodblock(dec, cap[pos:pos + dec.N], mtf)
        for p in g["planes"]:
            np.testing.assert_allclose(v[p], g["blk%d_%s" % (bi, p)], rtol=1e-13, atol=1e-7, err_msg=p)
        np.testing.assert_allclose(a["audio_left"], g["blk%d_audio_left" % bi], rtol=1e-13, atol=1e-7)
        np.testing.assert_allclose(a["audio_right"], g["blk%d_audio_right" % bi], rtol=1e-13, atol=1e-7)


@pytest.fixture(scope="module")
def stitched(golden):
    cache = {}

    def get(name):
        if name not in cache:
            g = golden(name)
            dec = _decoder(g)
            cache[name] = (dec, O.demod(dec, _loader(g["capture"]), 0, int(g["demod_length"]), 1))
        return cache[name]
    return get


@pytest.mark.parametrize("name", CASES)
def test_demod_stitching_and_audio(golden, stitched, name):
    g = golden(name)
    dec, (video, audio) = stitched(name)
    sp = int(g["sparse"])
    N = dec.N
    assert len(video["demod"]) == int(g["demod_len"])
    for p in g["planes"]:
        np.testing.assert_allclose(video[p][::sp], g["demod_sparse_" + p], rtol=1e-13, atol=1e-7)
        np.testing.assert_allclose(video[p][N - 1056 - 2048:N - 1056 + 2048], g["demod_seam_" + p], rtol=1e-13, atol=1e-7)
        np.testing.assert_allclose(video[p][-4096:], g["demod_tail_" + p], rtol=1e-13, atol=1e-7)
    np.testing.assert_allclose(audio["audio_left"], g["audio_left"], rtol=1e-12, atol=1e-6)
    np.testing.assert_allclose(audio["audio_right"], g["audio_right"], rtol=1e-12, atol=1e-6)
    v2, a2 = O.demod(dec, _loader(g["capture"]), 54321, 300000, 0)
    assert len(v2["demod"]) == int(g["demod2_len"])
    for p in g["planes"]:
        np.testing.assert_allclose(v2[p][::sp], g["demod2_sparse_" + p], rtol=1e-13, atol=1e-7)
    np.testing.assert_allclose(a2["audio_left"], g["demod2_audio_left"], rtol=1e-12, atol=1e-6)
    assert O.demod(dec, _loader(g["capture"]), len(g["capture"]) - 5000, 100000, 0) is None


@pytest.mark.parametrize("name", CASES)
def test_field_matches_reference(golden, stitched, name):
    g = golden(name)
    dec, (video, audio) = stitched(name)
    f = O.decode_field(dec, video, 0)
    assert f.valid == bool(g["field_valid"])
    assert np.array_equal(np.array(f.peaklist), g["field_peaklist"])          # bit-exact indices
    assert np.array_equal(np.array(f.vsyncs), g["field_vsyncs"])
    assert f.nextfieldoffset == int(g["field_nextfieldoffset"])
    assert int(f.istop) == int(g["field_istop"]) and f.linecount == int(g["field_linecount"])
    np.testing.assert_allclose(f.med_hsync, g["field_med_hsync"], rtol=1e-13)
    np.testing.assert_allclose(f.linelocs1, g["field_linelocs1"], rtol=0, atol=1e-9)
    assert np.array_equal(np.array(f.linebad, dtype=np.int8), g["field_linebad"])
    np.testing.assert_allclose(f.linelocs2, g["field_linelocs2"], rtol=0, atol=1e-7)
    if dec.system == "NTSC":
        np.testing.assert_allclose(f.linelocs3, g["field_linelocs3"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(f.linelocs4, g["field_linelocs4"], rtol=0, atol=1e-7)
        np.testing.assert_array_equal(f.burstlevel, g["field_burstlevel"])
    np.testing.assert_allclose(f.linelocs, g["field_linelocs"], rtol=0, atol=1e-7)
    assert np.array_equal(f.dspicture, g["field_dspicture"])                    # identical uint16 TBC output
    codes = [[-1] * 6 if f.linecode[l] is None else f.linecode[l] for l in dec.SP["philips_codelines"]]
    assert np.array_equal(np.array(codes), g["field_linecode"])


def test_downscale_audio_matches_reference(golden, stitched):
    """downscale_audio on the reference's own phase-2 audio and line table: identical int16 PCM and carried offset."""
    g = golden("ntsc")
    dec, (video, audio) = stitched("ntsc")
    ref_audio = {"audio_left": g["audio_left"], "audio_right": g["audio_right"]}
    pcm, nxt = O.downscale_audio(dec, ref_audio, g["field_linelocs"], int(g["field_linecount"]), 0)
    assert pcm.dtype == np.int16 and np.array_equal(pcm, g["field_dsaudio"])
    assert nxt == float(g["field_audio_next_offset"])
    pcm2, _ = O.downscale_audio(dec, audio, g["field_linelocs"], int(g["field_linecount"]), 0)       # the oracle's own audio
    assert np.abs(pcm2.astype(np.int64) - g["field_dsaudio"]).max() <= 1


def test_framer_audio_walk_matches_reference_frame(golden):
    """The Framer.readframe audio chain of the oracle against the frame recorded from the reference's own Framer."""
    g = golden("ntsc_frame")
    dec = O.Decoder(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]))
    fields, offset = O.framer_audio_walk(dec, _loader(g["capture"]), 2)
    assert [f.linecount for _, f, _ in fields] == list(g["field_readlens"])
    con = np.concatenate([p for _, _, p in fields if p is not None])
    assert len(con) == len(g["conaudio"]) and np.abs(con.astype(np.int64) - g["conaudio"]).max() <= 1
    np.testing.assert_allclose(offset, float(g["audio_offset"]), rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["ntsc", "pal"])
def test_scale_and_notaknot_restatement(golden, stitched, name):
    """scale() == reference scale(); the written-out not-a-knot spline equals it to <1e-6 Hz."""
    g = golden(name)
    dec, (video, _) = stitched(name)
    ll = g["field_linelocs"]
    W = dec.SP["outlinelen"]
    for k, l in enumerate(g["scale_lines"]):
        ref = g["scale_out"][k]
        np.testing.assert_allclose(O.scale(video["demod"], ll[l], ll[l + 1], W), ref, rtol=1e-12)
        np.testing.assert_allclose(O.scale_notaknot(video["demod"], ll[l], ll[l + 1], W), ref, rtol=0, atol=1e-6)


def test_unpackers_match_reference(golden):
    g = golden("unpack")
    assert np.array_equal(O.unpack_r30_ddunpack(g["r30_words"]), g["r30_ddunpack_i16"])
    for s, exp in zip(g["r30_py_starts"], g["r30_py"]):
        w0 = int(s) // 3
        assert np.array_equal(O.unpack_r30_raw(g["r30_words"][w0:], int(s) % 3, 1000), exp)
    for s, exp in zip(g["lds_py_starts"], g["lds_py"]):
        b0 = (int(s) // 4) * 5
        assert np.array_equal(O.unpack_lds(g["lds_bytes"][b0:], int(s) % 4, 1000), exp)

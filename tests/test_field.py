"""Kernels (4) and (5) and the field chain through the drop-in Field classes, against the golden
vectors recorded from the reference's FieldNTSC / FieldPAL (lddecode_core.py:489-1191).

Bars (BASELINE.json north_star): sync-peak indices bit-exact; TBC output within +-1 LSB of uint16.
Line positions are held to 1e-5 samples, burst levels to 2 float32 ulps."""
import os

import numpy as np
import pytest

from lddecode_b200 import _lib, field, rfdecode
from oracle import ldd_oracle as O


def _setup(backend, g, name):
    system = "PAL" if name == "pal" else "NTSC"
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), system, int(g["blocklen"]), _backend=backend, precision="f64")
    cap = g["capture"]
    fmt = _lib.FMT_U8 if cap.dtype == np.uint8 else _lib.FMT_U16
    dd = rf.demod_device(backend.to_device(cap), fmt, 0, len(cap), 0, int(g["demod_length"]), 1)
    return rf, dd, system


@pytest.mark.parametrize("name", ["ntsc", "pal", "ntsc10"])
def test_field_golden(backend, golden, name):
    g = golden(name)
    rf, dd, system = _setup(backend, g, name)
    f = (field.FieldNTSC if system == "NTSC" else field.FieldPAL)(rf, dd, 0)
    assert f.valid == bool(g["field_valid"])
    assert np.array_equal(np.array(f.peaklist), g["field_peaklist"])                 # bit-exact peak indices
    assert np.array_equal(np.array(f.vsyncs), g["field_vsyncs"])
    assert f.nextfieldoffset == int(g["field_nextfieldoffset"]) and f.tbcstart == int(g["field_tbcstart"])
    assert int(f.istop) == int(g["field_istop"]) and f.linecount == int(g["field_linecount"])
    np.testing.assert_allclose(f.med_hsync, g["field_med_hsync"], rtol=1e-12)
    np.testing.assert_allclose(f.hsync_tolerance, g["field_hsync_tolerance"], rtol=1e-9)
    np.testing.assert_array_equal(np.array(f.linelocs1), g["field_linelocs1"])
    assert np.array_equal(np.array(f.linebad, dtype=np.int8), g["field_linebad"])
    np.testing.assert_allclose(f.linelocs2, g["field_linelocs2"], rtol=0, atol=1e-5)
    if system == "NTSC":
        np.testing.assert_allclose(f.linelocs3, g["field_linelocs3"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(f.linelocs4, g["field_linelocs4"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(f.burstlevel, g["field_burstlevel"], rtol=3e-7, atol=0)
    np.testing.assert_allclose(f.linelocs, g["field_linelocs"], rtol=0, atol=1e-5)
    d = f.dspicture.astype(np.int64) - g["field_dspicture"].astype(np.int64)
    assert np.abs(d).max() <= 1                                                      # +-1 LSB of uint16
    assert np.count_nonzero(d) < 0.002 * d.size
    codes = [[-1] * 6 if f.linecode[l] is None else f.linecode[l] for l in rf.SysParams["philips_codelines"]]
    assert np.array_equal(np.array(codes), g["field_linecode"])
    assert f.vbi["framenr"] == int(g["field_framenr"])


@pytest.mark.parametrize("name", ["ntsc", "pal"])
def test_reference_step_methods(backend, golden, name):
    """The reference's per-step methods on the drop-in classes (get_hsync_median, is_regular_hsync, determine_field,
    determine_vsyncs, compute_linelocs, refine_linelocs_hsync, decodephillipscode, processphilipscode,
    refine_linelocs_burst | refine_linelocs_pilot; lddecode_core.py:518-787, 814-884, 962-1021, 1054-1133), called one
    by one the way the reference's constructors chain them, against the vectors recorded from the reference."""
    g = golden(name)
    rf, dd, system = _setup(backend, g, name)
    f = (field.FieldNTSC if system == "NTSC" else field.FieldPAL)(rf, dd, 0)
    med, tol = f.get_hsync_median()
    np.testing.assert_allclose(med, g["field_med_hsync"], rtol=1e-12)
    np.testing.assert_allclose(tol, g["field_hsync_tolerance"], rtol=1e-9)
    vs = f.determine_vsyncs()
    assert [list(map(int, v)) for v in vs] == [list(v) for v in g["field_vsyncs"]]
    # determine_field / is_regular_hsync against the oracle's restatement on every candidate around the first interval
    dec = O.Decoder(float(g["fs_mhz"]), system, int(g["blocklen"]))
    ds = np.zeros(f._n)
    ds[np.array(f.peaklist)] = f._peakvals
    for k in list(range(0, 14)) + list(range(vs[0][0] - 12, vs[0][0] + 12)) + [len(f.peaklist) - 1, len(f.peaklist) + 3]:
        assert f.is_regular_hsync(k) == O._regular(ds, f.peaklist, k, med, tol)
        if k < len(f.peaklist) - 20:
            assert f.determine_field(k) == O._field_vote(dec, ds, f.peaklist, k, med, tol)
    ll1, bad = f.compute_linelocs()
    np.testing.assert_array_equal(np.array(ll1), g["field_linelocs1"])
    f.linelocs1, f.linebad = ll1, bad
    ll2 = f.refine_linelocs_hsync()
    np.testing.assert_allclose(ll2, g["field_linelocs2"], rtol=0, atol=1e-5)
    assert np.array_equal(np.array(f.linebad, dtype=np.int8), g["field_linebad"])
    f.linelocs = ll2
    codes = [f.decodephillipscode(l) for l in rf.SysParams["philips_codelines"]]
    assert np.array_equal(np.array([[-1] * 6 if c is None else c for c in codes]), g["field_linecode"])
    assert f.decodephillipscode(100) is None                                         # a picture line carries no code
    f.linecode = dict(zip(rf.SysParams["philips_codelines"], codes))
    f.processphilipscode()
    assert f.vbi["framenr"] == int(g["field_framenr"])
    if system == "NTSC":
        ll3, bl = f.refine_linelocs_burst(ll2)
        np.testing.assert_allclose(ll3, g["field_linelocs3"], rtol=0, atol=1e-5)
        ll4, bl = f.refine_linelocs_burst(ll3)
        np.testing.assert_allclose(ll4, g["field_linelocs4"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(bl, g["field_burstlevel"], rtol=3e-7, atol=0)
        np.testing.assert_allclose(f.apply_offsets(ll4, 91.5 * (np.pi / 180) - 8), g["field_linelocs"], rtol=0, atol=1e-5)
    else:
        np.testing.assert_allclose(f.refine_linelocs_pilot(), g["field_linelocs"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(f.refine_linelocs_pilot(ll2), g["field_linelocs"], rtol=0, atol=1e-5)


def test_burst_refinement_on_abnormal_line_spans(backend, golden):
    """refine_linelocs_burst on a line table with spans the fast path's staging window does not hold in one piece (1.7 x
    and 3.2 x nominal: the 40 burst samples are then produced in several passes) and a short one (0.6 x), against the
    oracle's restatement of lddecode_core.py:1054-1133 -- the reference resamples whatever span it is given."""
    g = golden("ntsc")
    rf, dd, system = _setup(backend, g, "ntsc")
    f = field.FieldNTSC(rf, dd, 0)
    L = rf.linelen
    ll = np.array(f.linelocs2)
    ll[100:] += 0.7 * L
    ll[150:] += 2.2 * L
    ll[200:] -= 0.4 * L
    dec = O.Decoder(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]))
    cap = g["capture"]
    video, _ = O.demod(dec, lambda a, n: cap[a:a + n] if a + n <= len(cap) else None, 0, int(g["demod_length"]), 1)
    want3, wantbl = O.refine_burst_ntsc(dec, video["demod_burst"], list(ll), f.linecount)
    got3, gotbl = f.refine_linelocs_burst(ll)
    np.testing.assert_allclose(got3, want3, rtol=0, atol=1e-5)
    np.testing.assert_allclose(gotbl, np.asarray(wantbl, dtype=np.float32), rtol=3e-6, atol=0)
    want4, _ = O.refine_burst_ntsc(dec, video["demod_burst"], list(want3), f.linecount)
    got4, _ = f.refine_linelocs_burst(got3)
    np.testing.assert_allclose(got4, want4, rtol=0, atol=1e-5)


@pytest.mark.parametrize("name", ["ntsc", "pal"])
def test_downscale_float_matches_scale(backend, golden, name):
    """Field.downscale in float64 mode equals the reference's lddutils.scale per line."""
    g = golden(name)
    rf, dd, system = _setup(backend, g, name)
    f = field.Field(rf, dd, 0)
    f.linecount = int(g["field_linecount"])
    W = rf.SysParams["outlinelen"]
    out, _ = f.downscale(lineoffset=0, lineinfo=g["field_linelocs"], wow=False, channel="demod")
    for k, l in enumerate(g["scale_lines"]):
        np.testing.assert_allclose(out[l * W:(l + 1) * W], g["scale_out"][k], rtol=0, atol=0.25)     # Hz
    if system == "NTSC":
        sb, _ = f.downscale(lineoffset=0, lineinfo=g["field_linelocs2"], channel="demod_burst")
        np.testing.assert_allclose(sb[::7], g["field_scaledburst_sparse"], rtol=0, atol=0.05)


def test_sync_peaks_bit_exact_random_planes(backend):
    """get_syncpeaks against the oracle on planes that make chains merge late or never:
    noise, silence, and a real sync plane, for several segment sizes and start offsets."""
    fs = 8 * 315 / 88
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend, precision="f64")
    rng = np.random.default_rng(7)
    planes = [rng.uniform(0, 0.5, 300000), np.zeros(100000), rng.uniform(0, 0.21, 200000),
              np.clip(np.sin(np.arange(250000) * 0.00345) + rng.normal(0, .05, 250000), 0, 1)]
    old = os.environ.get("LDD_PEAK_SEG_LINES")
    try:
        for seg in ("4", "48"):
            os.environ["LDD_PEAK_SEG_LINES"] = seg
            for ds in planes:
                for start in (0, 777):
                    buf = backend.to_device(np.ascontiguousarray(ds))
                    pk, vl = field.sync_peaks_device(rf, buf, len(ds), start)
                    ref = O.sync_peaks(ds, start, rf.linelen)
                    assert np.array_equal(pk, np.array(ref, dtype=np.int64))
                    assert np.array_equal(vl, ds[ref] if len(ref) else np.zeros(0))
        # the host restatement the field walk uses for short prefixes returns the same lists
        for ds in planes:
            buf = backend.to_device(np.ascontiguousarray(ds))
            n = 90000
            pk, vl = field.sync_peaks_device(rf, buf[:n], n, 0)
            hpk, hvl = field.sync_peaks_prefix_host(rf, buf[:n], n, {})
            assert np.array_equal(pk, hpk) and np.array_equal(vl, hvl)
        # shorter than two lines: empty list
        buf = backend.to_device(np.ones(3000))
        pk, _ = field.sync_peaks_device(rf, buf, 3000, 0)
        assert len(pk) == 0
    finally:
        if old is None:
            os.environ.pop("LDD_PEAK_SEG_LINES", None)
        else:
            os.environ["LDD_PEAK_SEG_LINES"] = old


def test_field_early_outs(backend):
    """No vsync in the window / a single vsync: same nextfieldoffset rules as Field.__init__
    (lddecode_core.py:909-924), checked against the oracle."""
    from lddecode_b200 import synth
    fs = 8 * 315 / 88
    cap = synth.SynthRF("NTSC", fs, seed=2, lead_lines=120).generate(700000)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=backend, precision="f64")
    dec = O.Decoder(fs, "NTSC", 16384, analog_audio=False)
    ld = lambda s, n: cap[s:s + n] if s + n <= len(cap) else None
    for length in (150000, 500000):          # no vsync at all / exactly one vsync
        dd = rf.demod_device(backend.to_device(cap), _lib.FMT_U8, 0, len(cap), 0, length, 0)
        f = field.FieldNTSC(rf, dd, 0)
        ov, _ = O.demod(dec, ld, 0, length, 0)
        of = O.decode_field(dec, ov, 0)
        assert not f.valid and not of.valid
        assert f.nextfieldoffset == of.nextfieldoffset
        assert np.array_equal(f.peaklist, of.peaklist)


@pytest.mark.parametrize("name", ["ntsc", "pal"])
def test_field_golden_mixed_lane(backend, golden, name):
    """The mixed-precision lane against the same golden fields, at the north-star bars only: peak
    indices bit-exact, line tables identical where they are integers, TBC within +-1 LSB."""
    g = golden(name)
    system = "PAL" if name == "pal" else "NTSC"
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), system, int(g["blocklen"]), precision="mixed", _backend=backend)
    cap = g["capture"]
    dd = rf.demod_device(backend.to_device(cap), _lib.FMT_U8, 0, len(cap), 0, int(g["demod_length"]), 1)
    f = (field.FieldNTSC if system == "NTSC" else field.FieldPAL)(rf, dd, 0)
    assert f.valid
    assert np.array_equal(np.array(f.peaklist), g["field_peaklist"])
    assert np.array_equal(np.array(f.vsyncs), g["field_vsyncs"])
    np.testing.assert_array_equal(np.array(f.linelocs1), g["field_linelocs1"])
    np.testing.assert_allclose(f.linelocs, g["field_linelocs"], rtol=0, atol=2e-3)
    d = f.dspicture.astype(np.int64) - g["field_dspicture"].astype(np.int64)
    assert np.abs(d).max() <= 1


def test_pilot_refine_synthetic_against_oracle(backend):
    """refine_linelocs_pilot (the per-line and the per-field kernel) on synthetic planes, including
    crossings that land exactly on a sample (the reference then also skips the next sample: the kernel's
    sequential re-walk), lines without any crossing, and many equal offsets (median selection with ties)."""
    fs = 35.46895
    rf = rfdecode.RFDecode(fs, "PAL", 16384, _backend=backend, precision="f64")
    dec = O.Decoder(fs, "PAL", 16384, analog_audio=False)
    be = backend
    rng = np.random.default_rng(11)
    linecount = 60
    nll = linecount + 4
    L = rf.linelen
    n = (nll + 3) * L
    for case in range(4):
        k = np.arange(n)
        phase = rng.uniform(0, 2 * np.pi)
        pil = 220000.0 * np.sin(2 * np.pi * 3.75 / fs * k + phase)
        if case == 1:
            pil += rng.normal(0, 30000.0, n)
        if case == 2:
            pil[::7] = 0.0                                   # exact zeros: some close a negative run
        if case == 3:
            pil[: n // 2] = -5.0                              # no crossings on the first lines
        d05 = rng.normal(0, 1000.0, n).astype(np.float32)
        demod = (pil.astype(np.float32) + d05).astype(np.float32)
        if case == 2:
            demod[::7] = d05[::7]                             # demod - demod_05 == 0 exactly
        ll2 = (np.arange(LL := field.LL_STRIDE) * (L + 0.37) + 2 * L + rng.uniform(0, 1)).astype(np.float64)
        video = {"demod": demod.astype(np.float64), "demod_05": d05.astype(np.float64)}
        ref = O.refine_pilot_pal(dec, video, ll2[:nll])
        d_out = be.empty(LL, np.float64)
        d_status = be.zeros(1, np.int32)
        d_demod, d_d05, d_ll = be.to_device(demod), be.to_device(d05), be.to_device(ll2)
        d_base, d_lc = be.to_device(np.zeros(1, dtype=np.int64)), be.to_device(np.array([linecount], dtype=np.int32))
        rf._check(be.lib.ldd_refine_pilot(rf._h, be.ptr(d_demod), be.ptr(d_d05), n, be.ptr(d_base), be.ptr(d_lc), 1, LL,
                                          be.ptr(d_ll), be.ptr(d_out), be.ptr(d_status), be.stream()))
        be.synchronize()
        assert int(be.to_host(d_status)[0]) == 0, case
        np.testing.assert_allclose(be.to_host(d_out)[:nll], ref, rtol=0, atol=1e-9, err_msg=str(case))


def test_tbc_random_geometry_against_oracle_spline(backend):
    """ldd_tbc_fields in float64 mode on random planes and irregular line positions (short, long and
    fractional spans) against the oracle's not-a-knot spline (lddutils.scale)."""
    fs = 8 * 315 / 88
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend, precision="f64")
    be = backend
    rng = np.random.default_rng(5)
    L, W = rf.linelen, rf.SysParams['outlinelen']
    n = 40 * L
    plane = np.cumsum(rng.normal(0, 20000.0, n)).astype(np.float32)
    gaps = np.concatenate([rng.uniform(0.8, 1.2, 12) * L, [0.3 * L, 40.0, 7.5, 1.24 * L]])
    ll = np.zeros(field.LL_STRIDE)
    ll[: len(gaps) + 1] = 500.25 + np.concatenate([[0], np.cumsum(gaps)])
    nlines = len(gaps)
    d_out = be.empty(nlines * W, np.float64)
    d_status = be.zeros(1, np.int32)
    d_plane, d_ll, d_lc = be.to_device(plane), be.to_device(ll), be.to_device(np.array([nlines], dtype=np.int32))
    rf._check(be.lib.ldd_tbc_fields(rf._h, be.ptr(d_plane), n, 0.0, None, be.ptr(d_ll), field.LL_STRIDE,
                                    be.ptr(d_lc), 1, nlines, 0, 0.0, W, 1, 0,
                                    be.ptr(d_out), nlines * W, None, 1.45, be.ptr(d_status), be.stream()))
    be.synchronize()
    assert int(be.to_host(d_status)[0]) == 0
    got = be.to_host(d_out).reshape(nlines, W)
    p64 = plane.astype(np.float64)
    for l in range(nlines):
        ref = O.scale(p64, ll[l], ll[l + 1], W) * ((ll[l + 1] - ll[l]) / L)
        np.testing.assert_allclose(got[l], ref, rtol=0, atol=2e-6 * np.abs(p64).max())


@pytest.mark.parametrize("precision,mode", [("f64", 0), ("mixed", 1), ("f64", 1)])
def test_tbc_long_lines_second_pass(backend, precision, mode):
    """Lines longer than the staging window of the TBC pass (1.25 x nominal; they come from fields whose line location partly
    failed, and lddutils.scale resamples any span): the first pass leaves them out with status 1 | LDD_ST_LINE_LONG and does
    every other line, ldd_tbc_long_lines does exactly those lines (exact kernel, spans up to 4032 samples) and nothing
    else; a span beyond that and a window outside the plane are LDD_ST_LINE_BAD in both passes."""
    fs = 8 * 315 / 88
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend, precision=precision)
    be = backend
    rng = np.random.default_rng(8)
    L, W = rf.linelen, rf.SysParams['outlinelen']
    n = 60 * L
    SP = rf.SysParams
    plane = (np.cumsum(rng.normal(0, 9000.0, n)) * 0.02 + 400000.0 * np.sin(np.arange(n) * 0.01)).astype(np.float32)
    add = float(SP['ire0']) if mode == 1 else 0.0

    def run(gaps, first=500.25):
        ll = np.zeros(field.LL_STRIDE)
        ll[: len(gaps) + 1] = first + np.concatenate([[0], np.cumsum(gaps)])
        nl = len(gaps)
        d_out = be.zeros(nl * W, np.uint16 if mode else np.float64)
        d_plane, d_ll, d_lc = be.to_device(plane), be.to_device(ll), be.to_device(np.array([nl], dtype=np.int32))
        outs = []
        for fn in (be.lib.ldd_tbc_fields_ex, be.lib.ldd_tbc_long_lines):
            d_st = be.zeros(1, np.int32)
            rf._check(fn(rf._h, be.ptr(d_plane), n, add, None, be.ptr(d_ll), field.LL_STRIDE, be.ptr(d_lc), 1, nl, 0, 0.0, W, 1, mode,
                         be.ptr(d_out), nl * W, None, 0, None, 1.45, be.ptr(d_st), be.stream()))
            be.synchronize()
            outs.append((int(be.to_host(d_st)[0]), be.to_host(d_out).reshape(nl, W).copy()))
        return ll, outs

    gaps = [1.0 * L, 1.3 * L, 0.9 * L, 1.9 * L, 4000.0, 1.1 * L, 1.26 * L + 70]
    long_lines = [1, 3, 4, 6]
    ll, ((st1, out1), (st2, out2)) = run(gaps)
    assert st1 == (1 | _lib.ST_LINE_LONG) and st2 == 0
    p64 = plane.astype(np.float64) + add
    for l in range(len(gaps)):
        hz = O.scale(p64, ll[l], ll[l + 1], W) * ((ll[l + 1] - ll[l]) / L)
        if mode == 0:
            ref, tol = hz, 2e-6 * np.abs(p64).max()
        else:
            ire = (hz - SP['ire0']) / SP['hz_ire'] - SP['vsync_ire']
            ref, tol = np.uint16(np.clip(ire * ((0xc800 - 0x0400) / (100 - SP['vsync_ire'])) + 1024, 0, 65535) + 0.5).astype(np.float64), 1.0
        if l in long_lines:
            assert not np.any(out1[l])                                     # left alone by the first pass ...
            assert np.abs(out2[l].astype(np.float64) - ref).max() <= tol   # ... and done by the second
        else:
            assert np.abs(out1[l].astype(np.float64) - ref).max() <= tol
            assert np.array_equal(out1[l], out2[l])                        # untouched by the second pass
    # beyond the exact kernel's reach, or outside the plane: not resampled by either pass
    for gaps, first in (([1.0 * L, 4100.0, 1.0 * L], 500.25), ([1.0 * L, -0.5 * L, 1.0 * L], 500.25), ([1.0 * L, 1.5, 1.0 * L], 500.25),
                        ([1.0 * L, 1.0 * L], n - 1.5 * L)):
        _, ((st1, _), (st2, _)) = run(gaps, first)
        assert st1 == (1 | _lib.ST_LINE_BAD) and st2 in (0, 1 | _lib.ST_LINE_BAD)
    assert st2 == 0                      # (the short lines of the last case are none of the second pass's business)


def test_window_peaks_from_global_matches_own_chase(backend):
    """The host walk cuts a window's peak list out of the capture-wide chase (on a global peak, or off the chain:
    the second read of a capture).  Whenever ldd_window_peaks_from_global decides, its answer must be the list the
    window's own chase produces; and on regular video it must decide for a window that starts 1024 samples before
    a peak (PAL: half a line > blockcut)."""
    import ctypes as C
    lib = backend.lib
    rng = np.random.default_rng(3)
    decided = 0
    for L in (2270, 1820, 2540):
        half, skip = L // 2, int(L * .4)
        n = 120 * L
        for kind in range(4):
            ds = np.zeros(n)
            if kind == 0:        # regular line sync peaks with mild noise
                pos = (np.arange(5, 115) * L + rng.integers(-3, 4, 110)).astype(np.int64)
            elif kind == 1:      # half-line pulses (vsync region) mixed with lines
                pos = np.sort(np.concatenate([np.arange(5, 60) * L, 60 * L + np.arange(0, 40) * half, np.arange(81, 115) * L])).astype(np.int64)
            elif kind == 2:      # irregular
                pos = np.sort(rng.choice(np.arange(2 * L, n - 3 * L), 150, replace=False)).astype(np.int64)
            else:                # noise floor close to the threshold plus peaks
                pos = (np.arange(5, 115) * L).astype(np.int64)
                ds += rng.uniform(0, 0.25, n)
            for p in pos:
                w = np.arange(-40, 41)
                ds[p + w] = np.maximum(ds[p + w], (0.7 + 0.1 * rng.uniform()) * np.exp(-(w / 15.0) ** 2))
            gpk = np.array(O.sync_peaks(ds, 0, L), dtype=np.int64)
            gend = n - 2 * L
            # (a window that starts ON a global peak is taken from the global list under the pipeline's documented
            # assumption that no higher peak follows within half a line: only regular planes are asked that here)
            starts = [int(p) - 1024 for p in gpk[5:60:3]] + list(rng.integers(L, n // 2, 25))
            if kind in (0, 1):
                starts += [int(p) for p in gpk[7:30:5]]
            for b in starts:
                wl = int(rng.integers(20 * L, 60 * L))
                if b < 1 or b + wl > n or (kind >= 2 and b in gpk):
                    continue
                own = np.array(O.sync_peaks(ds[b:b + wl], 0, L), dtype=np.int64) + b
                k0, k1 = C.c_int(0), C.c_int(0)
                ok = lib.ldd_window_peaks_from_global(gpk.ctypes.data_as(C.c_void_p), len(gpk), 0, gend, int(b), wl, L,
                                                      C.byref(k0), C.byref(k1))
                if ok:
                    decided += 1
                    assert np.array_equal(gpk[k0.value:k1.value], own), (L, kind, b, wl)
                elif kind == 0 and L == 2270 and b + 1024 in gpk:
                    raise AssertionError("regular PAL video: a window 1024 samples before a peak must be decided")
    assert decided > 100

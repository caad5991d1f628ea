"""Kernels (1)-(3) + the sync scan of (4): demodblock / demod / audio_phase2 through the C ABI
against the golden vectors recorded from the reference and against the oracle on fresh seeds.

Tolerances (BASELINE.json north_star): demodulated Hz within 1e-4 relative of the float64
reference -- the float64 lane is held to 1e-7 here (float32 plane storage relative to ire0 is
the only loss); demod_sync within 1e-6 absolute; audio within 1e-4 relative (held to 1e-9)."""
import os

import numpy as np
import pytest

from lddecode_b200 import _lib, rfdecode, synth
from oracle import ldd_oracle as O

RTOL_HZ = 1e-7


def _system(g):
    return "PAL" if "demod_pilot" in list(g["planes"]) else "NTSC"


def _mem_loader(cap):
    def ld(infile, sample, n):
        if sample + n > len(cap):
            return None
        return cap[sample:sample + n]
    return ld


def _close_hz(actual, desired, rtol=RTOL_HZ):
    np.testing.assert_allclose(actual, desired, rtol=rtol, atol=0.2)


@pytest.mark.parametrize("name", ["ntsc", "pal", "ntsc10"])
def test_demodblock_golden(backend, golden, name):
    g = golden(name)
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), _system(g), int(g["blocklen"]), _backend=backend, precision="f64")
    cap = g["capture"]
    for bi in range(3):
        pos, mtf = int(g["blk%d_pos" % bi]), float(g["blk%d_mtf" % bi])
        v, a = rf.demodblock(cap[pos:pos + rf.blocklen], mtf)
        for p in g["planes"]:
            ref = g["blk%d_%s" % (bi, p)]
            if p == "demod_sync":
                np.testing.assert_allclose(v[p], ref, rtol=0, atol=1e-6)
            else:
                _close_hz(v[p], ref)
        np.testing.assert_allclose(a["audio_left"], g["blk%d_audio_left" % bi], rtol=1e-9)
        np.testing.assert_allclose(a["audio_right"], g["blk%d_audio_right" % bi], rtol=1e-9)


@pytest.mark.parametrize("name", ["ntsc", "pal", "ntsc10"])
def test_demod_stitched_golden(backend, golden, name):
    g = golden(name)
    N = int(g["blocklen"])
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), _system(g), N, _backend=backend, precision="f64")
    cap = g["capture"]
    rfdecode.loader = _mem_loader(cap)
    video, audio = rf.demod(None, 0, int(g["demod_length"]), 1)
    sp = int(g["sparse"])
    assert len(video) == int(g["demod_len"])
    for p in g["planes"]:
        tol = dict(rtol=0, atol=1e-6) if p == "demod_sync" else dict(rtol=RTOL_HZ, atol=0.2)
        np.testing.assert_allclose(video[p][::sp], g["demod_sparse_" + p], **tol)
        np.testing.assert_allclose(video[p][N - 1056 - 2048:N - 1056 + 2048], g["demod_seam_" + p], **tol)
        np.testing.assert_allclose(video[p][-4096:], g["demod_tail_" + p], **tol)
    np.testing.assert_allclose(audio["audio_left"], g["audio_left"], rtol=1e-9)
    np.testing.assert_allclose(audio["audio_right"], g["audio_right"], rtol=1e-9)
    # start > blockcut and mtf_level = 0 (lddecode_core.py:376-379)
    v2, a2 = rf.demod(None, 54321, 300000, 0)
    assert len(v2) == int(g["demod2_len"])
    for p in g["planes"]:
        tol = dict(rtol=0, atol=1e-6) if p == "demod_sync" else dict(rtol=RTOL_HZ, atol=0.2)
        np.testing.assert_allclose(v2[p][::sp], g["demod2_sparse_" + p], **tol)
    np.testing.assert_allclose(a2["audio_left"], g["demod2_audio_left"], rtol=1e-9)
    # short read -> None, like the reference (lddecode_core.py:386-392)
    assert rf.demod(None, len(cap) - 5000, 100000, 0) is None


def test_runfilter_audio_phase2_one_block(backend, golden):
    """RFDecode.runfilter_audio_phase2 (lddecode_core.py:335-346): one block of the second audio stage.  Block 0 is the head
    of the reference's own audio_phase2 output (it keeps the first block whole, :353-355); any other start against the
    oracle's restatement of the same lines."""
    g = golden("ntsc")
    N = int(g["blocklen"])
    rf = rfdecode.RFDecode(float(g["fs_mhz"]), "NTSC", N, _backend=backend, precision="f64")
    cap = g["capture"]
    a1 = rf.demod_device(backend.to_device(cap), _lib.FMT_U8, 0, len(cap), 0, int(g["demod_length"]), 1, phase2=False).audio_recarray()
    out = rf.runfilter_audio_phase2(a1, 0)
    assert len(out) == N // 4
    np.testing.assert_allclose(out["audio_left"], g["audio_left"][:N // 4], rtol=1e-9)
    np.testing.assert_allclose(out["audio_right"], g["audio_right"][:N // 4], rtol=1e-9)
    dec = O.Decoder(float(g["fs_mhz"]), "NTSC", N)
    want = O._audio2_block(dec, a1, 5001)
    out = rf.runfilter_audio_phase2(a1, 5001)
    np.testing.assert_allclose(out["audio_left"], want["audio_left"], rtol=1e-9)
    np.testing.assert_allclose(out["audio_right"], want["audio_right"], rtol=1e-9)
    with pytest.raises(ValueError):
        rf.runfilter_audio_phase2(a1, len(a1) - N + 1)
    assert rfdecode.calclinelen(rf.SysParams, 4, 'fsc_mhz') == rf.SysParams['outlinelen'] == 910


def test_sync_decisions_bit_exact_vs_oracle(backend):
    """The binary sync decision behind demod_sync (lddecode_core.py:308) must be the reference's,
    sample for sample: recover it from demod_05 and compare with the oracle on a fresh seed."""
    fs = 8 * 315 / 88
    cap = synth.SynthRF("NTSC", fs, seed=11).generate(200000)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend, precision="f64")
    rfdecode.loader = _mem_loader(cap)
    video, _ = rf.demod(None, 0, 150000, 0)
    dec = O.Decoder(fs, "NTSC", 16384)
    ov, _ = O.demod(dec, lambda s, n: cap[s:s + n] if s + n <= len(cap) else None, 0, 150000, 0)
    np.testing.assert_allclose(video["demod_sync"], ov["demod_sync"], rtol=0, atol=1e-6)
    assert np.array_equal(O.sync_peaks(video["demod_sync"], 0, rf.linelen), O.sync_peaks(ov["demod_sync"], 0, dec.linelen))


@pytest.mark.parametrize("fmt", ["r30", "lds"])
def test_packed_input_formats(backend, fmt):
    """10-bit packed captures demodulate straight from the packed bytes (unpack fused into the
    block load) to the same planes as the unpacked samples."""
    import ctypes as C
    from lddecode_b200 import _lib
    fs = 8 * 315 / 88
    s10 = synth.SynthRF("NTSC", fs, seed=5, bits=10).generate(60000)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend, precision="f64")
    start, length = 4000, 30000
    ref = rf.demod_device(backend.to_device(s10), _lib.FMT_U16, 0, len(s10), start, length, 0, phase2=False).to_recarrays()
    if fmt == "r30":
        packed, f = synth.pack_r30(s10), _lib.FMT_R30
    else:
        packed, f = synth.pack_lds(s10), _lib.FMT_LDS40
    out = rf.demod_device(backend.to_device(packed), f, 0, len(s10) // 12 * 12, start, length, 0, phase2=False).to_recarrays()
    for p in ("demod", "demod_05", "demod_sync", "demod_burst"):
        assert np.array_equal(out[0][p], ref[0][p]), p
    assert np.array_equal(out[1]["audio_left"], ref[1]["audio_left"])


def test_unpack_kernels_bit_exact(backend, golden):
    import ctypes as C
    from lddecode_b200 import _lib
    g = golden("unpack")
    lib = backend.lib
    words = backend.to_device(g["r30_words"])
    out = backend.empty(len(g["r30_words"]) * 3, np.int16)
    assert lib.ldd_unpack_r30_ddunpack(backend.ptr(words), len(g["r30_words"]), backend.ptr(out), backend.stream()) == 0
    backend.synchronize()
    assert np.array_equal(backend.to_host(out), g["r30_ddunpack_i16"])          # == compiled ddunpack.c
    for s, exp in zip(g["r30_py_starts"], g["r30_py"]):
        o = backend.empty(1000, np.uint16)
        assert lib.ldd_unpack_raw(backend.ptr(words), _lib.FMT_R30, int(s), 1000, backend.ptr(o), backend.stream()) == 0
        backend.synchronize()
        assert np.array_equal(backend.to_host(o).astype(np.int16), exp)        # == load_packed_data_3_32
    lds = backend.to_device(g["lds_bytes"])
    for s, exp in zip(g["lds_py_starts"], g["lds_py"]):
        o = backend.empty(1000, np.uint16)
        assert lib.ldd_unpack_raw(backend.ptr(lds), _lib.FMT_LDS40, int(s), 1000, backend.ptr(o), backend.stream()) == 0
        backend.synchronize()
        assert np.array_equal(backend.to_host(o), exp)                         # == load_packed_data_4_40
        of = backend.empty(1001, np.float32)
        assert lib.ldd_unpack_f32(backend.ptr(lds), _lib.FMT_LDS40, int(s), 999, backend.ptr(of), backend.stream()) == 0
        backend.synchronize()
        assert np.array_equal(backend.to_host(of)[:999], exp[:999].astype(np.float32))
    # aligned bulk path (whole packing groups, vector kernels) + ragged remainder, against the oracle
    rng = np.random.default_rng(3)
    s10 = rng.integers(0, 1024, 12 * 64 * 5 + 7, dtype=np.uint16)
    for fmt, packed, first in ((_lib.FMT_R30, synth.pack_r30(s10), 24), (_lib.FMT_LDS40, synth.pack_lds(s10), 128)):
        pk = backend.to_device(packed)
        n = len(s10) - first - 3
        o = backend.empty(n, np.uint16)
        assert lib.ldd_unpack_raw(backend.ptr(pk), fmt, first, n, backend.ptr(o), backend.stream()) == 0
        of = backend.empty(n, np.float32)
        assert lib.ldd_unpack_f32(backend.ptr(pk), fmt, first, n, backend.ptr(of), backend.stream()) == 0
        backend.synchronize()
        assert np.array_equal(backend.to_host(o), s10[first:first + n])
        assert np.array_equal(backend.to_host(of), s10[first:first + n].astype(np.float32))
    # ragged tail and empty input
    o = backend.zeros(16, np.int16)
    assert lib.ldd_unpack_r30_ddunpack(backend.ptr(words), 5, backend.ptr(o), backend.stream()) == 0
    backend.synchronize()
    assert np.array_equal(backend.to_host(o)[:15], g["r30_ddunpack_i16"][:15]) and backend.to_host(o)[15] == 0
    assert lib.ldd_unpack_r30_ddunpack(backend.ptr(words), 0, backend.ptr(o), backend.stream()) == 0


@pytest.mark.parametrize("name", ["ntsc", "pal"])
def test_mixed_lane_sync_is_the_float64_lane(backend, golden, name):
    """precision='mixed' (float32 shared-memory pass + float64 re-run of every block with a demod_05
    sample within the 16 Hz guard band of a sync threshold): demod_sync must be BIT-identical to the
    float64 lane, the float32 error must stay far inside the guard band, and the Hz planes inside
    the 1e-4 relative bar."""
    from lddecode_b200 import _lib
    g = golden(name)
    cap = g["capture"]
    capd = backend.to_device(cap)
    out = {}
    for prec in ("f64", "mixed"):
        rf = rfdecode.RFDecode(float(g["fs_mhz"]), _system(g), int(g["blocklen"]), precision=prec, _backend=backend)
        dd = rf.demod_device(capd, _lib.FMT_U8, 0, len(cap), 0, int(g["demod_length"]), 1, phase2=False)
        out[prec] = {k: backend.to_host(v) for k, v in dd.planes.items()}
        if prec == "mixed":
            flagged, total = rf.mixed_stats()
            assert 0 <= flagged <= total // 4
    assert np.array_equal(out["mixed"]["demod_sync"], out["f64"]["demod_sync"])
    err05 = np.abs(out["mixed"]["demod_05"].astype(np.float64) - out["f64"]["demod_05"])
    assert err05.max() < 4.0                      # Hz; guard band is 16 Hz
    ire0 = 8100000.0 if _system(g) == "NTSC" else 7100000.0
    rel = np.abs(out["mixed"]["demod"].astype(np.float64) - out["f64"]["demod"]) / (np.abs(out["f64"]["demod"].astype(np.float64) + ire0))
    assert rel.max() < 1e-6                       # bar: 1e-4


def test_inplace_transforms_against_host_fft(tmp_path):
    """The in-place DIF / DIT transform pair of the default block length (ldd_fft2.cuh) and the Stockham plan
    (ldd_fft.cuh), each against a float64 host FFT: spectrum (through the digit permutation) and round trip.
    tools/fft_bench.cu is the same program the B200 micro-benchmark runs; here it is compiled for the CPU emulation."""
    import re
    import subprocess
    exe = str(tmp_path / "fft_bench_emu")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-DLDD_EMU", "-I", os.path.join(root, "tests", "emu"), "-I",
                           os.path.join(root, "lddecode_b200", "csrc"), "-x", "c++", os.path.join(root, "tools", "fft_bench.cu"),
                           os.path.join(root, "tests", "emu", "cuda_emu.cpp"), "-o", exe, "-Wno-unknown-pragmas"])
    out = subprocess.run([exe, "1", "1"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = re.findall(r"(\S+)\s+check: spectrum max err (\S+) \(max \|X\| (\S+)\), round trip max err (\S+)", out)
    assert [r[0] for r in rows] == ["stockham", "inplace/1", "inplace/2"], out
    for name, e1, mx, e2 in rows:
        assert float(e1) < 1e-6 * 8192 ** 0.5 * float(mx) and float(e2) < 1e-5, (name, e1, mx, e2)


@pytest.mark.parametrize("system,audio", [("PAL", False), ("NTSC", True)])
def test_inplace_block_equals_stockham_block(backend, system, audio, monkeypatch):
    """The float32 block of the default block length on in-place transforms (ldd_demod8k.cuh) against the Stockham block
    it replaced (LDD_STOCKHAM_BLOCK=1): planes and audio equal to float32 rounding, the sync plane bit for bit; and the
    mixed lane's float64 re-run on in-place transforms against the generic float64 block (LDD_STOCKHAM_RERUN=1) with a
    guard band wide enough to send many blocks through it: sync plane bit for bit, demod_05 to the float32 ulp (the two
    float64 computations differ by ~1e-16 relative, which moves a float32 rounding now and then)."""
    from lddecode_b200 import _lib
    fs = 8 * 315 / 88 if system == "NTSC" else 35.46895
    cap = synth.SynthRF(system, fs, seed=3).generate(150000)
    dev = backend.to_device(cap)

    def run(prec):
        rf = rfdecode.RFDecode(fs, system, 16384, decode_analog_audio=audio, _backend=backend, precision=prec)
        out = rf.demod_device(dev, _lib.FMT_U8, 0, len(cap), 2000, 120000, 1 if audio else 0, phase2=False).to_recarrays()
        return out, rf

    monkeypatch.delenv("LDD_STOCKHAM_BLOCK", raising=False)
    new, _ = run("f32")
    monkeypatch.setenv("LDD_STOCKHAM_BLOCK", "1")
    old, _ = run("f32")
    monkeypatch.delenv("LDD_STOCKHAM_BLOCK")
    assert np.array_equal(new[0]["demod_sync"], old[0]["demod_sync"])
    for name in new[0].dtype.names:
        if name != "demod_sync":
            np.testing.assert_allclose(new[0][name], old[0][name], rtol=1e-6, atol=1.0, err_msg=name)
    if audio:
        for name in new[1].dtype.names:
            np.testing.assert_allclose(new[1][name], old[1][name], rtol=2e-6, err_msg=name)
    monkeypatch.setenv("LDD_FLAG_MARGIN_HZ", "400")
    a, rfa = run("mixed")
    flagged, total = rfa.mixed_stats()
    assert 0 < flagged <= total
    monkeypatch.setenv("LDD_STOCKHAM_RERUN", "1")
    b, _ = run("mixed")
    assert np.array_equal(a[0]["demod_sync"], b[0]["demod_sync"])
    np.testing.assert_allclose(a[0]["demod_05"], b[0]["demod_05"], rtol=2e-7, atol=0)


def test_packed_captures_equal_unpacked_in_the_default_lane(backend):
    """The float32 block of the default lane fetches 8-bit and .lds samples as aligned words and everything else sample by
    sample, but its arithmetic is the same for every format and alignment: .lds, .r30 and 16-bit captures of the same
    10-bit samples give bit-equal planes and audio, and so does a capture window that starts at an odd byte."""
    from lddecode_b200 import _lib
    fs = 8 * 315 / 88
    s10 = synth.SynthRF("NTSC", fs, seed=5, bits=10).generate(120000)
    n = len(s10) // 12 * 12
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=True, _backend=backend, precision="mixed")
    outs = {}
    for name, packed, f in (("u16", s10, _lib.FMT_U16), ("lds", synth.pack_lds(s10), _lib.FMT_LDS40), ("r30", synth.pack_r30(s10), _lib.FMT_R30)):
        outs[name] = rf.demod_device(backend.to_device(packed), f, 0, n, 4000, 90000, 1, phase2=False).to_recarrays()
    for name in ("lds", "r30"):
        for p in outs["u16"][0].dtype.names:
            assert np.array_equal(outs[name][0][p], outs["u16"][0][p]), (name, p)
        assert np.array_equal(outs[name][1]["audio_left"], outs["u16"][1]["audio_left"]), name
    # 8-bit: word-aligned window against the same samples at an odd offset of a longer buffer (sample-by-sample fetch)
    s8 = synth.SynthRF("NTSC", fs, seed=6).generate(100001)
    a = rf.demod_device(backend.to_device(s8[1:]), _lib.FMT_U8, 0, 100000, 2000, 70000, 1, phase2=False).to_recarrays()
    dev = backend.to_device(s8)
    b = rf.demod_device(dev[1:], _lib.FMT_U8, 0, 100000, 2000, 70000, 1, phase2=False).to_recarrays()
    for p in a[0].dtype.names:
        assert np.array_equal(a[0][p], b[0][p]), p

"""Whole-capture pipeline against the reference's field-by-field flow (oracle), and sharding."""
import os
import subprocess
import sys

import numpy as np
import pytest

from lddecode_b200 import _lib, pipeline, rfdecode, synth
from oracle import ldd_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_walk(dec, cap, nmax=10):
    """Framer.readfield's loop (lddecode_core.py:1194-1223) with the oracle: list of FieldResult."""
    ld = lambda s, n: cap[s:s + n] if s + n <= len(cap) else None
    out, readsample = [], 0
    while len(out) < nmax:
        d = O.demod(dec, ld, readsample, 1000000, 1)
        if d is None:
            break
        f = O.decode_field(dec, d[0], 0)
        out.append((readsample, f))
        readsample += f.nextfieldoffset
    return out


@pytest.mark.parametrize("precision", ["f64", "mixed"])
@pytest.mark.parametrize("system", ["NTSC", "PAL"])
def test_pipeline_matches_reference_flow(backend, system, precision):
    fs = 8 * 315 / 88 if system == "NTSC" else 35.46895
    ncap = 2100000 if system == "NTSC" else 2600000
    cap = synth.SynthRF(system, fs, seed=9).generate(ncap)
    rf = rfdecode.RFDecode(fs, system, 16384, decode_analog_audio=False, _backend=backend, precision=precision)
    # the exact lane is held to 1e-5 samples / 0.2 % of TBC samples differing; the mixed (default) lane to
    # the north-star bars: identical peak lists and integer line tables, TBC within +-1 LSB
    ll_tol, frac = (1e-5, 0.002) if precision == "f64" else (2e-3, 1.0)
    dec = O.Decoder(fs, system, 16384, analog_audio=False)
    cd = pipeline.CaptureDecoder(rf)
    res = cd.decode(backend.to_device(cap), _lib.FMT_U8, ncap, want_tables=True)
    pics = cd.pictures(res)
    ref = _oracle_walk(dec, cap)
    assert res.nwindows == len(ref) == 3
    for k, (readsample, f) in enumerate(ref):
        info = res.infos[k]
        assert int(res.readsamples[k]) == readsample
        assert info.stage == _lib.FIELD_LOCATED and f.valid
        assert info.nextfieldoffset == f.nextfieldoffset and info.npeaks == len(f.peaklist)
        assert [[info.vsyncs[i][q] for q in range(3)] for i in range(2)] == [list(v) for v in f.vsyncs[:2]]
        nll = f.linecount + 4
        np.testing.assert_array_equal(res.linelocs1[k][:nll], np.array(f.linelocs1))
        j = res.located.index(k)
        np.testing.assert_allclose(res.refined.linelocs2[j][:nll], f.linelocs2, rtol=0, atol=ll_tol)
        np.testing.assert_allclose(res.refined.final[j][:nll] + res.refined.lineloc_add, f.linelocs, rtol=0, atol=ll_tol)
        d = pics[j][2].astype(np.int64) - f.dspicture.astype(np.int64)
        assert np.abs(d).max() <= 1 and np.count_nonzero(d) <= frac * d.size


@pytest.mark.parametrize("system", ["NTSC", "PAL"])
def test_pipeline_at_40_msps_from_packed_lds(backend, system):
    """The reference's default sample rate (RFDecode(inputfreq=40), lddecode_core.py:120) and the native rate of the
    Domesday Duplicator's packed .lds format (lddutils.py:190-229): line length, filter tables, audio decimation (64) and
    every staging window differ from the 8fsc cases.  Whole chain in the default lane from packed bytes, both audio
    channels on, against the oracle walk: read positions, peak counts and integer line tables equal, TBC +-1 LSB,
    PCM of the first window +-1 LSB."""
    fs = 40.0
    ncap = int(fs * 1e6 / (30 if system == "NTSC" else 25) * 1.65) // 4 * 4
    s10 = synth.SynthRF(system, fs, seed=5, bits=10).generate(ncap)
    rf = rfdecode.RFDecode(fs, system, 16384, _backend=backend)
    cd = pipeline.CaptureDecoder(rf)
    res = cd.decode(backend.to_device(synth.pack_lds(s10)), _lib.FMT_LDS40, ncap)
    pics = cd.pictures(res)
    dec = O.Decoder(fs, system, 16384, analog_audio=True)
    ref = _oracle_walk(dec, s10)
    assert res.nwindows == len(ref) >= 2 and len(res.located) == len(ref)
    for k, (readsample, f) in enumerate(ref):
        info = res.infos[k]
        assert int(res.readsamples[k]) == readsample and f.valid and info.stage == _lib.FIELD_LOCATED
        assert info.nextfieldoffset == f.nextfieldoffset and info.npeaks == len(f.peaklist)
        nll = f.linecount + 4
        np.testing.assert_array_equal(res.linelocs1[k][:nll], np.array(f.linelocs1))
        np.testing.assert_allclose(res.refined.final[k][:nll] + res.refined.lineloc_add, f.linelocs, rtol=0, atol=2e-3)
        d = pics[k][2].astype(np.int64) - f.dspicture.astype(np.int64)
        assert np.abs(d).max() <= 1
    assert not np.any(res.refined.status)
    pcm, _, _ = cd.pcm(res, chain='fields')
    ld = lambda a, n: s10[a:a + n] if a + n <= ncap else None
    want, _ = O.downscale_audio(dec, O.demod(dec, ld, 0, 1000000, 1)[1], ref[0][1].linelocs, ref[0][1].linecount, 0.0)
    assert len(pcm[0]) == len(want) and np.abs(pcm[0].astype(np.int64) - want.astype(np.int64)).max() <= 1


@pytest.mark.parametrize("system,nranges", [("NTSC", 2), ("PAL", 1)])
def test_pipeline_pcm_against_framer_audio(backend, system, nranges):
    """48 kHz PCM of a range (ldd_pipe_pcm) against the reference's per-field PCM chained as Framer.readframe chains it
    (oracle.framer_audio_walk, pinned to the reference's own Framer in test_oracle_golden).  Sample counts, dropped fields
    and the carried time offset are exact.  Values: the reference resamples each field from the phase-2 audio of that
    field's own read window; the pipeline reads the same instants out of the range's audio by four-point interpolation.
    A window that starts on the range's audio grid (the first) is the reference's within +-1 LSB; elsewhere the two
    differ where the reference's own block-edge ripple sits (its per-window block grid is not the range's): measured
    <= 21 LSB of int16 (2e-3 of the tone), mean < 0.9 LSB.  The state carries over from one range to the next."""
    from lddecode_b200 import parallel
    fs = 8 * 315 / 88 if system == "NTSC" else 35.46895
    ncap = 2600000 if system == "NTSC" else 3300000
    cap = synth.SynthRF(system, fs, seed=9).generate(ncap)
    rf = rfdecode.RFDecode(fs, system, 16384, _backend=backend)              # default lane, analog audio on
    cd = pipeline.CaptureDecoder(rf)
    got, offset, state = [], 0.0, 2
    for r0, r1 in parallel.shard_bounds(ncap, nranges):
        lo, hi = parallel.needed_window(cd, ncap, r0, r1)
        res = cd.decode_range(backend.to_device(cap[lo:hi]), _lib.FMT_U8, lo, hi - lo, ncap, r0, r1)
        pcm, offset, state = cd.pcm(res, offset, state)
        got += [(int(res.readsamples[j]), p) for j, p in zip(res.located, pcm)]
    dec = O.Decoder(fs, system, 16384, analog_audio=True)
    ld = lambda s, n: cap[s:s + n] if s + n <= len(cap) else None
    ref, ref_offset = O.framer_audio_walk(dec, ld, len(got))
    assert len(got) == len(ref) == 4
    assert offset == ref_offset
    for k, ((rs, p), (ors, f, op)) in enumerate(zip(got, ref)):
        assert rs == ors
        assert (p is None) == (op is None)
        if op is None:
            continue
        assert p.dtype == np.int16 and len(p) == len(op)
        d = np.abs(p.astype(np.int64) - op.astype(np.int64))
        if rs == 0:
            assert d.max() <= 1
        else:
            assert d.max() <= 32 and d.mean() <= 1.5, (k, d.max(), d.mean())
    # every field continuing the previous one: the offsets numpy's arange gives
    res = cd.decode(backend.to_device(cap), _lib.FMT_U8, ncap)
    # scale=None: the decoder's true audio decimation.  The synthetic left channel is a 1 kHz tone; with the reference's
    # fixed scale=64 an 8fsc capture (true factor 32) plays at half speed, a 35.5 MSPS one (64) at speed.
    assert rf.audio_decimation == (32 if system == "NTSC" else 64)
    tone = lambda p: np.argmax(np.abs(np.fft.rfft(p[0::2].astype(np.float64) * np.hanning(len(p) // 2)))[1:]) + 1
    k64, ktrue = [(lambda p: tone(p) * 48000.0 * 2 / len(p))(cd.pcm(res, 0.0, 2, chain='fields', scale=sc)[0][1]) for sc in (64, None)]
    assert abs(ktrue - 1000.0) < 70.0 and abs(k64 - (500.0 if system == "NTSC" else 1000.0)) < 70.0
    pcm, offset, _ = cd.pcm(res, 0.0, 2, chain='fields')
    t = 0.0
    for j, p in zip(res.located, pcm):
        frametime = rf.SysParams['line_period'] * int(res.infos[j].linecount) / 1000000
        ar = np.arange(t, frametime + 1 / 48000.0, 1 / 48000.0, dtype=np.double)
        assert p is not None and len(p) == 2 * (len(ar) - 1)
        t = ar[-1] - frametime
    assert offset == t


def test_error_behaviour_of_the_c_abi(backend):
    """Calls that cannot be served fail with an LDD_E* code and a message instead of decoding something else: sample rates
    the staging windows do not cover, tables of the wrong size, a second audio stage shorter than a block (the reference
    raises there, lddecode_core.py:337), PCM without audio / before a finish / with an unknown chain, a PCM buffer that
    is too small."""
    import ctypes as C
    fs = 8 * 315 / 88
    for bad_fs in (42.0, 45.0):
        with pytest.raises(_lib.LddError):
            rfdecode.RFDecode(bad_fs, "NTSC", 16384, _backend=backend)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend)
    lib = backend.lib
    with pytest.raises(_lib.LddError):
        rf._set_filter(_lib.F_AUDIO_LPF2, np.zeros(100, dtype=np.complex128))
    with pytest.raises(_lib.LddError):
        rf.audio_phase2(np.rec.array([np.zeros(1000), np.zeros(1000)], names=['audio_left', 'audio_right']))
    ao, st = C.c_double(0.0), C.c_int(2)
    assert lib.ldd_pcm_chain(0, 48000.0, 63.5, 7, 0, None, None, C.byref(ao), C.byref(st), None) == -1        # LDD_EINVAL: chain
    assert lib.ldd_pcm_chain(0, 0.0, 63.5, 0, 0, None, None, C.byref(ao), C.byref(st), None) == -1
    assert lib.ldd_pcm_chain(0, 48000.0, 63.5, 0, 0, None, None, C.byref(ao), C.byref(st), None) == 0 and ao.value == 0.0
    cap = synth.SynthRF("NTSC", fs, seed=3).generate(1100000)
    rf_mute = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=backend)
    cd = pipeline.CaptureDecoder(rf_mute)
    res = cd.decode(backend.to_device(cap), _lib.FMT_U8, len(cap))
    assert len(res.located) == 1
    with pytest.raises(ValueError):
        cd.pcm(res)                                             # decoded without the analog audio
    cd = pipeline.CaptureDecoder(rf)
    slot = cd._slot(0, len(cap))
    off = (C.c_longlong * 8)()
    buf, stat = backend.empty(16, np.int16), backend.zeros(8, np.int32)
    args = (48000.0, 64.0, 63.5, 2.3e6, 2.8e6, 1, C.byref(ao), C.byref(st), backend.ptr(buf), 16, off, backend.ptr(stat), backend.stream())
    assert lib.ldd_pipe_pcm(slot.h, *args) == -1                # no finish yet
    res = cd.decode(backend.to_device(cap), _lib.FMT_U8, len(cap))
    assert lib.ldd_pipe_pcm(slot.h, *args) == _lib.ECAP         # 16 int16 do not hold a field's samples
    assert ao.value == 0.0 and st.value == 2                    # ... and the state is untouched
    assert len(cd.pcm(res, chain='fields')[0][0]) == 1606


def test_empty_and_ragged_captures(backend):
    """Captures that are empty, shorter than a block, end off the block grid or are a sample short of the first read
    (Framer.readfield returns None when RFDecode.demod does, lddecode_core.py:386-392, 1199-1201): no window, no crash;
    the first window appears exactly where the oracle's demod of the first read succeeds."""
    fs = 8 * 315 / 88
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend)
    cd = pipeline.CaptureDecoder(rf)
    cap = synth.SynthRF("NTSC", fs, seed=3).generate(1100000)
    dec = O.Decoder(fs, "NTSC", 16384, analog_audio=False)
    stride = 16384 - 1024 - dec.blockcut_end
    need = (1000001 // stride) * stride + 16384                              # the last block demod(0, 1e6) reads ends here
    for n in (0, 1, 100, 16383, 16384, 16385, 40001, 999999, 1000001, need - 1, need, need + 1, 1100000):
        res = cd.decode(backend.to_device(cap[:max(n, 1)]), _lib.FMT_U8, n)
        want = 1 if n >= need else 0
        assert res.nwindows == want, n
        assert len(cd.pictures(res)) == len(res.located) == want
    for n, ok in ((need - 1, False), (need, True)):
        ld = lambda s, k: cap[s:s + k] if s + k <= n else None
        assert (O.demod(dec, ld, 0, 1000000, 1) is not None) == ok


def test_ranges_are_bit_identical_to_one_range(backend):
    """Two read-position ranges (what two GPUs or two chunks would do), the second holding only the
    part of the capture it needs, give exactly the fields of the single-range decode."""
    from lddecode_b200 import parallel
    fs = 8 * 315 / 88
    ncap = 2600000
    cap = synth.SynthRF("NTSC", fs, seed=9).generate(ncap)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=backend, precision="f64")
    cd = pipeline.CaptureDecoder(rf)
    one = cd.pictures(cd.decode(backend.to_device(cap), _lib.FMT_U8, ncap))
    parts = []
    for r0, r1 in parallel.shard_bounds(ncap, 2):
        lo, hi = parallel.needed_window(cd, ncap, r0, r1)
        res = cd.decode_range(backend.to_device(cap[lo:hi]), _lib.FMT_U8, lo, hi - lo, ncap, r0, r1)
        parts += cd.pictures(res)
    assert len(parts) == len(one) == 4
    for a, b in zip(parts, one):
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_pipelined_ranges_equal_single_decode(backend):
    """decode_pipelined (ranges on their own handles / streams, host walk overlapped with device work)
    returns exactly the fields of decode()."""
    fs = 8 * 315 / 88
    ncap = 2600000
    cap = synth.SynthRF("NTSC", fs, seed=9).generate(ncap)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=backend, precision="f64")
    cd = pipeline.CaptureDecoder(rf)
    capd = backend.to_device(cap)
    one = cd.pictures(cd.decode(capd, _lib.FMT_U8, ncap))
    parts = []
    for res in cd.decode_pipelined(capd, _lib.FMT_U8, ncap, nranges=2):
        parts += cd.pictures(res)
    assert len(parts) == len(one) == 4
    for a, b in zip(parts, one):
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])


def test_host_stream_decoder_equals_decode(backend):
    """HostStreamDecoder (uploads / downloads on their own streams, double-buffered) returns for every
    chunk exactly what decode() of that chunk returns, also when buffers are reused."""
    fs = 8 * 315 / 88
    n = 1300000
    caps = [synth.SynthRF("NTSC", fs, seed=s).generate(n) for s in (9, 10)]
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=backend, precision="mixed")
    cd = pipeline.CaptureDecoder(rf)
    want = cd.pictures(cd.decode(backend.to_device(caps[1]), _lib.FMT_U8, n))
    pins = []
    for c in caps:
        p = backend.pinned(n, np.uint8)
        backend.host_view(p)[:] = c
        pins.append(p)
    sd = pipeline.HostStreamDecoder(cd, _lib.FMT_U8, n, max_fields=8)
    got = []
    for res, pics in sd.run([(pins[0], n), (pins[1], n), (pins[0], n)]):
        assert not np.any(res.status_host & 15)
        got.append([(int(res.readsamples[j]), int(res.infos[j].istop), pics[k, :res.infos[j].linecount * 910].copy())
                    for k, j in enumerate(res.located)])
    assert len(got) == 3 and len(got[1]) == len(want) >= 1
    for a, b in zip(got[1], want):
        assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])
    assert len(got[0]) == len(got[2])
    for a, b in zip(got[0], got[2]):
        assert a[0] == b[0] and np.array_equal(a[2], b[2])
    assert not np.array_equal(got[0][0][2], got[1][0][2])


def test_host_stream_decoder_downloads_audio(backend):
    """With analog audio decoding on, HostStreamDecoder brings the two phase-2 channels to the host with the
    fields; they equal the device result of decode() on the same chunk."""
    fs = 8 * 315 / 88
    n = 1300000
    cap = synth.SynthRF("NTSC", fs, seed=12).generate(n)
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=True, _backend=backend, precision="mixed")
    cd = pipeline.CaptureDecoder(rf)
    ref = cd.decode(backend.to_device(cap), _lib.FMT_U8, n)
    want = (backend.to_host(ref.audio['audio_left']).copy(), backend.to_host(ref.audio['audio_right']).copy())
    p = backend.pinned(n, np.uint8)
    backend.host_view(p)[:] = cap
    sd = pipeline.HostStreamDecoder(cd, _lib.FMT_U8, n, max_fields=8)
    outs = list(sd.run([(p, n), (p, n)]))
    assert len(outs) == 2
    for res, pics in outs:
        assert res.audio_host is not None and len(res.audio_host[0]) == len(want[0]) > 1000
    res, _ = outs[-1]
    assert np.array_equal(res.audio_host[0], want[0]) and np.array_equal(res.audio_host[1], want[1])


@pytest.mark.parametrize("fmt", ["lds", "u8"])
def test_file_stream_decoder_equals_whole_decode(backend, tmp_path, fmt):
    """FileStreamDecoder (reader thread -> page-locked buffers -> HostStreamDecoder, the file worked through as
    consecutive read-position ranges of one capture; lddutils.py:131-229 loaders, lddecode.py:88-98 loop) yields exactly
    the fields of decoding the whole file at once, in order, from packed .lds bytes and from 8-bit samples."""
    fs = 8 * 315 / 88
    n = 2400000
    if fmt == "lds":
        s = synth.SynthRF("NTSC", fs, seed=21, bits=10).generate(n)
        raw, f, path = synth.pack_lds(s), _lib.FMT_LDS40, tmp_path / "cap.lds"
    else:
        s = synth.SynthRF("NTSC", fs, seed=21).generate(n)
        raw, f, path = s, _lib.FMT_U8, tmp_path / "cap.u8"
    raw.tofile(str(path))
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, decode_analog_audio=False, _backend=backend, precision="mixed")
    cd = pipeline.CaptureDecoder(rf)
    want = cd.pictures(cd.decode(backend.to_device(raw), f, n))
    assert len(want) >= 3
    fsd = pipeline.FileStreamDecoder(cd, str(path), chunk_samples=700000)
    assert fsd.fmt == f and fsd.ncap == n and len(fsd.ranges) == 4
    got = []
    for res, pics in fsd:
        assert not np.any(res.status_host & 15)
        got += [(int(res.readsamples[j]), int(res.infos[j].istop), pics[k, :res.infos[j].linecount * 910].copy())
                for k, j in enumerate(res.located)]
    assert [g[0] for g in got] == [w[0] for w in want]
    for a, b in zip(got, want):
        assert a[1] == b[1] and np.array_equal(a[2], b[2])


def test_file_stream_decoder_pcm(backend, tmp_path):
    """FileStreamDecoder(pcm=True): the 48 kHz PCM of the file's fields, chained from chunk to chunk as Framer.readframe
    chains it -- the sample counts, the dropped fields and the final offset of the whole-capture decode; values within the
    block-edge ripple of the second audio stage (its blocks are anchored per range)."""
    fs = 8 * 315 / 88
    n = 2400000
    s = synth.SynthRF("NTSC", fs, seed=21, bits=10).generate(n)
    raw, path = synth.pack_lds(s), tmp_path / "cap.lds"
    raw.tofile(str(path))
    rf = rfdecode.RFDecode(fs, "NTSC", 16384, _backend=backend)
    cd = pipeline.CaptureDecoder(rf)
    res = cd.decode(backend.to_device(raw), _lib.FMT_LDS40, n)
    want, want_off, want_state = cd.pcm(res)
    assert len(want) >= 3 and sum(w is not None for w in want) >= 3
    fsd = pipeline.FileStreamDecoder(cd, str(path), chunk_samples=700000, pcm=True)
    got = []
    for r, pics in fsd:
        assert not np.any(r.status_host & 15)
        got += [None if p is None else p.copy() for p in r.pcm_host]
    assert fsd.sd.pcm_state == (want_off, want_state)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if a is not None:
            assert len(a) == len(b)
            d = np.abs(a.astype(np.int64) - b.astype(np.int64))
            assert d.max() <= 32 and d.mean() <= 1.5


def test_two_rank_gloo_gather():
    """world_size 2 over gloo on the CPU (emulated kernels): shard, decode, gather on rank 0; the sharded capture's PCM
    with the offset chain carried across the ranks (parallel.sharded_pcm)."""
    script = os.path.join(ROOT, "tests", "dist_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "GATHER_OK 4" in r.stdout, r.stdout[-3000:]
    assert "PCM_OK 4" in r.stdout, r.stdout[-3000:]


def test_two_rank_strong_scaling_gloo():
    """world_size 2 over gloo: one tiled capture sharded by block range, chunked per rank, gathered per chunk."""
    script = os.path.join(ROOT, "tests", "dist_strong_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "STRONG_OK" in r.stdout, r.stdout[-3000:]

#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference has no tests or fixtures of its own for this path, so these
files are the parity pin: outputs of the reference's own numpy/scipy code (lddecode_core.py
RFDecode.demodblock / demod / audio_phase2, Field / FieldNTSC / FieldPAL; lddutils loaders;
ddunpack.c compiled by oracle/Makefile) on seeded synthetic RF from lddecode_b200/synth.py.
The captures themselves are stored too, so the vectors do not depend on libm/SIMD details of the
machine that replays them.
"""
import io
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import refshim  # noqa: E402
from lddecode_b200 import synth  # noqa: E402

SPARSE = 37          # stride of the sparse sample of the stitched planes
CONFIGS = {
    # name: (system, fs_mhz, blocklen, capture samples, demod length, seed, bits)
    "ntsc": ("NTSC", 8 * 315 / 88, 16384, 700000, 640000, 0, 8),
    "pal": ("PAL", 35.46895, 16384, 1000000, 940000, 1, 8),
    "ntsc10": ("NTSC", 8 * 315 / 88, 32768, 700000, 640000, 0, 10),
}


def field_record(ref, f, system):
    rec = dict(valid=np.array(f.valid), peaklist=np.array(f.peaklist, dtype=np.int64),
               vsyncs=np.array(f.vsyncs, dtype=np.int64), nextfieldoffset=np.array(f.nextfieldoffset))
    if not f.valid:
        return rec
    rec.update(istop=np.array(int(f.istop)), linecount=np.array(f.linecount),
               med_hsync=np.array(f.med_hsync), hsync_tolerance=np.array(f.hsync_tolerance),
               linelocs1=np.array(f.linelocs1, dtype=np.float64), linebad=np.array(f.linebad, dtype=np.int8),
               linelocs2=np.array(f.linelocs2, dtype=np.float64), linelocs=np.array(f.linelocs, dtype=np.float64),
               dspicture=np.asarray(f.dspicture), tbcstart=np.array(f.tbcstart))
    codes = []
    for l in f.rf.SysParams["philips_codelines"]:
        lc = f.linecode[l]
        codes.append([-1] * 6 if lc is None else [int(x) for x in lc])
    rec["linecode"] = np.array(codes, dtype=np.int64)
    rec["framenr"] = np.array(-1 if f.vbi["framenr"] is None else int(f.vbi["framenr"]))
    if system == "NTSC":
        rec.update(linelocs3=np.array(f.linelocs3), linelocs4=np.array(f.linelocs4),
                   burstlevel=np.array(f.burstlevel, dtype=np.float32))
    if f.dsaudio is not None:
        rec["dsaudio"] = np.asarray(f.dsaudio)
        rec["audio_next_offset"] = np.array(f.audio_next_offset)
    return rec


def main():
    ref = refshim.load_reference()
    for name, (system, fs, N, ncap, dlen, seed, bits) in CONFIGS.items():
        cap = synth.SynthRF(system, fs, seed=seed, bits=bits).generate(ncap)
        out = dict(capture=cap, fs_mhz=np.array(fs), blocklen=np.array(N), demod_length=np.array(dlen),
                   sparse=np.array(SPARSE))
        ref.loader = refshim.make_array_loader(cap)
        rf = ref.RFDecode(inputfreq=fs, system=system, blocklen_=N)
        planes = list(rf.demodblock(cap[:N])[0].dtype.names)
        out["planes"] = np.array(planes)
        # --- block level, three MTF levels, two block positions
        for bi, (pos, mtf) in enumerate(((0, 0), (200000, 1), (403217, 0.73))):
            v, a = rf.demodblock(cap[pos:pos + N], mtf_level=mtf)
            out["blk%d_pos" % bi] = np.array(pos)
            out["blk%d_mtf" % bi] = np.array(mtf)
            for p in planes:
                out["blk%d_%s" % (bi, p)] = np.asarray(v[p], dtype=np.float64)
            out["blk%d_audio_left" % bi] = np.asarray(a["audio_left"])
            out["blk%d_audio_right" % bi] = np.asarray(a["audio_right"])
        # --- stitched demod as the Framer calls it (lddecode_core.py:1199, mtf_level=1)
        video, audio = rf.demod(refshim.MemFile(), 0, dlen, 1)
        out["demod_len"] = np.array(len(video))
        for p in planes:
            full = np.asarray(video[p], dtype=np.float64)
            out["demod_sparse_" + p] = full[::SPARSE].copy()
            out["demod_seam_" + p] = full[N - 1056 - 2048 + 0:N - 1056 + 2048].copy()   # around the first block seam
            out["demod_tail_" + p] = full[-4096:].copy()
        out["audio_left"] = np.asarray(audio["audio_left"])
        out["audio_right"] = np.asarray(audio["audio_right"])
        # a second window that does not start at 0 (exercises start > blockcut, lddecode_core.py:376-379)
        video2, audio2 = rf.demod(refshim.MemFile(), 54321, 300000, 0)
        out["demod2_len"] = np.array(len(video2))
        for p in planes:
            out["demod2_sparse_" + p] = np.asarray(video2[p], dtype=np.float64)[::SPARSE].copy()
        out["demod2_audio_left"] = np.asarray(audio2["audio_left"])
        # short read -> None (lddecode_core.py:386-392)
        assert rf.demod(refshim.MemFile(), ncap - 5000, 100000, 0) is None
        # --- field
        FC = ref.FieldNTSC if system == "NTSC" else ref.FieldPAL
        f = FC(rf, (video, audio), 0)
        assert f.valid, name
        for k, val in field_record(ref, f, system).items():
            out["field_" + k] = val
        if system == "NTSC":
            # intermediate of the burst pass: the resampled burst plane of the first pass (lddecode_core.py:1057)
            sb, _ = f.downscale(outwidth=f.outlinelen, lineinfo=f.linelocs2, channel="demod_burst", lineoffset=0)
            out["field_scaledburst_sparse"] = np.asarray(sb)[::7].copy()
        # scale() on a few lines, float64 (lddutils.py:83-97)
        ll = np.array(f.linelocs)
        sc = [ref.scale(video["demod"], ll[l], ll[l + 1], f.outlinelen) for l in (30, 31, 150)]
        out["scale_lines"] = np.array([30, 31, 150])
        out["scale_out"] = np.array(sc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "peaks", len(f.peaklist), "vsyncs", np.array(f.vsyncs).tolist(), "framenr", f.vbi["framenr"],
              "size", os.path.getsize(os.path.join(HERE, name + ".npz")))

    # --- one whole frame through Framer.readframe, as lddecode.py:88-98 does
    system, fs, N = "NTSC", 8 * 315 / 88, 16384
    cap = synth.SynthRF(system, fs, seed=3).generate(1530000)
    ref.loader = refshim.make_array_loader(cap)
    rf = ref.RFDecode(inputfreq=fs, system=system, blocklen_=N)
    fr = ref.Framer(rf)
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        combined, conaudio, nextsample, fields = fr.readframe(refshim.MemFile(), 0, True)
    np.savez_compressed(os.path.join(HERE, "ntsc_frame.npz"), capture=cap, fs_mhz=np.array(fs), blocklen=np.array(N),
                        combined=combined, conaudio=conaudio, nextsample=np.array(nextsample),
                        framenr=np.array(fr.vbi["framenr"]), mtf_level=np.array(fr.mtf_level),
                        audio_offset=np.array(fr.audio_offset),
                        field_readlens=np.array([f.linecount for f in fields]))
    print("frame: framenr", fr.vbi["framenr"], "nextsample", nextsample, "audio", None if conaudio is None else len(conaudio))

    # --- integer unpackers
    rng = np.random.default_rng(42)
    s10 = rng.integers(0, 1024, 3 * 4099, dtype=np.uint16)
    s10[:8] = [0, 1023, 512, 511, 1, 1022, 513, 0]
    words = synth.pack_r30(s10)
    lds = synth.pack_lds(rng.integers(0, 1024, 4 * 3001, dtype=np.uint16))
    out = dict(r30_words=words, lds_bytes=lds)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    exe = os.path.join(ROOT, "oracle", "_ref", "ddunpack")
    res = subprocess.run([exe], input=words.tobytes(), stdout=subprocess.PIPE, check=True).stdout
    out["r30_ddunpack_i16"] = np.frombuffer(res, dtype="<i2").copy()
    # ddpack round trip of the sign-extended samples gives the words back (ddpack.c:11-29)
    back = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ddpack")], input=res, stdout=subprocess.PIPE, check=True).stdout
    assert np.array_equal(np.frombuffer(back, dtype="<u4"), words)
    # the Python loaders (lddutils.py:150-229); load_packed_data_3_32 calls the removed binary
    # np.fromstring, so it is run with np.fromstring aliased to np.frombuffer for this call only
    import io
    old = getattr(np, "fromstring")
    np.fromstring = lambda buf, dtype, count=-1: np.frombuffer(buf, dtype=dtype, count=count)
    try:
        r30 = [np.asarray(ref.load_packed_data_3_32(io.BytesIO(words.tobytes()), s, 1000)) for s in (0, 1, 2, 301)]
    finally:
        np.fromstring = old
    out["r30_py_starts"] = np.array([0, 1, 2, 301])
    out["r30_py"] = np.array(r30)
    l40 = [np.asarray(ref.load_packed_data_4_40(io.BytesIO(lds.tobytes()), s, 1000)) for s in (0, 1, 2, 3, 402)]
    out["lds_py_starts"] = np.array([0, 1, 2, 3, 402])
    out["lds_py"] = np.array(l40)
    np.savez_compressed(os.path.join(HERE, "unpack.npz"), **out)
    print("unpack ok")


if __name__ == "__main__":
    main()

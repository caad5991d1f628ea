"""Fuzzing of the whole path on the CPU emulation of the kernel sources (tests/emu) against the oracle's Framer walk.

Every case draws a system, sample rate (8fsc / 40 MSPS), capture format (u8, u16, packed .lds / .r30), disc speed error,
noise level, lead-in, CLV / CAV codes, number of fields and number of read-position ranges, decodes the capture with
CaptureDecoder in the default lane (both audio channels on) and compares read positions, peak counts, VBI codes, TBC
pictures (+-1 LSB), PCM sample counts / dropped fields / carried offset (exact) and PCM values (interpolation bound)
with oracle.framer_audio_walk.  Harsher campaigns of round 2 (speed errors to 3 %, noise to 15 % of the carrier, dropouts,
holes, noise-only lead-ins, clipped video, captures starting inside a vertical interval) used variants of this script;
what they found is in DESIGN.md section 10 and pinned by tests/test_parity_gpu.py::
test_unreadable_and_noisy_captures_follow_the_reference and tests/test_field.py (long / abnormal line spans).

    python tools/fuzz_emu.py FIRST_SEED LAST_SEED        # ~10-30 s per case on one core
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lddecode_b200 import _lib, pipeline, rfdecode, synth, parallel
from oracle import ldd_oracle as O
from emu_util import emu_backend

def case(seed):
    rng = np.random.default_rng(seed)
    system = "NTSC" if rng.random() < 0.6 else "PAL"
    fs = 8 * 315 / 88 if system == "NTSC" else 35.46895
    if rng.random() < 0.25: fs = 40.0
    speed = 1.0 + rng.uniform(-0.003, 0.003)
    noise = float(rng.choice([0.5, 2.0, 5.0]))
    lead = int(rng.integers(12, 200))
    fmtname = str(rng.choice(["u8", "u16", "lds", "r30"]))
    nfields = int(rng.integers(2, 4))
    ncap = int(fs * 1e6 / (60 if system == "NTSC" else 50) * (nfields + 1.3)) // 12 * 12
    bits = 8 if fmtname == "u8" else 10
    s = synth.SynthRF(system, fs * speed, seed=seed, bits=bits, noise=noise, lead_lines=lead, clv=bool(rng.random() < 0.5)).generate(ncap)
    raw, fmt = {"u8": (s, _lib.FMT_U8), "u16": (s, _lib.FMT_U16), "lds": (synth.pack_lds(s), _lib.FMT_LDS40), "r30": (synth.pack_r30(s), _lib.FMT_R30)}[fmtname]
    be = emu_backend()
    rf = rfdecode.RFDecode(fs, system, 16384, _backend=be)
    cd = pipeline.CaptureDecoder(rf)
    nr = int(rng.integers(1, 4))
    got, gpcm, off, st = [], [], 0.0, 2
    for r0, r1 in parallel.shard_bounds(ncap, nr):
        if nr == 1:
            res = cd.decode(be.to_device(raw), fmt, ncap)
        else:
            lo, hi = parallel.needed_window(cd, ncap, r0, r1)
            lo = lo // 12 * 12; hi = min(ncap, -(-hi // 12) * 12)
            grp, bpg = {_lib.FMT_U8: (1, 1), _lib.FMT_U16: (1, 1), _lib.FMT_LDS40: (4, 5), _lib.FMT_R30: (3, 1)}[fmt]
            sub = raw[lo // grp * bpg: -(-hi // grp) * bpg]
            res = cd.decode_range(be.to_device(sub), fmt, lo, hi - lo, ncap, r0, r1)
        pics = cd.pictures(res)
        pcm, off, st = cd.pcm(res, off, st)
        for k, j in enumerate(res.located):
            got.append((int(res.readsamples[j]), pics[k][2], pcm[k], int(res.infos[j].npeaks), res.vbi_codes()[k].tolist()))
    dec = O.Decoder(fs, system, 16384, analog_audio=True)
    ld = lambda a, n: s[a:a + n] if a + n <= ncap else None
    ref, roff = O.framer_audio_walk(dec, ld, 99)
    msgs = []
    if len(ref) != len(got): msgs.append("fields %d vs oracle %d" % (len(got), len(ref)))
    if off != roff: msgs.append("audio offset %r vs %r" % (off, roff))
    for k, ((rs, pic, pcm, npk, codes), (ors, f, opcm)) in enumerate(zip(got, ref)):
        if rs != ors: msgs.append("f%d readsample %d vs %d" % (k, rs, ors)); break
        if npk != len(f.peaklist): msgs.append("f%d npeaks" % k)
        if pic is None: msgs.append("f%d flagged" % k); continue
        dd = np.abs(pic.astype(np.int64) - f.dspicture.astype(np.int64))
        if dd.max() > 1: msgs.append("f%d TBC max %d (%d > 1)" % (k, dd.max(), int((dd > 1).sum())))
        want = [O.decode_philips(dec, None, None, None) if False else f.linecode[l] for l in dec.SP["philips_codelines"]]
        from lddecode_b200 import field as F
        if [F.code_nibbles(c) for c in codes[:3]] != want: msgs.append("f%d VBI" % k)
        if (pcm is None) != (opcm is None): msgs.append("f%d pcm dropped mismatch" % k)
        elif pcm is not None:
            if len(pcm) != len(opcm): msgs.append("f%d pcm len" % k)
            else:
                d = np.abs(pcm.astype(int) - opcm.astype(int))
                if d.max() > 60 or d.mean() > 3: msgs.append("f%d pcm max %d mean %.2f" % (k, d.max(), d.mean()))
    return "seed %d %s fs %.3f speed %.4f noise %.1f lead %d %s clv? nr %d fields %d: %s" % (seed, system, fs, speed, noise, lead, fmtname, nr, len(got), "OK" if not msgs else "; ".join(msgs))

if __name__ == "__main__":
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        t = time.time()
        try:
            print(case(seed), "(%.0f s)" % (time.time() - t), flush=True)
        except Exception as e:
            import traceback; traceback.print_exc()
            print("seed", seed, "EXC", type(e).__name__, str(e)[:200], flush=True)

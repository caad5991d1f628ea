"""Seeded synthetic LaserDisc RF (FM-modulated NTSC / PAL composite + two analog FM audio carriers).

Input generator for tests and bench.py (the reference ships no sample captures, SURVEY.md
section 4).  The recipe follows SURVEY.md section 8d, which in turn follows the reference's own
test-signal idea (attic2/ld-decoder-dev11.ipynb cells 21-22 and the 'Femp' filter kept for
"test signal generation" at lddecode_core.py:190-192): build baseband video in Hz on the sample
grid from the level constants of lddecode_core.py:30-84, band-limit, pre-emphasise with the
inverse of the decoder's de-emphasis, FM-modulate, add the audio carriers and noise, quantise.

Host-side numpy only; this is not part of the timed path.
"""
import numpy as np
import scipy.signal as sps

NTSC = dict(
    name="NTSC", fsc=315.0e6 / 88.0, line_period=227.5 / (315.0e6 / 88.0), frame_lines=525,
    ire0=8100000.0, hz_ire=1700000.0 / 140.0, sync_ire=-40.0,
    deemp=(120 * .32, 320 * .32),
    audio_lfreq=(1000000 * 315 / 88 / 227.5) * 146.25, audio_rfreq=(1000000 * 315 / 88 / 227.5) * 178.75,
    eq_us=2.3, n_eq=6, burst_cycles=9, burst_ire=20.0, lpf_hz=4.2e6,
    active=((20, 261), (283, 524)), codelines=((14, 19), (277, 282)),
    front_us=9.4, back_us=1.5,
)
PAL = dict(
    name="PAL", fsc=4433618.75, line_period=64e-6, frame_lines=625,
    ire0=7100000.0, hz_ire=8000.0, sync_ire=-.3 * (100 / .7),
    deemp=(100 * .4, 400 * .4),
    audio_lfreq=(1000000 / 64) * 43.75, audio_rfreq=(1000000 / 64) * 68.25,
    eq_us=2.35, n_eq=5, burst_cycles=10, burst_ire=21.4, lpf_hz=4.8e6,
    active=((23, 310), (336, 622)), codelines=((17, 22), (330, 335)),
    front_us=10.5, back_us=1.65,
)
SYSTEMS = {"NTSC": NTSC, "PAL": PAL}

T_LINESTART, T_LINEMID, T_EQ, T_BROAD = 0, 1, 2, 3


def _slot_table(sp):
    """Half-line slot types for one frame (2 fields)."""
    nhl = sp["frame_lines"] * 2
    k = sp["n_eq"]
    tab = np.where(np.arange(nhl) % 2 == 0, T_LINESTART, T_LINEMID).astype(np.int8)
    for base in (0, nhl // 2):
        tab[base:base + k] = T_EQ
        tab[base + k:base + 2 * k] = T_BROAD
        tab[base + 2 * k:base + 3 * k] = T_EQ
        # an odd slot straight after the last EQ pulse is a blank half line
        nxt = base + 3 * k
        if nxt % 2 == 1:
            tab[nxt] = T_LINEMID
    return tab


def _line_kind(sp):
    """Per frame line (hsync grid): 0 blank, 1 active picture, 2 Philips code line."""
    nl = sp["frame_lines"] + 1
    kind = np.zeros(nl, dtype=np.int8)
    for lo, hi in sp["active"]:
        kind[lo:hi + 1] = 1
    for lo, hi in sp["codelines"]:
        kind[lo:hi + 1] = 2
    return kind


def _bcd(n, digits):
    out = 0
    for d in range(digits):
        out |= (n % 10) << (4 * d)
        n //= 10
    return out


class SynthRF:
    """Streaming generator: call generate(n) repeatedly; output continues where the last call ended.

    system   'NTSC' | 'PAL'
    fs_mhz   sample rate in MHz (8fsc: NTSC 8*315/88 = 28.636..., PAL 35.46895)
    bits     8 -> uint8 (amplitude 100, +128); 10 -> uint16 0..1023 (amplitude 400, +512)
    lead_lines  number of lines before the first vertical interval (the reference crashes when a
             vsync lies within the first 11 peaks, lddecode_core.py:545-546, 604)
    """

    CHUNK = 1 << 20

    def __init__(self, system="NTSC", fs_mhz=None, seed=0, bits=8, audio=True, clv=False,
                 frame0=1, lead_lines=20, noise=0.5, random_luma=True):
        self.sp = sp = SYSTEMS[system]
        self.system = system
        if fs_mhz is None:
            fs_mhz = 8 * 315 / 88 if system == "NTSC" else 35.46895
        self.fs = fs_mhz * 1e6
        self.bits = bits
        self.audio = audio
        self.clv = clv
        self.frame0 = frame0
        self.noise = noise
        self.seed = seed
        self.random_luma = random_luma
        self.slots = _slot_table(sp)
        self.kind = _line_kind(sp)
        self.nhl = len(self.slots)
        self.halfline = sp["line_period"] / 2
        # capture starts lead_lines before the start of a frame's first vertical interval
        self.hl0 = self.nhl - 2 * lead_lines
        self.n_done = 0
        rng = np.random.default_rng([seed, 12345])
        self.line_f = rng.uniform(0.5, 6.0, 4096)
        self.line_ph = rng.uniform(0, 2 * np.pi, 4096)
        self.line_amp = rng.uniform(0.0, 12.0, 4096) if random_luma else np.zeros(4096)
        self.line_chroma_ph = rng.uniform(0, 2 * np.pi, 4096)
        # band-limit + pre-emphasis (inverse of the decoder's de-emphasis, lddecode_core.py:186-192)
        fs_half = self.fs / 2
        self.lpf = sps.butter(4, sp["lpf_hz"] / fs_half, "low")
        d0, d1 = sp["deemp"]
        tf_b, tf_a = sps.zpk2tf(-d0 * (10 ** -10), -d1 * (10 ** -10), d1 / d0)
        self.emp = sps.bilinear(tf_b, tf_a, 1.0 / fs_half)
        self.zi_lpf = None
        self.zi_emp = None
        self.ph_video = 0.0
        self.ph_l = 0.0
        self.ph_r = 0.0

    # -- baseband ------------------------------------------------------------------------
    def _baseband_ire(self, n0, n):
        sp = self.sp
        t = (np.arange(n0, n0 + n, dtype=np.float64)) / self.fs
        hl = t / self.halfline + self.hl0
        ihl = np.floor(hl)
        th = (hl - ihl) * self.halfline * 1e6          # microseconds inside the half-line slot
        frame = np.floor(ihl / self.nhl).astype(np.int64)
        pos = (ihl - frame * self.nhl).astype(np.int64)
        typ = self.slots[pos]
        H_us = sp["line_period"] * 1e6
        half_us = H_us / 2
        th0, tr = 0.3, 0.14

        def pulse(x, start, width):
            return np.clip((x - start) / tr + 0.5, 0, 1) - np.clip((x - start - width) / tr + 0.5, 0, 1)

        width = np.full(n, 4.7)
        width[typ == T_EQ] = sp["eq_us"]
        width[typ == T_BROAD] = half_us - 4.7
        width[typ == T_LINEMID] = 0.0
        sync = pulse(th, th0, width)
        ire = sp["sync_ire"] * sync

        # time since the start of the line, for slots on the hsync grid
        line = pos // 2
        tl = np.where(typ == T_LINEMID, th + half_us, th)
        online = (typ == T_LINESTART) | (typ == T_LINEMID)
        kind = np.where(online, self.kind[line], 0)
        absline = (frame * sp["frame_lines"] + line) & 4095
        sc_phase = 2 * np.pi * sp["fsc"] * t

        # colour burst on every line that has a normal hsync (not in the EQ/broad area)
        b0 = th0 + 4.7 + 0.6
        blen = sp["burst_cycles"] / (sp["fsc"] * 1e-6)
        burst_on = online & (tl >= b0) & (tl < b0 + blen) & (self.slots[(line * 2) % self.nhl] == T_LINESTART)
        ire = ire + np.where(burst_on, sp["burst_ire"] * np.sin(sc_phase + np.pi), 0.0)

        # active picture: ramp + per-line sinusoid + chroma
        a0 = th0 + sp["front_us"]
        a1 = H_us - sp["back_us"]
        act = (kind == 1) & (tl >= a0) & (tl < a1)
        u = (tl - a0) / (a1 - a0)
        luma = 10 + 80 * u + self.line_amp[absline] * np.sin(2 * np.pi * self.line_f[absline] * u + self.line_ph[absline])
        chroma = 15 * np.sin(sc_phase + self.line_chroma_ph[absline])
        ire = ire + np.where(act, luma + chroma, 0.0)

        # Philips code lines: 24 biphase cells of 2 us (decoder: lddecode_core.py:814-834)
        c0 = th0 + 4.7 + 5.5
        cell = np.floor((tl - c0) / 2.0)
        code_on = (kind == 2) & (cell >= 0) & (cell < 24)
        if code_on.any():
            fr = frame + self.frame0
            if self.clv:
                # CLV minutes code 0xF?DD?? (lddecode_core.py:850-854): fixed hour 0, minute = frame//1800
                code = np.int64(0xF0DD00) | np.array([_bcd(int(m), 2) for m in ((fr // 1800) % 60)], dtype=np.int64)
            else:
                code = np.int64(0xF00000) | np.array([_bcd(int(f), 5) for f in (fr % 80000)], dtype=np.int64)
            ci = np.clip(cell, 0, 23).astype(np.int64)
            bit = (code >> (23 - ci)) & 1
            second_half = ((tl - c0) - 2.0 * cell) >= 1.0
            high = np.where(bit == 1, second_half, ~second_half)
            ire = ire + np.where(code_on & high, 100.0, 0.0)

        hz = sp["ire0"] + sp["hz_ire"] * ire
        if self.system == "PAL":
            # 3.75 MHz pilot inside the sync tip of normal lines (decoder: lddecode_core.py:974-988)
            tip = (typ == T_LINESTART) & (th >= th0 + 0.3) & (th < th0 + 4.4)
            hz = hz + np.where(tip, 200000.0 * np.sin(2 * np.pi * 3.75e6 * t), 0.0)
        return hz, t

    # -- RF ------------------------------------------------------------------------------
    def _chunk(self, n):
        sp = self.sp
        n0 = self.n_done
        hz, t = self._baseband_ire(n0, n)
        if self.zi_lpf is None:
            self.zi_lpf = sps.lfilter_zi(*self.lpf) * hz[0]
            self.zi_emp = sps.lfilter_zi(*self.emp) * hz[0]
        hz, self.zi_lpf = sps.lfilter(self.lpf[0], self.lpf[1], hz, zi=self.zi_lpf)
        hz, self.zi_emp = sps.lfilter(self.emp[0], self.emp[1], hz, zi=self.zi_emp)
        amp = 100.0 if self.bits == 8 else 400.0
        ph = self.ph_video + 2 * np.pi * np.cumsum(hz) / self.fs
        self.ph_video = float(ph[-1] % (2 * np.pi))
        rf = amp * np.sin(ph)
        if self.audio:
            al = sp["audio_lfreq"] + 50000.0 * np.sin(2 * np.pi * 1000.0 * t)
            ar = sp["audio_rfreq"] + 50000.0 * np.sin(2 * np.pi * 400.0 * t)
            pl = self.ph_l + 2 * np.pi * np.cumsum(al) / self.fs
            pr = self.ph_r + 2 * np.pi * np.cumsum(ar) / self.fs
            self.ph_l = float(pl[-1] % (2 * np.pi))
            self.ph_r = float(pr[-1] % (2 * np.pi))
            rf += (amp / 10.0) * (np.sin(pl) + np.sin(pr))
        if self.noise > 0:
            rng = np.random.default_rng([self.seed, 777, n0 // self.CHUNK])
            rf += rng.normal(0.0, self.noise * (amp / 100.0), n)
        self.n_done += n
        if self.bits == 8:
            return np.clip(np.round(rf) + 128, 0, 255).astype(np.uint8)
        return np.clip(np.round(rf) + 512, 0, 1023).astype(np.uint16)

    def generate(self, n):
        """Next n samples (uint8 for bits=8, uint16 in 0..1023 for bits=10)."""
        out = np.empty(n, dtype=np.uint8 if self.bits == 8 else np.uint16)
        done = 0
        while done < n:
            # keep chunk boundaries on multiples of CHUNK so the noise is a function of position only
            room = self.CHUNK - (self.n_done % self.CHUNK)
            m = min(room, n - done)
            out[done:done + m] = self._chunk(m)
            done += m
        return out


def synth_capture(system="NTSC", nsamples=1200000, **kw):
    return SynthRF(system, **kw).generate(nsamples)


# -- 10-bit packers (test-vector generators; inverse of the loaders in lddutils.py:150-229) -----
def pack_r30(samples10):
    """3 x 10 bit in a little-endian u32, ddpack.c:22-29 layout, from raw 0..1023 values."""
    s = np.asarray(samples10, dtype=np.uint32)
    pad = (-len(s)) % 3
    if pad:
        s = np.concatenate([s, np.zeros(pad, dtype=np.uint32)])
    s = s.reshape(-1, 3)
    return (s[:, 0] | (s[:, 1] << 10) | (s[:, 2] << 20)).astype("<u4")


def pack_lds(samples10):
    """4 x 10 bit in 5 bytes, MSB first (Domesday Duplicator .lds; lddutils.py:176-229)."""
    s = np.asarray(samples10, dtype=np.uint16)
    pad = (-len(s)) % 4
    if pad:
        s = np.concatenate([s, np.zeros(pad, dtype=np.uint16)])
    s = s.reshape(-1, 4).astype(np.uint32)
    out = np.empty((len(s), 5), dtype=np.uint8)
    out[:, 0] = s[:, 0] >> 2
    out[:, 1] = ((s[:, 0] & 3) << 6) | (s[:, 1] >> 4)
    out[:, 2] = ((s[:, 1] & 0xF) << 4) | (s[:, 2] >> 6)
    out[:, 3] = ((s[:, 2] & 0x3F) << 2) | (s[:, 3] >> 8)
    out[:, 4] = s[:, 3] & 0xFF
    return out.reshape(-1)

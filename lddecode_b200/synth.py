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
                 frame0=1, lead_lines=20, noise=0.5, random_luma=True, tile_frames=0):
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
        self.tile_frames = tile_frames        # > 0: picture content and codes repeat every tile_frames frames (TiledCapture)
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
        if self.tile_frames:
            frame = frame % self.tile_frames
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


class TiledCapture:
    """An arbitrarily long NTSC CLV capture that is never materialised: a two-frame
    baseband template -- the four-field colour sequence, 2 * 525 * 1820 = 1 911 000 samples at 8fsc, so the tiling is
    period-exact -- FM-modulated ON THE DEVICE with a running carrier phase (no phase jump at tile seams), plus the two
    analog audio carriers (closed-form phase) and seeded noise.  Sample n of the capture is a pure function of n, so
    any window can be generated anywhere: every rank generates its own shard in HBM, and the parity tests regenerate
    the windows they check on the host side of the same function.

    Test / benchmark tooling built on torch ops; not part of the decoder and never timed."""

    CHUNK = 1 << 24                                   # samples generated per pass (bounds the temporaries)

    def __init__(self, seed=2, device="cuda", noise=0.5):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.seed, self.noise = seed, noise
        sp = NTSC
        self.fs = 8 * 315e6 / 88
        self.P = 2 * 525 * 1820
        gen = SynthRF("NTSC", self.fs / 1e6, seed=seed, clv=True, audio=False, noise=0.0, tile_frames=2)
        # three periods of filtered, pre-emphasised baseband; the middle one is the periodic steady state
        hz, _ = gen._baseband_ire(0, 3 * self.P)
        hz = sps.lfilter(gen.lpf[0], gen.lpf[1], hz)
        hz = sps.lfilter(gen.emp[0], gen.emp[1], hz)
        hz = hz[self.P:2 * self.P]
        # phase advance up to and including sample t of the period, in cycles (float64; < 6e5 cycles per period)
        cyc = np.cumsum(hz / self.fs)
        self.cycles_per_period = float(cyc[-1])
        self.tmpl = torch.from_numpy(cyc).to(self.device)
        self.sp = sp

    def generate(self, n0, n, out=None):
        """Samples [n0, n0 + n) as a uint8 tensor on the device."""
        torch = self.torch
        if out is None:
            out = torch.empty(n, dtype=torch.uint8, device=self.device)
        done = 0
        while done < n:
            a = n0 + done
            m = min(self.CHUNK, n - done)
            out[done:done + m] = self._chunk(a, m)
            done += m
        return out

    @staticmethod
    def _mix64(torch, z):
        """splitmix64 finaliser on int64 tensors (wrapping arithmetic, logical shifts by masking)."""
        def s64(c):
            return c - (1 << 64) if c >= (1 << 63) else c
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * s64(0x94D049BB133111EB)
        return z ^ ((z >> 31) & ((1 << 33) - 1))

    def _gauss(self, idx):
        """N(0, 1) per sample as a pure function of (seed, sample index): counter-based hash + Box-Muller."""
        torch = self.torch
        k = (self.seed * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
        k = k - (1 << 64) if k >= (1 << 63) else k              # as a signed 64-bit value (tensor arithmetic wraps)
        a = self._mix64(torch, idx * 2 + 1 + k)
        b = self._mix64(torch, idx * 2 + 2 + k)
        u1 = (((a >> 11) & ((1 << 53) - 1)).to(torch.float64) + 0.5) / float(1 << 53)
        u2 = ((b >> 11) & ((1 << 53) - 1)).to(torch.float64) / float(1 << 53)
        return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2 * np.pi * u2)

    def _chunk(self, a, m):
        torch = self.torch
        dev = self.device
        idx = torch.arange(a, a + m, dtype=torch.int64, device=dev)
        tile = idx // self.P
        t = idx - tile * self.P
        # carrier phase in cycles, reduced before the sine: tile * frac(cycles per period) + template
        fracp = self.cycles_per_period - np.floor(self.cycles_per_period)
        cyc = self.tmpl[t] + torch.frac(tile.to(torch.float64) * fracp)
        rf = 100.0 * torch.sin(2 * np.pi * torch.frac(cyc))
        # audio carriers: f / fs = 146.25 / 1820 and 178.75 / 1820 cycles per sample (exactly periodic over 4 lines = 7280
        # samples); FM by 1 kHz / 400 Hz tones with 50 kHz deviation: phase = carrier - (dev / f_tone) cos(2 pi f_tone t)
        k4 = (idx % 7280).to(torch.float64)
        # tone phases: f_tone / fs = 11 / 315000 (1 kHz) and 11 / 787500 (400 Hz) cycles per sample
        p1 = ((idx * 11) % 315000).to(torch.float64) / 315000.0
        p2 = ((idx * 11) % 787500).to(torch.float64) / 787500.0
        pl = 146.25 / 1820.0 * k4 - (50000.0 / 1000.0) / (2 * np.pi) * torch.cos(2 * np.pi * p1)
        pr = 178.75 / 1820.0 * k4 - (50000.0 / 400.0) / (2 * np.pi) * torch.cos(2 * np.pi * p2)
        rf = rf + 10.0 * (torch.sin(2 * np.pi * torch.frac(pl)) + torch.sin(2 * np.pi * torch.frac(pr)))
        if self.noise > 0:
            rf = rf + self.noise * self._gauss(idx)
        return torch.clamp(torch.round(rf) + 128.0, 0, 255).to(torch.uint8)


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

"""CPU ORACLE -- test infrastructure, NOT product code.

A float64 numpy/scipy restatement of the reference's RF demodulation + sync + TBC hot path
(wondras/ld-decode, lddecode_core.py / lddutils.py), used ONLY by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the checker
and the CPU baseline.  The product package (lddecode_b200/) never imports this module.

Parity pin: the reference ships no tests or golden vectors for this path,
so the pin is the reference ITSELF run in the build container on seeded synthetic RF:
tests/golden/make_golden.py imports /root/reference (unmodified, import shims only), runs it,
and stores its outputs under tests/golden/*.npz; tests/test_oracle_golden.py checks every
function below against those vectors.  Third-party arithmetic the reference relies on and
that is not under /root/reference: numpy.fft (pocketfft), numpy.angle/unwrap,
scipy.signal.{butter,freqz,zpk2tf,bilinear,firwin}, scipy.interpolate.{splrep,splev}
(FITPACK); no versions are pinned by the reference, the installed numpy 2.3 / scipy 1.18 are
the de-facto oracle and this file calls the same routines.

Each function cites the reference lines it restates.
"""
import numpy as np
import scipy.signal as sps
from scipy import interpolate

TAU = 2 * np.pi

# ---- system / decoder parameters (lddecode_core.py:30-117) -------------------------------------
_NTSC_FSC = 315.0 / 88.0


def sysparams(system):
    if system == "NTSC":
        lp = 1 / (_NTSC_FSC / 227.5)
        sp = dict(fsc_mhz=_NTSC_FSC, pilot_mhz=_NTSC_FSC, frame_lines=525, ire0=8100000,
                  hz_ire=1700000 / 140.0, vsync_ire=-40, analog_audio=True,
                  audio_lfreq=(1000000 * 315 / 88 / 227.5) * 146.25,
                  audio_rfreq=(1000000 * 315 / 88 / 227.5) * 178.75,
                  philips_codelines=[16, 17, 18], topfirst=True, line_period=lp)
        sp["FPS"] = 1000000 / (525 * lp)
        sp["outlinelen"] = int(np.round(lp * sp["fsc_mhz"] * 4))
        dp = dict(audio_notchwidth=350000, audio_notchorder=2, video_deemp=(120 * .32, 320 * .32),
                  video_bpf=[3500000, 13200000], video_bpf_order=3,
                  video_lpf_freq=4200000, video_lpf_order=5)
    elif system == "PAL":
        sp = dict(FPS=25, fsc_mhz=((1 / 64) * 283.75) + (25 / 1000000), pilot_mhz=3.75,
                  frame_lines=625, line_period=64, ire0=7100000, hz_ire=800000 / 100.0,
                  analog_audio=True, audio_lfreq=(1000000 / 64) * 43.75,
                  audio_rfreq=(1000000 / 64) * 68.25, philips_codelines=[19, 20, 21],
                  topfirst=False)
        sp["outlinelen"] = int(np.round(64 * sp["fsc_mhz"] * 4))
        sp["vsync_ire"] = -.3 * (100 / .7)
        dp = dict(audio_notchwidth=200000, audio_notchorder=2, video_deemp=(100 * .4, 400 * .4),
                  video_bpf=(2500000, 14500000), video_bpf_order=3,
                  video_lpf_freq=5200000, video_lpf_order=9)
    else:
        raise ValueError(system)
    return sp, dp


def _freqresp(ba, n):
    """lddutils.py:256-257 (filtfft)."""
    return sps.freqz(ba[0], ba[1], n, whole=1)[1]


def _hilbert_taps(terms=128):
    """lddutils.py:246-249."""
    return np.fft.fftshift(np.fft.ifft([0] + [1] * terms + [0] * terms))


class Decoder:
    """State of one RFDecode instance (lddecode_core.py:119-145): parameters and filter tables."""

    def __init__(self, inputfreq=40, system="NTSC", blocklen=16384, analog_audio=True):
        self.system = system
        self.N = self.blocklen = blocklen
        self.blockcut = 1024
        self.freq = inputfreq
        self.freq_hz = inputfreq * 1000000
        self.nyq_hz = inputfreq * 1000000 / 2
        self.nyq_mhz = inputfreq / 2
        self.SP, self.DP = sysparams(system)
        self.linelen = int(np.round(self.freq_hz / (1000000.0 / self.SP["line_period"])))
        self.analog_audio = analog_audio
        self.F = {}
        self._video_tables()
        if analog_audio:
            self._audio_tables()
        self.blockcut_end = self.F["F05_offset"]

    def iretohz(self, ire):
        return self.SP["ire0"] + self.SP["hz_ire"] * ire

    # lddecode_core.py:152-214
    def _video_tables(self):
        N, F, SP, DP = self.N, self.F, self.SP, self.DP
        pole_angles = (12.5, 27.5) if self.system == "NTSC" else (10, 28)
        poles = [.7 * np.exp(1j * np.pi * a / 20) for a in pole_angles]
        F["MTF"] = _freqresp(sps.zpk2tf([], poles, 1.11), N)
        F["hilbert"] = np.fft.fft(_hilbert_taps(), N)
        bpf = sps.butter(DP["video_bpf_order"], [f / self.nyq_hz for f in DP["video_bpf"]], btype="bandpass")
        rfv = _freqresp(bpf, N)
        if SP["analog_audio"]:
            for key in ("audio_lfreq", "audio_rfreq"):
                w = DP["audio_notchwidth"]
                notch = sps.butter(DP["audio_notchorder"],
                                   [(SP[key] - w) / self.nyq_hz, (SP[key] + w) / self.nyq_hz], btype="bandstop")
                rfv = rfv * _freqresp(notch, N)
        F["RFVideo"] = rfv * F["hilbert"]
        lpf = _freqresp(sps.butter(DP["video_lpf_order"], DP["video_lpf_freq"] / self.nyq_hz, "low"), N)
        d0, d1 = DP["video_deemp"]
        b, a = sps.zpk2tf(-d1 * (10 ** -10), -d0 * (10 ** -10), d0 / d1)
        deemp = _freqresp(sps.bilinear(b, a, 1.0 / self.nyq_hz), N)
        F["FVideo"] = lpf * deemp
        F["F05_offset"] = 32
        f05 = _freqresp((sps.firwin(65, [0.5 / self.nyq_mhz], pass_zero=True), [1.0]), N)
        F["FVideo05"] = lpf * deemp * f05
        fsc = SP["fsc_mhz"]
        burst = _freqresp(sps.butter(1, [(fsc - .1) / self.nyq_mhz, (fsc + .1) / self.nyq_mhz], btype="bandpass"), N)
        F["FVideoBurst"] = lpf * deemp * burst
        if self.system == "PAL":
            pilot = _freqresp(sps.butter(1, [3.7 / self.nyq_mhz, 3.8 / self.nyq_mhz], btype="bandpass"), N)
            F["FVideoPilot"] = lpf * deemp * pilot
        F["FPsync_ba"] = sps.butter(1, 0.05 / self.nyq_mhz, btype="low")
        F["FPsync"] = _freqresp(F["FPsync_ba"], N)

    # lddecode_core.py:223-279
    def _audio_tables(self):
        N, F, SP = self.N, self.F, self.SP
        fdiv1 = 32 if self.freq >= 32 else 16
        half = N // (fdiv1 * 2)
        F["freq_arf"] = self.freq_hz / (fdiv1 / 2)
        F["audio_fdiv1"] = fdiv1
        cfreq = (SP["audio_rfreq"] + SP["audio_lfreq"]) // 2
        center = int((cfreq / self.freq_hz) * N)
        lo, hi = int(center - half), int(center + half)
        F["slice_lo"] = slice(lo, hi)
        F["slice_hi"] = slice(N - hi, N - lo)
        F["audio_lowfreq"] = cfreq - (self.freq_hz / (2 * fdiv1))
        for key, out in (("audio_lfreq", "audio_lfilt"), ("audio_rfreq", "audio_rfilt")):
            taps = sps.firwin(800, [(SP[key] - 150000) / self.nyq_hz, (SP[key] + 150000) / self.nyq_hz], pass_zero=False)
            F[out] = self._slice1(_freqresp([taps, 1.0], N) * F["hilbert"])
        F["audio_fdiv2"] = 4
        F["freq_aud2"] = F["freq_arf"] / 4
        F["slice2_lo"] = slice(0, N // 8)
        F["slice2_hi"] = slice(N - N // 8, N)
        F["audio_lpf2"] = _freqresp([sps.firwin(65, [21000 / (F["freq_aud2"] / 2)]), [1.0]], N // 4)

    def _slice1(self, spec):
        return np.concatenate([spec[self.F["slice_lo"]], spec[self.F["slice_hi"]]])

    def _slice2(self, spec):
        return np.concatenate([spec[self.F["slice2_lo"]], spec[self.F["slice2_hi"]]])


# ---- FM discriminator (lddutils.py:320-334) ------------------------------------------------------
def fm_discriminate(analytic, freq_hz):
    ang = np.angle(analytic)
    d = np.concatenate([[0.0], np.diff(ang)])
    if d[0] < -np.pi:
        d[0] += TAU
    d = np.unwrap(d)
    while np.min(d) < 0:
        d[d < 0] += TAU
    while np.max(d) > TAU:
        d[d > TAU] -= TAU
    return d * (freq_hz / TAU)


# ---- block demodulation (lddecode_core.py:288-330) -----------------------------------------------
def demodblock(dec, data, mtf_level=0):
    """Returns (video planes dict, audio dict | None), all float64 length N (audio N/(fdiv1/2))."""
    F, N = dec.F, dec.N
    X = np.fft.fft(data[:N])
    Y = X * F["RFVideo"]
    if mtf_level != 0:
        Y = Y * F["MTF"] ** mtf_level
    demod = fm_discriminate(np.fft.ifft(Y), dec.freq_hz)
    D = np.fft.fft(demod)
    v = {}
    v["demod"] = np.fft.ifft(D * F["FVideo"]).real
    v["demod_05"] = np.roll(np.fft.ifft(D * F["FVideo05"]).real, -F["F05_offset"])
    v["demod_burst"] = np.fft.ifft(D * F["FVideoBurst"]).real
    insync = (v["demod_05"] >= dec.iretohz(-55)) & (v["demod_05"] <= dec.iretohz(-25))
    v["demod_sync"] = np.fft.ifft(np.fft.fft(insync) * F["FPsync"]).real
    if dec.system == "PAL":
        v["demod_pilot"] = np.fft.ifft(D * F["FVideoPilot"]).real
    if not dec.analog_audio:
        return v, None
    a = {}
    for ch, filt in (("audio_left", "audio_lfilt"), ("audio_right", "audio_rfilt")):
        a[ch] = fm_discriminate(np.fft.ifft(dec._slice1(X) * F[filt]), F["freq_arf"]) + F["audio_lowfreq"]
    return v, a


VIDEO_PLANES_NTSC = ("demod", "demod_05", "demod_sync", "demod_burst")
VIDEO_PLANES_PAL = VIDEO_PLANES_NTSC + ("demod_pilot",)


def planes_of(system):
    return VIDEO_PLANES_PAL if system == "PAL" else VIDEO_PLANES_NTSC


# ---- second audio stage (lddecode_core.py:335-371) -----------------------------------------------
def _audio2_block(dec, audio, start):
    out = {}
    for ch in ("audio_left", "audio_right"):
        spec = np.fft.fft(audio[ch][start:start + dec.N])
        out[ch] = np.fft.ifft(dec._slice2(spec) * dec.F["audio_lpf2"]).real / dec.F["audio_fdiv2"]
    return out


def audio_phase2(dec, audio):
    L = len(audio["audio_left"])
    nout = L // dec.F["audio_fdiv2"]
    out = {ch: np.zeros(nout) for ch in ("audio_left", "audio_right")}
    askip = 64
    hop = dec.N - askip * dec.F["audio_fdiv2"]
    first = _audio2_block(dec, audio, 0)
    blen = len(first["audio_left"])
    for ch in out:
        out[ch][:blen] = first[ch]
    pos = blen
    for s in range(hop, L - hop, hop):
        blk = _audio2_block(dec, audio, s)
        for ch in out:
            out[ch][pos:pos + blen - askip] = blk[ch][askip:]
        pos += blen - askip
    blk = _audio2_block(dec, audio, L - dec.N - 1)
    for ch in out:
        out[ch][nout - (blen - askip):] = blk[ch][askip:]
    return out


# ---- overlap-save streaming (lddecode_core.py:373-427) -------------------------------------------
def demod(dec, loader, start, length, mtf_level=0):
    """loader(sample, n) -> array | None.  Returns (video dict, audio dict|None) or None."""
    end = int(start + length) + 1
    start = int(start - dec.blockcut) if start > dec.blockcut else 0
    stride = dec.N - dec.blockcut - dec.blockcut_end
    names = planes_of(dec.system)
    total = end - start + 1
    video = {k: np.zeros(total) for k in names}
    audio = None
    for i in range(start, end, stride):
        try:
            raw = loader(i, dec.N)
        except Exception:
            return None
        if raw is None:
            return None
        v, a = demodblock(dec, raw, mtf_level)
        o = i - start
        n = total - o if o + (dec.N - dec.blockcut) > total else stride
        for k in names:
            video[k][o:o + n] = v[k][dec.blockcut:dec.blockcut + n]
        if a is not None:
            ds = dec.N // len(a["audio_left"])
            if audio is None:
                audio = {k: np.zeros((end - start) // ds + 1) for k in a}
            for k in a:
                audio[k][o // ds:(o + n) // ds] = a[k][dec.blockcut // ds:(dec.blockcut + n) // ds]
    if audio is None:
        return video, None
    return video, audio_phase2(dec, audio)


# ---- helpers (lddutils.py:259-303, 83-97) --------------------------------------------------------
def inrange(a, lo, hi):
    return (a >= lo) & (a <= hi)


def calczc(data, start, target, edge="both", reverse=False, count=10):
    """First (or last) sample crossing `target` within data[start:start+count+1], linearly
    interpolated; None when there is none (lddutils.py:265-303)."""
    s = int(start)
    n = int(count + 1)
    if edge == "both":
        edge = "rising" if data[s] < target else "falling"
    win = data[s:s + n]
    hits = np.where(win >= target)[0] if edge == "rising" else np.where(win <= target)[0]
    if len(hits) == 0:
        return None
    x = s + hits[-1 if reverse else 0]
    if x == 0:
        return None
    a = data[x - 1] - target
    b = data[x] - target
    return x - 1 + (-a / (-a + b))


def scale(buf, begin, end, tgtlen):
    """Cubic interpolating spline (FITPACK splrep s=0, not-a-knot) through the integer samples
    spanning [begin, end], evaluated at tgtlen equidistant points (lddutils.py:83-97)."""
    ib, ie = int(begin), int(end)
    dist = ie - ib
    knots = np.linspace(0, dist, num=dist + 1)
    spl = interpolate.splrep(knots, buf[ib:ib + dist + 1])
    x = np.linspace(begin - ib, (end - begin) + (begin - ib), tgtlen + 1)
    return interpolate.splev(x, spl)[:-1]


def scale_notaknot(buf, begin, end, tgtlen):
    """The same function written out: textbook not-a-knot cubic spline on
    unit spacing, second-derivative form.  Used to cross-check `scale` and as the readable
    statement of what kernel (5) computes."""
    ib, ie = int(begin), int(end)
    n = ie - ib
    y = np.asarray(buf[ib:ib + n + 1], dtype=np.float64)
    A = np.zeros((3, n + 1))           # banded storage would hide the two dense end rows: solve dense-free
    rhs = np.zeros(n + 1)
    rhs[1:n] = 6 * (y[:-2] - 2 * y[1:-1] + y[2:])
    # Solve with not-a-knot rows M0-2M1+M2=0, Mn-2Mn-1+Mn-2=0 by eliminating M0 and Mn:
    # row1: M0+4M1+M2=r1 with M0=2M1-M2 -> 6M1 = r1 ; similarly row n-1 -> 6M(n-1) = r(n-1)
    lower = np.ones(n - 1)
    diag = np.full(n - 1, 4.0)
    upper = np.ones(n - 1)
    diag[0], upper[0] = 6.0, 0.0
    diag[-1], lower[-1] = 6.0, 0.0
    r = rhs[1:n].copy()
    ab = np.zeros((3, n - 1))
    ab[0, 1:] = upper[:-1]
    ab[1] = diag
    ab[2, :-1] = lower[1:]
    from scipy.linalg import solve_banded
    Mi = solve_banded((1, 1), ab, r)
    M = np.empty(n + 1)
    M[1:n] = Mi
    M[0] = 2 * M[1] - M[2]
    M[n] = 2 * M[n - 1] - M[n - 2]
    del A
    x = np.linspace(begin - ib, (end - begin) + (begin - ib), tgtlen + 1)[:-1]
    i = np.minimum(np.floor(x).astype(int), n - 1)
    t = x - i
    return (M[i] * (1 - t) ** 3 / 6 + M[i + 1] * t ** 3 / 6
            + (y[i] - M[i] / 6) * (1 - t) + (y[i + 1] - M[i + 1] / 6) * t)


# ---- sync peaks (lddecode_core.py:497-516) -------------------------------------------------------
def sync_peaks(ds, start, linelen):
    peaks = []
    i = start
    half = linelen // 2
    skip = int(linelen * .4)
    limit = len(ds) - 2 * linelen
    while i < limit:
        p = int(np.argmax(ds[i:i + half]))
        if ds[i + p] > .2:
            peaks.append(i + p)
            i += p + skip
        else:
            i += half
    return peaks


# ---- field analysis (lddecode_core.py:518-713, 889-957) ------------------------------------------
class FieldResult:
    """Plain record of what the reference's Field / FieldNTSC / FieldPAL objects expose."""
    pass


def _hsync_stats(ds, peaks):
    lv = np.array([ds[p] for p in peaks if inrange(ds[p], 0.6, 0.8)])
    med = np.median(lv)
    tol = max(np.std(lv) * 2, .01)
    return med, tol


def _regular(ds, peaks, k, med, tol):
    if k >= len(peaks) or peaks[k] > len(ds):
        return False
    return bool(inrange(ds[peaks[k]], med - tol, med + tol))


def _field_vote(dec, ds, peaks, k, med, tol):
    """lddecode_core.py:544-588."""
    if k < 11:
        return None
    vote = 0
    line0 = None
    gap1 = None
    for i in range(k - 1, k - 20, -1):
        if _regular(ds, peaks, i, med, tol):
            line0 = i
            gap1 = peaks[i + 1] - peaks[i]
            break
    if gap1 is not None and gap1 > dec.linelen * .75:
        vote -= 1
    gap2 = None
    for i in range(k, k + 20):
        if _regular(ds, peaks, i, med, tol):
            gap2 = peaks[i] - peaks[i - 1]
            break
    if gap2 is not None and gap2 > dec.linelen * .75:
        vote += 1 if dec.system == "NTSC" else -1
    if dec.system == "PAL":
        vote += 1
    return line0, vote


def find_vsyncs(dec, ds, peaks):
    """lddecode_core.py:590-636.  Returns (list of [peakidx, line0idx, istop], med, tol)."""
    if len(peaks) < 200:
        return [], None, None
    med, tol = _hsync_stats(ds, peaks)
    raw = []
    prev = 1.0
    for i, p in enumerate(peaks):
        v = ds[p]
        if v > .9 and prev < med - tol * 2:
            line0, vote = _field_vote(dec, ds, peaks, i, med, tol)   # reference crashes on None here
            if line0 is not None:
                raw.append([i, line0, vote])
        prev = v
    if len(raw) < 2:
        return raw, med, tol
    out = [list(r) for r in raw]
    for i in range(len(out)):
        if out[i][2] == 0:
            out[i][1] = -1
            if i < len(out) - 1 and raw[i + 1][2] != 0:
                out[i][2] = -out[i + 1][2]
            elif i >= 1 and raw[i - 1][2] != 0:
                out[i][2] = -out[i - 1][2]
        if out[i][1] <= 0:
            out[i][1] = out[i][0] - (6 if dec.system == "PAL" else 7)
        out[i][2] = int(out[i][2] < 0)
    return out, med, tol


def line_locations(dec, ds, peaks, vsyncs, med, tol, linecount):
    """lddecode_core.py:638-713 -> (linelocs1, linebad)."""
    L = dec.linelen
    found = {}
    lens = [L]
    prev_i = None
    prev_n = None
    for i in range(0, vsyncs[1][1]):
        medlen = np.median(lens[-25:])
        if not _regular(ds, peaks, i, med, tol):
            continue
        if prev_i is not None:
            gap = peaks[i] - peaks[prev_i]
            if inrange(gap / L, .98, 1.02):
                lens.append(gap)
                n = prev_n + 1
            else:
                n = prev_n + int(np.round(gap / medlen))
        else:
            n = int(np.round((peaks[i] - peaks[vsyncs[0][1]]) / medlen))
        found[n] = peaks[i]
        prev_i, prev_n = i, n
    filled = dict(found)
    for l in range(1, linecount + 5):
        if l in found:
            continue
        before = next((i for i in range(l, -10, -1) if i in found), None)
        after = next((i for i in range(l, linecount + 1) if i in found), None)
        if before is None:
            filled[l] = found[after] - L * (after - l)
        elif after is not None:
            avg = (found[after] - found[before]) / (after - before)
            filled[l] = found[before] + avg * (l - before)
        else:
            avg = found[before] - filled[before - 1]
            filled[l] = found[before] + avg * (l - before)
    locs = [filled[l] for l in range(1, linecount + 5)]
    bad = [l not in found for l in range(1, linecount + 5)]
    bad[:10] = [False] * 10
    return locs, bad


def refine_hsync(dec, d05, linelocs1, linebad):
    """lddecode_core.py:715-787 -> linelocs2 (linebad is updated in place)."""
    fq = dec.freq
    out = list(linelocs1)
    for i in range(len(out)):
        if i < 9:
            out[i] -= 200
        ll1 = out[i]
        zc = calczc(d05, out[i], dec.iretohz(-20), reverse=False, count=400)
        if zc is not None and not linebad[i]:
            out[i] = zc
            if i >= 10:
                w1 = d05[int(ll1 - fq * 2):int(ll1 + fq * 2)]
                w = d05[int(zc - fq * 1):int(zc + fq * 3)]
                wb = d05[int(zc + fq * 1):int(zc + fq * 3)]
                if ((np.min(w) < dec.iretohz(-60) or np.max(w) > dec.iretohz(20))
                        or (np.min(w1) < dec.iretohz(-60) or np.max(w1) > dec.iretohz(100))
                        or (np.min(wb) < dec.iretohz(-10) or np.max(wb) > dec.iretohz(10))):
                    linebad[i] = True
                else:
                    mid = (np.mean(w[0:20]) + np.mean(w[100:120])) / 2
                    zc2 = calczc(w, 0, mid, reverse=False, count=len(w))
                    zc2 += int(zc) - fq * 1
                    if np.abs(zc2 - zc) < fq / 4:
                        out[i] = zc2
                    else:
                        linebad[i] = True
        else:
            linebad[i] = True
        if i < 10:
            out[i] += 4.72 * fq
        if i > 10 and linebad[i]:
            out[i] = out[i - 1] + (out[i - 1] - out[i - 2])
    lo, hi = dec.linelen - fq * .2, dec.linelen + fq * .2
    for i in range(9, -1, -1):
        gap = out[i + 1] - out[i]
        if not inrange(gap, lo, hi):
            gap = dec.linelen
        out[i] = out[i + 1] - gap
    for i in range(len(out) - 10, len(out)):
        gap = out[i] - out[i - 1]
        if not inrange(gap, lo, hi):
            gap = dec.linelen
        out[i] = out[i - 1] + gap
    return out


def resample_lines(dec, plane, lineinfo, linecount, lineoffset, outwidth, wow=True):
    """Field.downscale without the audio leg (lddecode_core.py:789-812)."""
    out = np.zeros(linecount * outwidth)
    for l in range(lineoffset, linecount + lineoffset):
        seg = scale(plane, lineinfo[l], lineinfo[l + 1], outwidth)
        if wow:
            seg = seg * ((lineinfo[l + 1] - lineinfo[l]) / dec.linelen)
        out[(l - lineoffset) * outwidth:(l + 1 - lineoffset) * outwidth] = seg
    return out


def refine_burst_ntsc(dec, burstplane, linelocs2, linecount):
    """FieldNTSC.refine_linelocs_burst (lddecode_core.py:1054-1133) -> (linelocs3, burstlevel)."""
    W = dec.SP["outlinelen"]
    hz_ire = 1700000 / 140
    sb = resample_lines(dec, burstplane, linelocs2, linecount, 0, W, wow=True)
    locs = np.array(linelocs2, dtype=np.float64).copy()
    level = np.zeros(len(locs), dtype=np.float32)
    phase = np.zeros([len(locs), 2])
    for l in range(linecount):
        ba = sb[W * l + 20:W * l + 60].copy()
        ba -= np.mean(ba)
        level[l] = np.max(np.abs(ba))
        if (level[l] / hz_ire) > 30 or (np.std(ba) / hz_ire) < 3:
            level[l] = 0
            continue
        offs = {False: [], True: []}
        bi = 0
        while bi < len(ba):
            if np.abs(ba[bi]) > level[l] * .6:
                zc = calczc(ba, bi, 0)
                if zc is not None:
                    o = zc - ((np.floor(zc / 4) * 4) - 1)
                    if o > 3.5:
                        o -= 4
                    offs[bool(ba[bi] > 0)].append(o)
                    bi = int(zc)
            bi += 1
        if len(offs[False]) < 3 or len(offs[True]) < 3:
            continue
        mf = np.mean(np.array(offs[False][1:-1]))
        mt = np.mean(np.array(offs[True][1:-1]))
        phase[l] = (2 - mt, 2 - mf) if l % 2 else (2 - mf, 2 - mt)
    cut = phase[np.logical_or(phase[:, 0] != 0, phase[:, 1] != 0)]
    group = 0 if np.abs(np.median(cut[:, 0])) < np.abs(np.median(cut[:, 1])) else 1
    adj = phase[:, group]
    level[group::2] = -level[group::2]
    for l in range(len(locs)):
        if np.abs(adj[l]) > 2:
            level[l] = 0
            continue
        locs[l] -= adj[l] * (dec.freq / (4 * 315 / 88)) * 1
    for l in range(2, len(locs) - 1):
        if level[l] == 0:
            locs[l] = (locs[l - 1] + locs[l + 1]) / 2
    return locs, level


def refine_pilot_pal(dec, video, linelocs2):
    """FieldPAL.refine_linelocs_pilot (lddecode_core.py:962-1021)."""
    locs = np.array(linelocs2, dtype=np.float64).copy()
    fq = dec.freq
    offsets = {}
    alloffsets = []
    for l in range(len(locs)):
        a, b = int(locs[l] - 4.7 * fq), int(locs[l])
        pilot = (video["demod"][a:b] - video["demod_05"][a:b])[::-1].copy()
        offsets[l] = []
        adjfreq = fq
        if l > 1:
            adjfreq = fq / ((locs[l] - locs[l - 1]) / dec.linelen)
        i = 0
        while i < len(pilot):
            if inrange(pilot[i], -300000, -100000):
                zc = calczc(pilot, i, 0)
                if zc is not None:
                    zcp = zc / (adjfreq / 3.75)
                    offsets[l].append(zcp - np.floor(zcp))
                    i = int(zc + 1)
            i += 1
        if len(offsets) >= 3:
            offsets[l] = offsets[l][1:-1]
            if i >= 11:
                alloffsets += offsets[l]
        else:
            offsets[l] = []
    with np.errstate(all="ignore"):
        med = np.median(alloffsets) if len(alloffsets) else np.nan
    tgt = .5 if inrange(med, 0.25, 0.75) else 0
    for l in range(len(locs)):
        if len(offsets[l]):
            locs[l] += (tgt - np.median(offsets[l])) * (fq / 3.75) * .25
    return locs


def tbc_to_u16(dec, hz):
    """Final quantisation (NTSC lddecode_core.py:1139-1142, PAL :1027-1030)."""
    SP = dec.SP
    ire = (hz - SP["ire0"]) / SP["hz_ire"] - SP["vsync_ire"]
    if dec.system == "NTSC":
        sc, off = np.double(0xc800 - 0x0400) / (100 - SP["vsync_ire"]), 1024
    else:
        sc, off = np.double(0xd300 - 0x0100) / (100 - SP["vsync_ire"]), 256
    return np.uint16(np.clip(ire * sc + off, 0, 65535) + 0.5)


def decode_philips(dec, demod_plane, linelocs, linenum):
    """lddecode_core.py:814-834."""
    fq = dec.freq
    thr = dec.iretohz(50)
    cur = calczc(demod_plane, int(linelocs[linenum] + 2 * fq), thr, count=int(12 * fq))
    zcs = []
    while cur is not None:
        zcs.append((cur, demod_plane[int(cur - 0.5 * fq)] < thr))
        cur = calczc(demod_plane, cur + 1.9 * fq, thr, count=int(0.2 * fq))
    if len(zcs) != 24:
        return None
    gaps = np.diff([z[0] for z in zcs]) / fq
    if not (np.min(gaps) > 1.85 and np.max(gaps) < 2.15):
        return None
    bits = [int(z[1]) for z in zcs]
    return [bits[b] * 8 + bits[b + 1] * 4 + bits[b + 2] * 2 + bits[b + 3] for b in range(0, 24, 4)]


def decode_field(dec, video, start=0, colorlevel=1.45, colorphase=91.5, full=True):
    """Field.__init__ + FieldNTSC/FieldPAL.__init__ (lddecode_core.py:889-957, 1037-1048, 1165-1191).

    video: dict of float64 planes as returned by demod().  Returns a FieldResult."""
    r = FieldResult()
    ds = video["demod_sync"]
    L = dec.linelen
    r.valid = False
    r.peaklist = sync_peaks(ds, start, L)
    r.vsyncs, r.med_hsync, r.hsync_tolerance = find_vsyncs(dec, ds, r.peaklist)
    r.dspicture = None
    if len(r.vsyncs) == 0:
        r.nextfieldoffset = start + L * 200
        return r
    if len(r.vsyncs) == 1 or len(r.peaklist) < r.vsyncs[1][1] + 4:
        jump = r.peaklist[r.vsyncs[0][1] - 10]
        r.nextfieldoffset = start + jump if jump != 0 else start + L * 240
        return r
    r.nextfieldoffset = r.peaklist[r.vsyncs[1][1] - 10]
    r.istop = r.vsyncs[0][2]
    r.linecount = dec.SP["frame_lines"] // 2 + (1 if r.istop else 0)
    try:
        r.linelocs1, r.linebad = line_locations(dec, ds, r.peaklist, r.vsyncs, r.med_hsync, r.hsync_tolerance, r.linecount)
        r.linelocs2 = refine_hsync(dec, video["demod_05"], r.linelocs1, r.linebad)
    except Exception:
        return r
    r.linelocs = r.linelocs2
    r.linecode = {l: decode_philips(dec, video["demod"], r.linelocs, l) for l in dec.SP["philips_codelines"]}
    r.valid = True
    r.tbcstart = r.peaklist[r.vsyncs[1][1] - 10]
    if not full:
        return r
    W = dec.SP["outlinelen"]
    try:
        if dec.system == "NTSC":
            r.linelocs3, r.burstlevel = refine_burst_ntsc(dec, video["demod_burst"], r.linelocs2, r.linecount)
            r.linelocs4, r.burstlevel = refine_burst_ntsc(dec, video["demod_burst"], r.linelocs3, r.linecount)
            shift = colorphase * (np.pi / 180) - 8
            r.linelocs = np.array(r.linelocs4) + shift * (dec.freq / (4 * 315 / 88))
            hz = resample_lines(dec, video["demod"], r.linelocs, r.linecount, 1, W, wow=True)
            pic = tbc_to_u16(dec, hz)
            for i in range(1, r.linecount - 1):
                pic[i * W] = 16384 if r.burstlevel[i] > 0 else 32768
                clevel = (1 / colorlevel) / (1700000 / 140)
                pic[i * W + 1] = np.uint16(327.67 * clevel * np.abs(r.burstlevel[i]))
            r.dspicture = pic
        else:
            r.linelocs = refine_pilot_pal(dec, video, r.linelocs2)
            hz = resample_lines(dec, video["demod"], r.linelocs, r.linecount, 3, W, wow=True)
            r.dspicture = tbc_to_u16(dec, hz)
    except Exception:
        r.valid = False
    return r


# ---- 48 kHz PCM along the line positions (lddecode_core.py:431-484) ---------------------------------
def downscale_audio(dec, audio, lineinfo, linecount, timeoffset=0.0, freq=48000.0, scale=64):
    """downscale_audio (lddecode_core.py:431-484): audio = phase-2 audio of the field's read window, lineinfo = the
    field's final line positions.  Returns (int16 interleaved L/R, time offset for the next field)."""
    SP = dec.SP
    frametime = (SP["line_period"] * linecount) / 1000000
    soundgap = 1 / freq
    times = np.arange(timeoffset, frametime + soundgap, soundgap, dtype=np.double)
    n = len(times) - 1
    out = np.zeros(2 * max(n, 0), dtype=np.int32)
    for i in range(n):
        linenum = ((times[i] * 1000000) / SP["line_period"]) + 1
        li = int(linenum)
        cur = lineinfo[li]
        nxt = lineinfo[li + 1] if li + 1 < len(lineinfo) else cur + dec.linelen
        loc = cur + (nxt - cur) * (linenum - np.floor(linenum))
        wow = (nxt - cur) / dec.linelen
        k = int(loc / scale)
        out[2 * i] = int(np.round((audio["audio_left"][k] * wow - SP["audio_lfreq"]) * 32767 / 150000))
        out[2 * i + 1] = int(np.round((audio["audio_right"][k] * wow - SP["audio_rfreq"]) * 32767 / 150000))
    return np.clip(out, -32766, 32766).astype(np.int16), times[-1] - frametime


def framer_audio_walk(dec, loader, nfields, readlen=1000000, mtf_level=1, firstframe=True):
    """The audio side of Framer.readframe called in a loop (lddecode_core.py:1194-1223, 1256-1289; CLV pairing by field
    parity): every field of one readframe call is built with the Framer's audio_offset as the call found it, the field that
    closes the frame hands its audio_next_offset on, and fields read ahead of the very first frame are not written.
    Returns [(readsample, FieldResult, pcm | None)] for the first nfields valid fields, and the final audio_offset."""
    topfirst = dec.SP["topfirst"]
    out, offset, rs, fieldcount = [], 0.0, 0, 0
    while len(out) < nfields:
        d = demod(dec, loader, rs, readlen, mtf_level)
        if d is None:
            break
        f = decode_field(dec, d[0], 0)
        here = rs
        rs += f.nextfieldoffset
        if not f.valid:
            if len(f.peaklist) < 100:                # readfield's jumps over unreadable stretches (:1208-1213)
                rs = here + int(dec.freq_hz * 10)
            elif len(f.vsyncs) == 0:
                rs = here + int(dec.freq_hz * 1)
            continue
        pcm, nxt = downscale_audio(dec, d[1], f.linelocs, f.linecount, offset)
        if f.istop == topfirst:
            fieldcount = 1
        elif fieldcount == 1:
            fieldcount = 2
        out.append((here, f, pcm if (fieldcount or not firstframe) else None))
        if fieldcount == 2:
            offset, fieldcount, firstframe = nxt, 0, False
    return out, offset


# ---- integer unpackers (ddunpack.c:11-36, lddutils.py:150-229) -----------------------------------
def unpack_r30_raw(words, offset, n):
    """lddutils.py:150-173: three 10-bit fields per LE u32, raw 0..1023 as int16."""
    w = np.asarray(words, dtype=np.uint32)
    out = np.empty(len(w) * 3, dtype=np.int16)
    out[0::3] = w & 0x3ff
    out[1::3] = (w >> 10) & 0x3ff
    out[2::3] = (w >> 20) & 0x3ff
    return out[offset:offset + n]


def unpack_r30_ddunpack(words):
    """ddunpack.c:11-21,28-36: ((field) - 512) << 6 as int16."""
    raw = unpack_r30_raw(words, 0, len(words) * 3).astype(np.int32)
    return ((raw - 512) << 6).astype(np.int16)


def unpack_lds(data, offset, n):
    """lddutils.py:195-229: four 10-bit samples in five bytes, MSB first."""
    b = np.asarray(data, dtype=np.uint8).astype(np.uint16)
    g = len(b) // 5
    b = b[:g * 5].reshape(g, 5)
    out = np.empty((g, 4), dtype=np.uint16)
    out[:, 0] = (b[:, 0] << 2) | (b[:, 1] >> 6)
    out[:, 1] = ((b[:, 1] & 0x3f) << 4) | (b[:, 2] >> 4)
    out[:, 2] = ((b[:, 2] & 0x0f) << 6) | (b[:, 3] >> 2)
    out[:, 3] = ((b[:, 3] & 0x03) << 8) | b[:, 4]
    return out.reshape(-1)[offset:offset + n]

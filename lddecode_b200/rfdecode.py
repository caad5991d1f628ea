"""RFDecode: host-side mirror of the reference's demodulator class, backed by libldd_b200.so.

Keeps the reference's public surface for this path (lddecode_core.py:119-427): constructor
arguments, the attributes callers read (blocklen, blockcut, blockcut_end, freq, freq_hz,
linelen, SysParams, DecoderParams, Filters ...), and the methods computefilters(),
demodblock(), demod(), audio_phase2(), iretohz(), hztoire() with the same argument meaning,
return layout (numpy record arrays in Hz) and error behaviour (None on a short read).
Filter tables are built on the host with scipy exactly as the reference does and uploaded;
all per-sample work runs in the CUDA kernels.  PyTorch is used for device buffers only.
"""
import ctypes as C
import os

import numpy as np
import scipy.signal as sps

from . import _lib
from ._backend import CudaBackend

DEFAULT_PRECISION = 'mixed'

# module-global loader, as in the reference (lddecode_core.py:387, assigned at lddecode.py:53-58):
# loader(infile, sample, readlen) -> array | None
loader = None

_FSC_NTSC = 315.0 / 88.0

# lddecode_core.py:30-56
SysParams_NTSC = {
    'fsc_mhz': _FSC_NTSC, 'pilot_mhz': _FSC_NTSC, 'frame_lines': 525,
    'ire0': 8100000, 'hz_ire': 1700000 / 140.0, 'vsync_ire': -40, 'analog_audio': True,
    'audio_lfreq': (1000000 * 315 / 88 / 227.5) * 146.25, 'audio_rfreq': (1000000 * 315 / 88 / 227.5) * 178.75,
    'philips_codelines': [16, 17, 18], 'topfirst': True,
    'line_period': 1 / (_FSC_NTSC / 227.5),
}
SysParams_NTSC['FPS'] = 1000000 / (525 * SysParams_NTSC['line_period'])
SysParams_NTSC['outlinelen'] = int(np.round(SysParams_NTSC['line_period'] * SysParams_NTSC['fsc_mhz'] * 4))

# lddecode_core.py:58-84
SysParams_PAL = {
    'FPS': 25, 'fsc_mhz': ((1 / 64) * 283.75) + (25 / 1000000), 'pilot_mhz': 3.75, 'frame_lines': 625,
    'line_period': 64, 'ire0': 7100000, 'hz_ire': 800000 / 100.0, 'analog_audio': True,
    'audio_lfreq': (1000000 / 64) * 43.75, 'audio_rfreq': (1000000 / 64) * 68.25,
    'philips_codelines': [19, 20, 21], 'topfirst': False,
}
SysParams_PAL['outlinelen'] = int(np.round(64 * SysParams_PAL['fsc_mhz'] * 4))
SysParams_PAL['outlinelen_pilot'] = int(np.round(64 * SysParams_PAL['pilot_mhz'] * 4))
SysParams_PAL['vsync_ire'] = -.3 * (100 / .7)

# lddecode_core.py:86-117
RFParams_NTSC = {
    'audio_notchwidth': 350000, 'audio_notchorder': 2, 'video_deemp': (120 * .32, 320 * .32),
    'video_bpf': [3500000, 13200000], 'video_bpf_order': 3, 'video_lpf_freq': 4200000, 'video_lpf_order': 5,
}
RFParams_PAL = {
    'audio_notchwidth': 200000, 'audio_notchorder': 2, 'video_deemp': (100 * .4, 400 * .4),
    'video_bpf': (2500000, 14500000), 'video_bpf_order': 3, 'video_lpf_freq': 5200000, 'video_lpf_order': 9,
}

VIDEO_FIELDS = {'NTSC': ['demod', 'demod_05', 'demod_sync', 'demod_burst'],
                'PAL': ['demod', 'demod_05', 'demod_sync', 'demod_burst', 'demod_pilot']}
_PLANE_OF = {'demod': _lib.P_DEMOD, 'demod_05': _lib.P_DEMOD05, 'demod_sync': _lib.P_SYNC,
             'demod_burst': _lib.P_BURST, 'demod_pilot': _lib.P_PILOT}
_REL_IRE0 = ('demod', 'demod_05')      # planes the library stores relative to ire0 (include/ldd_b200.h)

_FMT_OF_DTYPE = {np.dtype('uint8'): _lib.FMT_U8, np.dtype('int16'): _lib.FMT_S16, np.dtype('uint16'): _lib.FMT_U16}


def calclinelen(SP, mult, mhz):
    """lddecode_core.py:23-27."""
    if type(mhz) == str:
        mhz = SP[mhz]
    return int(np.round(SP['line_period'] * mhz * mult))


def filtfft(filt, blocklen):
    """Frequency response of a (b, a) filter on `blocklen` points of the unit circle (lddutils.py:256-257)."""
    return sps.freqz(filt[0], filt[1], blocklen, whole=1)[1]


def hilbert_taps(terms=128):
    """lddutils.py:246-249."""
    return np.fft.fftshift(np.fft.ifft([0] + [1] * terms + [0] * terms))


class DeviceDemod:
    """Result of a demodulation kept in device memory: float32 planes + float64 audio."""

    def __init__(self, rf, planes, audio, rng):
        self.rf = rf
        self.planes = planes          # dict name -> device buffer (float32; demod/demod_05 relative to ire0)
        self.audio = audio            # None | dict name -> device buffer (float64, after phase 2)
        self.range = rng
        self.length = int(rng.total_out)

    def plane_host(self, name):
        """float64 numpy array in the reference's units (Hz, absolute)."""
        v = self.rf._be.to_host(self.planes[name]).astype(np.float64)
        if name in _REL_IRE0:
            v += self.rf.SysParams['ire0']
        return v

    def audio_recarray(self):
        if self.audio is None:
            return None
        return np.rec.array([self.rf._be.to_host(self.audio['audio_left']), self.rf._be.to_host(self.audio['audio_right'])],
                            names=['audio_left', 'audio_right'])

    def to_recarrays(self):
        names = VIDEO_FIELDS[self.rf.system]
        video = np.rec.array([self.plane_host(n) for n in names], names=names)
        audio = None
        if self.audio is not None:
            audio = np.rec.array([self.rf._be.to_host(self.audio['audio_left']),
                                  self.rf._be.to_host(self.audio['audio_right'])],
                                 names=['audio_left', 'audio_right'])
        return video, audio


class RFDecode:
    def __init__(self, inputfreq=40, system='NTSC', blocklen_=16384, decode_analog_audio=True,
                 have_analog_audio=True, device=None, precision=None, _backend=None):
        self.blocklen = blocklen_
        self.blockcut = 1024
        self.system = system
        self.freq = inputfreq
        self.freq_half = inputfreq / 2
        self.freq_hz = self.freq * 1000000
        self.freq_hz_half = self.freq * 1000000 / 2
        if system == 'NTSC':
            self.SysParams, self.DecoderParams = SysParams_NTSC, RFParams_NTSC
        elif system == 'PAL':
            self.SysParams, self.DecoderParams = SysParams_PAL, RFParams_PAL
        else:
            raise ValueError("system must be 'NTSC' or 'PAL'")
        self.linelen = int(np.round(self.freq_hz / (1000000.0 / self.SysParams['line_period'])))
        self.decode_analog_audio = decode_analog_audio
        # precision lanes: 'f64' exact, 'f32' fast, 'mixed' = f32 + f64 re-run of the
        # blocks whose sync decisions float32 could get wrong (sync plane bit-identical to 'f64')
        self.precision = precision or os.environ.get("LDD_PRECISION", DEFAULT_PRECISION)
        self._be = _backend if _backend is not None else CudaBackend(device)
        self._h = None
        self._mtf_uploaded = None
        self.computefilters()
        self.blockcut_end = self.Filters['F05_offset']
        self._create_handle()

    # ---- filter tables (host, scipy), lddecode_core.py:147-279 ---------------------------------
    def computefilters(self):
        self.computevideofilters()
        if self.decode_analog_audio:
            self.computeaudiofilters()
        self._mtf_uploaded = None
        if self._h is not None:
            self._upload_filters()

    def computevideofilters(self):
        N, SP, DP = self.blocklen, self.SysParams, self.DecoderParams
        nyq = self.freq_hz_half
        resp = lambda ba: filtfft(ba, N)
        angles = (12.5, 27.5) if self.system == 'NTSC' else (10, 28)
        SF = self.Filters = {'MTF': resp(sps.zpk2tf([], [.7 * np.exp(1j * np.pi * a / 20) for a in angles], 1.11))}
        SF['hilbert'] = np.fft.fft(hilbert_taps(), N)
        SF['RFVideo'] = resp(sps.butter(DP['video_bpf_order'], [DP['video_bpf'][0] / nyq, DP['video_bpf'][1] / nyq],
                                        btype='bandpass'))
        if SP['analog_audio']:
            for key, name in (('audio_lfreq', 'Fcutl'), ('audio_rfreq', 'Fcutr')):
                edges = [(SP[key] - DP['audio_notchwidth']) / nyq, (SP[key] + DP['audio_notchwidth']) / nyq]
                SF[name] = resp(sps.butter(DP['audio_notchorder'], edges, btype='bandstop'))
            SF['RFVideo'] *= (SF['Fcutl'] * SF['Fcutr'])
        SF['RFVideo'] *= SF['hilbert']
        SF['Fvideo_lpf'] = resp(sps.butter(DP['video_lpf_order'], DP['video_lpf_freq'] / nyq, 'low'))
        d0, d1 = DP['video_deemp']
        for name, (z, p_, k) in (('Fdeemp', (-d1 * (10 ** -10), -d0 * (10 ** -10), d0 / d1)),
                                 ('Femp', (-d0 * (10 ** -10), -d1 * (10 ** -10), d1 / d0))):
            tf_b, tf_a = sps.zpk2tf(z, p_, k)
            SF[name] = resp(sps.bilinear(tf_b, tf_a, 1.0 / nyq))
        SF['FVideo'] = SF['Fvideo_lpf'] * SF['Fdeemp']
        SF['F05_offset'] = 32
        SF['F05'] = resp((sps.firwin(65, [0.5 / self.freq_half], pass_zero=True), [1.0]))
        SF['FVideo05'] = SF['Fvideo_lpf'] * SF['Fdeemp'] * SF['F05']
        band = lambda lo, hi: resp(sps.butter(1, [lo / self.freq_half, hi / self.freq_half], btype='bandpass'))
        SF['Fburst'] = band(SP['fsc_mhz'] - .1, SP['fsc_mhz'] + .1)
        SF['FVideoBurst'] = SF['Fvideo_lpf'] * SF['Fdeemp'] * SF['Fburst']
        if self.system == 'PAL':
            SF['Fpilot'] = band(3.7, 3.8)
            SF['FVideoPilot'] = SF['Fvideo_lpf'] * SF['Fdeemp'] * SF['Fpilot']
        self._fpsync_ba = sps.butter(1, 0.05 / self.freq_half, btype='low')
        SF['FPsync'] = resp(self._fpsync_ba)

    def audio_fdslice(self, freqdomain):
        SF = self.Filters
        return np.concatenate([freqdomain[SF['audio_fdslice_lo']], freqdomain[SF['audio_fdslice_hi']]])

    def audio_fdslice2(self, freqdomain):
        SF = self.Filters
        return np.concatenate([freqdomain[SF['audio_fdslice2_lo']], freqdomain[SF['audio_fdslice2_hi']]])

    def computeaudiofilters(self):
        N, SF, SP = self.blocklen, self.Filters, self.SysParams
        nyq = self.freq_hz_half
        fdiv1 = 32 if self.freq >= 32 else 16
        halfwidth = N // (fdiv1 * 2)
        SF['freq_arf'] = self.freq_hz / (fdiv1 / 2)
        SF['audio_fdiv1'] = fdiv1
        SP['audio_cfreq'] = (SP['audio_rfreq'] + SP['audio_lfreq']) // 2
        center = int((SP['audio_cfreq'] / self.freq_hz) * N)
        lo, hi = int(center - halfwidth), int(center + halfwidth)
        SF['audio_fdslice_lo'] = slice(lo, hi)
        SF['audio_fdslice_hi'] = slice(N - hi, N - lo)
        SF['audio_lowfreq'] = SP['audio_cfreq'] - (self.freq_hz / (2 * fdiv1))
        apass, ntaps = 150000, 800
        for key, name in (('audio_lfreq', 'audio_lfilt'), ('audio_rfreq', 'audio_rfilt')):
            taps = sps.firwin(ntaps, [(SP[key] - apass) / nyq, (SP[key] + apass) / nyq], pass_zero=False)
            SF[name] = self.audio_fdslice(filtfft([taps, 1.0], N) * SF['hilbert'])
        SF['audio_fdiv2'] = 4
        SF['audio_fdiv'] = fdiv1 * 4
        SF['freq_aud2'] = SF['freq_arf'] / 4
        SF['audio_fdslice2_lo'] = slice(0, N // 8)
        SF['audio_fdslice2_hi'] = slice(N - N // 8, N)
        SF['audio_lpf2'] = filtfft([sps.firwin(65, [21000 / (SF['freq_aud2'] / 2)]), [1.0]], N // 4)
        d75freq = 1000000 / (2 * np.pi * 75)
        SF['audio_deemp2'] = filtfft(sps.butter(1, [d75freq / (SF['freq_aud2'] / 2)], btype='lowpass'), N // 4)

    @property
    def audio_decimation(self):
        """RF samples per sample of the phase-2 audio demod() returns: the first stage keeps 2 * blocklen / audio_fdiv1 bins
        of the block's spectrum (lddecode_core.py:253-259), the second every fourth sample (:271, 343) -- 32 at 8fsc NTSC,
        64 from 32 MSPS up.  (downscale_audio's scale=64 and Filters['audio_fdiv'] assume the latter.)"""
        return 4 * (self.blocklen // len(self.Filters['audio_lfilt']))

    def iretohz(self, ire):
        return self.SysParams['ire0'] + (self.SysParams['hz_ire'] * ire)

    def hztoire(self, hz):
        return (hz - self.SysParams['ire0']) / self.SysParams['hz_ire']

    # ---- device handle ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self._be.lib.ldd_last_error(self._h) if self._h else b""
            raise _lib.LddError(rc, (msg or b"").decode())

    def _create_handle(self):
        SF, SP = self.Filters, self.SysParams
        cfg = _lib.Config()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.device = self._be.device_index
        cfg.system = _lib.SYSTEM[self.system]
        cfg.blocklen, cfg.blockcut, cfg.blockcut_end = self.blocklen, self.blockcut, self.blockcut_end
        cfg.f05_offset = SF['F05_offset']
        cfg.precision = {'f64': _lib.PREC_F64, 'f32': _lib.PREC_F32, 'mixed': _lib.PREC_MIXED}[self.precision]
        cfg.decode_analog_audio = int(bool(self.decode_analog_audio))
        if self.decode_analog_audio:
            cfg.audio_slice_lo, cfg.audio_slice_hi = SF['audio_fdslice_lo'].start, SF['audio_fdslice_lo'].stop
            cfg.freq_arf, cfg.audio_lowfreq = SF['freq_arf'], SF['audio_lowfreq']
        cfg.linelen, cfg.outlinelen = self.linelen, SP['outlinelen']
        cfg.freq_hz = self.freq_hz
        cfg.ire0, cfg.hz_ire, cfg.vsync_ire = SP['ire0'], SP['hz_ire'], SP['vsync_ire']
        cfg.sync_lo_hz, cfg.sync_hi_hz = self.iretohz(-55), self.iretohz(-25)
        b, a = self._fpsync_ba
        cfg.fpsync_b0, cfg.fpsync_b1, cfg.fpsync_a1 = b[0], b[1], a[1]
        h = C.c_void_p()
        rc = self._be.lib.ldd_create(C.byref(cfg), C.byref(h))
        self._h = h if h.value else None
        self._check(rc)
        self._upload_filters()

    def set_blockcut(self, blockcut):
        """RFDecode.blockcut is a plain instance attribute in the reference (lddecode_core.py:122); here the
        device handle has to follow it."""
        self.blockcut = int(blockcut)
        if self._h is not None:
            self._be.lib.ldd_destroy(self._h)
            self._h = None
        self._create_handle()

    def _set_filter(self, fid, table):
        t = np.ascontiguousarray(table, dtype=np.complex128)
        self._check(self._be.lib.ldd_set_filter(self._h, fid, t.ctypes.data_as(C.c_void_p), len(t)))

    def _upload_filters(self):
        SF = self.Filters
        self._set_filter(_lib.F_VIDEO, SF['FVideo'])
        self._set_filter(_lib.F_VIDEO05, SF['FVideo05'])
        self._set_filter(_lib.F_BURST, SF['FVideoBurst'])
        if self.system == 'PAL':
            self._set_filter(_lib.F_PILOT, SF['FVideoPilot'])
        if self.decode_analog_audio:
            self._set_filter(_lib.F_AUDIO_L, SF['audio_lfilt'])
            self._set_filter(_lib.F_AUDIO_R, SF['audio_rfilt'])
            self._set_filter(_lib.F_AUDIO_LPF2, SF['audio_lpf2'])
        self._mtf_uploaded = None

    def _set_mtf(self, mtf_level):
        """RFVideo * MTF**mtf_level (lddecode_core.py:290-293).  The two tables are uploaded once; a level change is
        one small kernel on the current stream (ldd_set_mtf_level), stream-ordered with the demodulations around it."""
        if self._mtf_uploaded is None:
            self._set_filter(_lib.F_RFVIDEO, self.Filters['RFVideo'])
            self._set_filter(_lib.F_MTF, self.Filters['MTF'])
            self._mtf_uploaded = True
        self._check(self._be.lib.ldd_set_mtf_level(self._h, float(mtf_level), self._be.stream()))     # no-op when unchanged

    def set_mtf_ramp(self, pos0_sample=0.0, period_samples=0.0, step_per_period=0.0, hold_until_sample=-1e300, hold_level=1.0):
        """Per-block MTF level for whole-range decodes of CAV discs (ldd_set_mtf_ramp): blocks in the n-th frame period
        after capture sample pos0_sample use max(mtf_level + n * step_per_period, 0); blocks before hold_until_sample use
        hold_level (the reference's start-up level for the first frame of a run).  period_samples = 0: off."""
        if self._mtf_uploaded is None:
            self._set_mtf(0)
        self._check(self._be.lib.ldd_set_mtf_ramp(self._h, float(pos0_sample), float(period_samples), float(step_per_period),
                                                  float(hold_until_sample), float(hold_level)))

    def __del__(self):
        try:
            if self._h is not None:
                self._be.lib.ldd_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- demodulation ------------------------------------------------------------------------------
    def range_query(self, start, length):
        r = _lib.Range()
        self._check(self._be.lib.ldd_demod_range_query(self._h, int(start), int(length), C.byref(r)))
        return r

    def _alloc_planes(self, total, alloc=None):
        names = VIDEO_FIELDS[self.system]
        if alloc is None:
            alloc = lambda name, n, dt: self._be.empty(n, dt)
        bufs = {n: alloc(n, total, np.float64 if n == 'demod_sync' else np.float32) for n in names}
        arr = (C.c_void_p * 5)()
        for n in names:
            arr[_PLANE_OF[n]] = self._be.ptr(bufs[n])
        return bufs, arr

    def demod_device(self, rf_buf, fmt, rf_base, rf_len, start, length, mtf_level=0, phase2=True):
        """demod(start, length) on a capture that is already in device memory.

        rf_buf: device buffer holding capture samples [rf_base, rf_base+rf_len) in format `fmt`.
        Returns a DeviceDemod, or None when the capture is too short (the reference returns None)."""
        self._set_mtf(mtf_level)
        r = self.range_query(start, length)
        planes, parr = self._alloc_planes(r.total_out)
        a1l = a1r = None
        if self.decode_analog_audio:
            a1l = self._be.empty(r.audio1_len, np.float64)
            a1r = self._be.empty(r.audio1_len, np.float64)
        rc = self._be.lib.ldd_demod_range(self._h, self._be.ptr(rf_buf), fmt, int(rf_base), int(rf_len), int(start),
                                          int(length), parr, self._be.ptr(a1l) if a1l is not None else None,
                                          self._be.ptr(a1r) if a1r is not None else None, self._be.stream())
        if rc == _lib.ESHORT:
            return None
        self._check(rc)
        audio = None
        if self.decode_analog_audio:
            if phase2:
                audio = self._audio_phase2_device(a1l, a1r, r.audio1_len)
            else:
                audio = {'audio_left': a1l, 'audio_right': a1r}
        return DeviceDemod(self, planes, audio, r)

    def mixed_stats(self):
        """(blocks re-run in float64, blocks) of the last demodulation in precision='mixed'."""
        self._be.synchronize()
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._check(self._be.lib.ldd_mixed_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def _audio_phase2_device(self, a1l, a1r, n):
        outl = self._be.empty(n // 4, np.float64)
        outr = self._be.empty(n // 4, np.float64)
        self._check(self._be.lib.ldd_audio_phase2(self._h, self._be.ptr(a1l), self._be.ptr(a1r), int(n),
                                                  self._be.ptr(outl), self._be.ptr(outr), self._be.stream()))
        return {'audio_left': outl, 'audio_right': outr}

    def demod(self, infile, start, length, mtf_level=0):
        """RFDecode.demod (lddecode_core.py:373-427): (video recarray, audio recarray | None) or None.

        The samples of all blocks are fetched with one call of the module-global `loader`
        (same contract as the reference's, lddutils.py:117-129) instead of one call per block."""
        out = self.demod_raw(infile, start, length, mtf_level)
        if out is None:
            return None
        return out.to_recarrays()

    def demod_raw(self, infile, start, length, mtf_level=0):
        """Like demod() but the result stays in device memory (a DeviceDemod, which Field accepts)."""
        r = self.range_query(start, length)
        need = int(r.last_needed - r.first_sample)
        try:
            data = loader(infile, int(r.first_sample), need)
        except Exception:
            return None
        if data is None or len(data) < need:
            return None
        data = np.ascontiguousarray(data[:need])
        if data.dtype not in _FMT_OF_DTYPE:
            data = data.astype(np.int16)
        dev = self._be.to_device(data)
        return self.demod_device(dev, _FMT_OF_DTYPE[data.dtype], r.first_sample, need, start, length, mtf_level)

    def demodblock(self, data, mtf_level=0):
        """RFDecode.demodblock (lddecode_core.py:288-330) for one block of `blocklen` samples:
        (video recarray, audio recarray | None), every field `blocklen` long."""
        N = self.blocklen
        data = np.ascontiguousarray(np.asarray(data)[:N])
        if len(data) < N:
            raise ValueError("demodblock needs blocklen samples")
        if data.dtype not in _FMT_OF_DTYPE:
            data = data.astype(np.int16)
        self._set_mtf(mtf_level)
        dev = self._be.to_device(data)
        planes, parr = self._alloc_planes(N)
        al = ar = None
        if self.decode_analog_audio:
            A = len(self.Filters['audio_lfilt'])
            al, ar = self._be.empty(A, np.float64), self._be.empty(A, np.float64)
        self._check(self._be.lib.ldd_demodblock(self._h, self._be.ptr(dev), _FMT_OF_DTYPE[data.dtype], N, parr,
                                                self._be.ptr(al) if al is not None else None,
                                                self._be.ptr(ar) if ar is not None else None, self._be.stream()))
        rng = _lib.Range()
        rng.total_out = N
        audio = None if al is None else {'audio_left': al, 'audio_right': ar}
        return DeviceDemod(self, planes, audio, rng).to_recarrays()

    def runfilter_audio_phase2(self, frame_audio, start):
        """RFDecode.runfilter_audio_phase2 (lddecode_core.py:335-346): ONE block of the second audio stage, blocklen
        phase-1 samples from `start` -> blocklen/4 samples, nothing cut.  (ldd_audio_phase2 over blocklen + 1 samples is
        exactly that block: the first block is kept whole and the re-anchored last block is the same block.)"""
        N = self.blocklen
        n = len(frame_audio['audio_left'])
        if start < 0 or start + N > n:
            raise ValueError("runfilter_audio_phase2: the block leaves the array (the reference raises on the shape mismatch)")
        pad = lambda ch: np.concatenate([np.asarray(frame_audio[ch][start:start + N], dtype=np.float64), [0.0]])
        out = self._audio_phase2_device(self._be.to_device(pad('audio_left')), self._be.to_device(pad('audio_right')), N + 1)
        return np.rec.array([self._be.to_host(out['audio_left']), self._be.to_host(out['audio_right'])],
                            names=['audio_left', 'audio_right'])

    def audio_phase2(self, field_audio):
        """RFDecode.audio_phase2 (lddecode_core.py:348-371) on a host record array."""
        l = self._be.to_device(np.ascontiguousarray(field_audio['audio_left'], dtype=np.float64))
        r = self._be.to_device(np.ascontiguousarray(field_audio['audio_right'], dtype=np.float64))
        out = self._audio_phase2_device(l, r, len(field_audio['audio_left']))
        return np.rec.array([self._be.to_host(out['audio_left']), self._be.to_host(out['audio_right'])],
                            names=['audio_left', 'audio_right'])

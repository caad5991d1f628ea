"""Whole-capture decode with everything resident in HBM: the throughput path behind bench.py.

The reference walks a capture field by field (Framer.readfield, lddecode_core.py:1194-1223): every
field re-reads and re-demodulates 1e6 samples (2.1x redundant) and the next read position comes out
of the previous field.  Here the capture is demodulated ONCE on a fixed global block grid
(blocks are independent given the 1024/32-sample halos), one global sync-peak
chase replaces the per-field ones, the field-to-field walk runs on the host over the ~16 K peaks per
second of video (ldd_field_chain), and all located fields are refined and resampled in a handful of
batched launches.  Only the peak list crosses to the host in the middle.

Differences from calling Field() per window, by construction: planes come from one block grid
instead of one grid per field, so demod_sync differs by the FPsync wrap term (< 3e-5 absolute);
peak indices, line tables and TBC output are compared against the reference flow in tests/.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import field as F

READLEN = 1000000        # Framer.readlen (lddecode_core.py:1319, 1324)


class CaptureResult:
    pass


class CaptureDecoder:
    def __init__(self, rf, readlen=READLEN, mtf_level=1, colorlevel=1.45, colorphase=91.5, max_fields=4096):
        self.rf = rf
        self.readlen = readlen
        self.mtf_level = mtf_level          # Framer starts at 1 (lddecode_core.py:1334)
        self.colorlevel, self.colorphase = colorlevel, colorphase
        self.max_fields = max_fields

    # -- stage 1: planes + peaks on the device
    def demod_all(self, cap_dev, fmt, ncap, rf_base=0):
        rf, be = self.rf, self.rf._be
        N = rf.blocklen
        S = N - rf.blockcut - rf.blockcut_end
        nblocks = (ncap - N) // S + 1 if ncap >= N else 0
        total = nblocks * S
        rf._set_mtf(self.mtf_level)
        planes, parr = rf._alloc_planes(max(total, 1))
        a1l = a1r = None
        alen = 0
        if rf.decode_analog_audio:
            ds = N // len(rf.Filters['audio_lfilt'])
            alen = total // ds
            a1l, a1r = be.empty(max(alen, 1), np.float64), be.empty(max(alen, 1), np.float64)
        if nblocks:
            rf._check(be.lib.ldd_demod_blocks(rf._h, be.ptr(cap_dev), fmt, int(rf_base), int(ncap), int(rf_base), int(nblocks),
                                              int(total), parr, be.ptr(a1l) if a1l is not None else None,
                                              be.ptr(a1r) if a1r is not None else None, int(alen), be.stream()))
        return planes, total, (a1l, a1r, alen)

    def decode(self, cap_dev, fmt, ncap, want_tables=False, audio_phase2=True):
        """cap_dev: device buffer with the whole capture (format fmt, ncap samples)."""
        rf, be = self.rf, self.rf._be
        res = CaptureResult()
        planes, total, (a1l, a1r, alen) = self.demod_all(cap_dev, fmt, ncap)
        res.planes, res.plane_len = planes, total
        res.audio = None
        if rf.decode_analog_audio and audio_phase2 and alen > rf.blocklen:
            res.audio = rf._audio_phase2_device(a1l, a1r, alen)
        elif rf.decode_analog_audio:
            res.audio = {'audio_left': a1l, 'audio_right': a1r}
        # global peak chase, then the only device->host hop of the path: ~16 K peaks per second of video
        gpk, gvl = F.sync_peaks_device(rf, planes['demod_sync'], total, 0)
        res.gpeaks = gpk
        batch, infos, readsamples = self._walk(planes, total, ncap, gpk, gvl)
        res.infos, res.readsamples = infos, readsamples
        res.nwindows = len(infos)
        located = [i for i, f in enumerate(infos) if f.stage == _lib.FIELD_LOCATED]
        res.located = located
        res.batch = batch
        if located:
            sub = F.FieldBatch(rf, len(located))
            for j, i in enumerate(located):
                sub.base[j], sub.winlen[j], sub.linecount[j] = batch.base[i], batch.winlen[i], infos[i].linecount
                sub.linelocs1[j], sub.linebad[j] = batch.linelocs1[i], batch.linebad[i]
            ref = F.refine_and_tbc(rf, planes, total, sub, self.colorlevel, self.colorphase, want_intermediates=want_tables)
            res.refined = ref
            res.d_pic, res.out_stride, res.d_status = ref.d_pic, ref.out_stride, ref.d_status
        else:
            res.refined = None
            res.d_pic = None
        return res

    # -- stage 2: host walk
    def _walk(self, planes, total, ncap, gpk, gvl):
        rf, be = self.rf, self.rf._be
        mf = self.max_fields
        fields = (_lib.FieldInfo * mf)()
        batch = F.FieldBatch(rf, mf)
        readsample = np.zeros(mf, dtype=np.int64)
        nf = C.c_int(0)
        keep = {}

        def window_peaks(ctx, b, wl, ppk, pvl, pn):
            # slow path: this window does not start on a peak of the global chase
            sync = planes['demod_sync'][int(b):int(b + wl)]
            pk, vl = F.sync_peaks_device(rf, sync, int(wl), 0)
            keep['pk'], keep['vl'] = np.ascontiguousarray(pk), np.ascontiguousarray(vl)
            ppk[0] = keep['pk'].ctypes.data
            pvl[0] = keep['vl'].ctypes.data
            pn[0] = len(pk)
            return 0

        cb = _lib.WINDOW_PEAKS_FN(window_peaks)
        gpk = np.ascontiguousarray(gpk, dtype=np.int64)
        gvl = np.ascontiguousarray(gvl, dtype=np.float64)
        rf._check(be.lib.ldd_field_chain(rf._h, F._h(gpk), F._h(gvl), len(gpk), int(total), int(ncap), int(self.readlen), 0, mf,
                                         C.cast(cb, C.c_void_p), None, C.cast(fields, C.c_void_p), F._h(batch.base),
                                         F._h(batch.winlen), F._h(readsample), F._h(batch.linelocs1.reshape(-1)),
                                         F._h(batch.linebad.reshape(-1)), F.LL_STRIDE, C.byref(nf)))
        n = nf.value
        infos = [fields[i] for i in range(n)]
        return batch, infos, readsample[:n].copy()

    # -- host copies
    def pictures(self, res):
        """uint16 TBC fields of all located windows -> list of (window index, array linecount*outlinelen)."""
        rf, be = self.rf, self.rf._be
        if res.d_pic is None:
            return []
        be.synchronize()
        pic = be.to_host(res.d_pic).reshape(len(res.located), res.out_stride)
        st = be.to_host(res.d_status)
        W = rf.SysParams['outlinelen']
        out = []
        for j, i in enumerate(res.located):
            ok = (st[j] & (1 | 2 | 4 | 8)) == 0
            out.append((i, pic[j, :res.infos[i].linecount * W].copy() if ok else None))
        return out

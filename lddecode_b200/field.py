"""Field / FieldNTSC / FieldPAL: host-side mirrors of the reference's field classes
(lddecode_core.py:489-1191) on top of the CUDA kernels, plus the batched core they share with the
whole-capture pipeline (pipeline.py).

What runs where:
  sync-peak chase (get_syncpeaks)                         -> ldd_sync_peaks      (device)
  vsync detection / parity / line numbering (A7)           -> ldd_field_locate    (host C++, scalar)
  refine_linelocs_hsync                                    -> ldd_refine_hsync    (device)
  FieldNTSC.refine_linelocs_burst x2                       -> ldd_refine_burst    (device)
  FieldPAL.refine_linelocs_pilot                           -> ldd_refine_pilot    (device)
  downscale / scale / uint16 quantisation                  -> ldd_tbc_fields      (device)
  Philips code (VBI) decode, 3 lines per field             -> host numpy on three fetched line windows
"""
import ctypes as C

import numpy as np

from . import _lib
from .rfdecode import DeviceDemod

LL_STRIDE = 320          # >= linecount + 4 (PAL: 317)


def _h(x):
    return x.ctypes.data_as(C.c_void_p)


class FieldBatch:
    """Per-field tables of a batch of fields that share one set of device planes."""

    def __init__(self, rf, nfields):
        self.rf = rf
        self.n = nfields
        self.base = np.zeros(nfields, dtype=np.int64)        # plane index of each window's sample 0
        self.winlen = np.zeros(nfields, dtype=np.int64)
        self.linecount = np.zeros(nfields, dtype=np.int32)
        self.linelocs1 = np.zeros((nfields, LL_STRIDE), dtype=np.float64)
        self.linebad = np.zeros((nfields, LL_STRIDE), dtype=np.uint8)

    @classmethod
    def view(cls, rf, batch, idx, linecounts):
        """The rows `idx` of `batch` as a new batch (fancy-indexed copies, no Python loop per field)."""
        b = cls.__new__(cls)
        b.rf, b.n = rf, len(idx)
        b.base, b.winlen = batch.base[idx], batch.winlen[idx]
        b.linecount = np.ascontiguousarray(linecounts, dtype=np.int32)
        b.linelocs1, b.linebad = batch.linelocs1[idx], batch.linebad[idx]
        return b


class PendingPeaks:
    """An enqueued peak chase whose results are on their way to pinned host memory."""

    def __init__(self, be, ev, h_pk, h_vl, h_cnt, keep):
        self.be, self.ev, self.h_pk, self.h_vl, self.h_cnt, self.keep = be, ev, h_pk, h_vl, h_cnt, keep

    def result(self):
        be = self.be
        be.wait_event(self.ev)
        c = int(be.host_view(self.h_cnt)[0])
        return be.host_view(self.h_pk)[:c].copy(), be.host_view(self.h_vl)[:c].copy()


def sync_peaks_launch(rf, sync_buf, n, start=0, staging=None):
    """Enqueue Field.get_syncpeaks on a device float64 demod_sync plane and the device->host copy of its
    result (asynchronous); returns a PendingPeaks.  `staging`: dict reused across calls for the buffers."""
    be = rf._be
    cap = int(n // int(rf.linelen * .4)) + 8
    st = staging if staging is not None else {}
    if st.get('cap', 0) < cap:
        st['cap'] = cap
        st['pk'], st['vl'], st['cnt'] = be.empty(cap, np.int64), be.empty(cap, np.float64), be.zeros(2, np.int32)
        st['h_pk'], st['h_vl'], st['h_cnt'] = be.pinned(cap, np.int64), be.pinned(cap, np.float64), be.pinned(2, np.int32)
    pk, vl, cnt = st['pk'], st['vl'], st['cnt']
    rf._check(be.lib.ldd_sync_peaks(rf._h, be.ptr(sync_buf), int(n), int(start), be.ptr(pk), be.ptr(vl), st['cap'],
                                    be.ptr(cnt), be.stream()))
    # the list goes to pinned host memory through a kernel, not the copy engine (it would queue behind a field download)
    rf._check(be.lib.ldd_peaks_to_host(be.ptr(pk), be.ptr(vl), be.ptr(cnt), st['cap'], be.ptr(st['h_pk']), be.ptr(st['h_vl']),
                                       be.ptr(st['h_cnt']), be.stream()))
    return PendingPeaks(be, be.record_event(), st['h_pk'], st['h_vl'], st['h_cnt'], st)


def sync_peaks_device(rf, sync_buf, n, start=0, staging=None):
    """Field.get_syncpeaks on a device float64 demod_sync plane -> (peaks int64, values float64) on the host."""
    return sync_peaks_launch(rf, sync_buf, n, start, staging).result()


def sync_peaks_prefix_host(rf, sync_buf, n, staging):
    """The same chase for a short stretch of the plane, done on the HOST: the samples come over on a
    side stream (a copy engine, so the transfer does not queue behind kernels already enqueued on the
    main stream -- the whole-capture pipeline has the next capture's demodulation in flight at this
    point) and ldd_sync_peaks_host runs the comparisons.  The caller guarantees that sync_buf[:n] is
    complete (it has seen the result of a later operation on the producing stream)."""
    be = rf._be
    st = staging
    if st.get('hcap', 0) < n:
        st['hcap'] = int(n)
        st['h_sync'] = be.pinned(n, np.float64)
        st['side'] = st.get('side') or be.new_stream()
        cap = int(n // int(rf.linelen * .4)) + 8
        st['hpk'], st['hvl'] = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.float64)
    with be.stream_ctx(st['side']):
        be.copy_async(st['h_sync'][:n], sync_buf[:n])
        ev = be.record_event()
    be.wait_event(ev)
    cnt = C.c_int(0)
    hs = be.host_view(st['h_sync'])
    rf._check(be.lib.ldd_sync_peaks_host(rf._h, _h(hs), int(n), 0, _h(st['hpk']), _h(st['hvl']), len(st['hpk']), C.byref(cnt)))
    c = min(cnt.value, len(st['hpk']))
    return st['hpk'][:c].copy(), st['hvl'][:c].copy()


def locate(rf, peaks, vals, window_len, start=0):
    """ldd_field_locate for one window -> (FieldInfo, linelocs1, linebad)."""
    info = _lib.FieldInfo()
    peaks = np.ascontiguousarray(peaks, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    ll = np.zeros(LL_STRIDE, dtype=np.float64)
    bad = np.zeros(LL_STRIDE, dtype=np.uint8)
    rf._check(rf._be.lib.ldd_field_locate(rf._h, _h(peaks), _h(vals), len(peaks), int(window_len), int(start),
                                          C.byref(info), _h(ll), _h(bad), LL_STRIDE))
    return info, ll, bad


class RefinedBatch:
    pass


def refine_and_tbc(rf, planes, plane_len, batch, colorlevel=1.45, colorphase=91.5, want_intermediates=True, staging=None):
    """Device part of FieldNTSC/FieldPAL.__init__ for every located field of `batch`.

    planes: dict name -> device buffer (as in DeviceDemod.planes).  Returns a RefinedBatch with
    device buffers (dspicture uint16 [n][out_stride]) and, when asked, host copies of the line tables."""
    be, lib = rf._be, rf._be.lib
    n = batch.n
    out = RefinedBatch()
    W = rf.SysParams['outlinelen']
    maxlc = int(batch.linecount.max()) if n else 0
    out.out_stride = (rf.SysParams['frame_lines'] // 2 + 1) * W
    # one host->device copy for all per-field tables: [base | winlen | linelocs1 | linecount | linebad]
    # (through a pinned buffer of `staging` when given, so that the host does not wait for the stream)
    npk = n * (8 + 8 + 8 * LL_STRIDE + 4 + LL_STRIDE) + 64
    if staging is not None:
        if staging.get('tbl_cap', 0) < npk:
            staging['tbl_cap'] = npk + npk // 2
            staging['tbl'] = be.pinned(staging['tbl_cap'], np.uint8)
            staging['tbl_ev'] = None
        if staging['tbl_ev'] is not None:
            be.wait_event(staging['tbl_ev'])        # the previous upload from this buffer has been executed
        pk = be.host_view(staging['tbl'])[:npk]
    else:
        pk = np.empty(npk, dtype=np.uint8)
    o_base, o_win = 0, 8 * n
    o_l1 = 16 * n
    o_lc = o_l1 + 8 * n * LL_STRIDE
    o_bad = o_lc + 4 * n
    pk[o_base:o_base + 8 * n] = batch.base.view(np.uint8)
    pk[o_win:o_win + 8 * n] = batch.winlen.view(np.uint8)
    pk[o_l1:o_l1 + 8 * n * LL_STRIDE] = np.ascontiguousarray(batch.linelocs1).reshape(-1).view(np.uint8)
    pk[o_lc:o_lc + 4 * n] = batch.linecount.view(np.uint8)
    pk[o_bad:o_bad + n * LL_STRIDE] = np.ascontiguousarray(batch.linebad).reshape(-1)
    if staging is not None:
        d_pk = be.empty(npk, np.uint8)
        rf._check(be.lib.ldd_copy_small(be.ptr(d_pk), be.ptr(staging['tbl']), npk, be.stream()))     # not the copy engine: see ldd_b200.h
        staging['tbl_ev'] = be.record_event()
    else:
        d_pk = be.to_device(pk)
    d_base = d_pk[o_base:o_base + 8 * n].view(be.dtype_of(np.int64))
    d_win = d_pk[o_win:o_win + 8 * n].view(be.dtype_of(np.int64))
    d_l1 = d_pk[o_l1:o_l1 + 8 * n * LL_STRIDE].view(be.dtype_of(np.float64))
    d_lc = d_pk[o_lc:o_lc + 4 * n].view(be.dtype_of(np.int32))
    d_bad = d_pk[o_bad:o_bad + n * LL_STRIDE]
    d_l2 = be.empty(n * LL_STRIDE, np.float64)
    d_bad2 = be.empty(n * LL_STRIDE, np.uint8)
    d_status = be.zeros(n, np.int32)
    st = be.stream()
    rf._check(lib.ldd_refine_hsync(rf._h, be.ptr(planes['demod_05']), int(plane_len), be.ptr(d_base), be.ptr(d_win),
                                   be.ptr(d_lc), n, LL_STRIDE, be.ptr(d_l1), be.ptr(d_bad), be.ptr(d_l2), be.ptr(d_bad2),
                                   be.ptr(d_status), st))
    out.d_linelocs2, out.d_linebad = d_l2, d_bad2
    d_pic = be.empty(n * out.out_stride, np.uint16)
    ire0 = float(rf.SysParams['ire0'])
    if rf.system == 'NTSC':
        d_l3 = be.empty(n * LL_STRIDE, np.float64)
        d_l4 = be.empty(n * LL_STRIDE, np.float64)
        d_bl = be.empty(n * LL_STRIDE, np.float32)
        for src, dst in ((d_l2, d_l3), (d_l3, d_l4)):
            rf._check(lib.ldd_refine_burst(rf._h, be.ptr(planes['demod_burst']), int(plane_len), be.ptr(d_base), be.ptr(d_lc),
                                           n, LL_STRIDE, be.ptr(src), be.ptr(dst), be.ptr(d_bl), be.ptr(d_status), st))
        # apply_offsets(linelocs4, shift33 - 8) (lddecode_core.py:1161-1162, 1185-1186)
        shift33 = colorphase * (np.pi / 180)
        out.lineloc_add = (shift33 - 8) * (rf.freq / (4 * 315 / 88))
        rf._check(lib.ldd_tbc_fields(rf._h, be.ptr(planes['demod']), int(plane_len), ire0, be.ptr(d_base), be.ptr(d_l4),
                                     LL_STRIDE, be.ptr(d_lc), n, maxlc, 1, out.lineloc_add, W, 1, 1, be.ptr(d_pic),
                                     out.out_stride, be.ptr(d_bl), float(colorlevel), be.ptr(d_status), st))
        out.d_linelocs3, out.d_linelocs4, out.d_burstlevel, out.d_final = d_l3, d_l4, d_bl, d_l4
    else:
        d_lp = be.empty(n * LL_STRIDE, np.float64)
        rf._check(lib.ldd_refine_pilot(rf._h, be.ptr(planes['demod']), be.ptr(planes['demod_05']), int(plane_len),
                                       be.ptr(d_base), be.ptr(d_lc), n, LL_STRIDE, be.ptr(d_l2), be.ptr(d_lp),
                                       be.ptr(d_status), st))
        out.lineloc_add = 0.0
        rf._check(lib.ldd_tbc_fields(rf._h, be.ptr(planes['demod']), int(plane_len), ire0, be.ptr(d_base), be.ptr(d_lp),
                                     LL_STRIDE, be.ptr(d_lc), n, maxlc, 3, 0.0, W, 1, 1, be.ptr(d_pic), out.out_stride,
                                     None, float(colorlevel), be.ptr(d_status), st))
        out.d_final = d_lp
    out.d_pic, out.d_status, out.d_base, out.d_lc = d_pic, d_status, d_base, d_lc
    if want_intermediates:
        be.synchronize()
        out.status = be.to_host(d_status)
        need = ((out.status & _lib.ST_LINE_LONG) != 0) & ((out.status & _lib.ST_LINE_BAD) == 0)
        if need.any():
            # lines longer than the staging window of the TBC pass (the reference's scale() takes any span,
            # lddutils.py:83-97): second pass with the exact kernel, in place
            d_st2 = be.zeros(n, np.int32)
            ntsc = rf.system == 'NTSC'
            rf._check(lib.ldd_tbc_long_lines(rf._h, be.ptr(planes['demod']), int(plane_len), ire0, be.ptr(d_base), be.ptr(out.d_final),
                                             LL_STRIDE, be.ptr(d_lc), n, maxlc, 1 if ntsc else 3, out.lineloc_add, W, 1, 1,
                                             be.ptr(d_pic), out.out_stride, None, 0, be.ptr(out.d_burstlevel) if ntsc else None,
                                             float(colorlevel), be.ptr(d_st2), st))
            be.synchronize()
            done = need & (be.to_host(d_st2) == 0)
            out.status = np.array(out.status)
            out.status[done] &= ~(1 | _lib.ST_LINE_LONG)
        out.linelocs2 = be.to_host(d_l2).reshape(n, LL_STRIDE)
        out.linebad = be.to_host(d_bad2).reshape(n, LL_STRIDE)
        out.final = be.to_host(out.d_final).reshape(n, LL_STRIDE)
        if rf.system == 'NTSC':
            out.linelocs3 = be.to_host(d_l3).reshape(n, LL_STRIDE)
            out.linelocs4 = be.to_host(d_l4).reshape(n, LL_STRIDE)
            out.burstlevel = be.to_host(d_bl).reshape(n, LL_STRIDE)
    return out


# ---- VBI (lddecode_core.py:814-884): the cell walk runs on the device (ldd_vbi_decode), the code interpretation here
def code_nibbles(code):
    """24-bit Philips code -> the reference's `linecode` list of six nibbles (None for -1)."""
    code = int(code)
    if code < 0:
        return None
    return [(code >> s) & 15 for s in (20, 16, 12, 8, 4, 0)]


def vbi_decode_device(rf, demod_buf, n, d_linelocs, nfields, d_base=None, d_winlen=None):
    """ldd_vbi_decode for `nfields` line tables [nfields][LL_STRIDE] -> device int32 [nfields][4] (enqueued)."""
    be = rf._be
    lines = rf.SysParams['philips_codelines']
    arr = (C.c_int * len(lines))(*lines)
    d_codes = be.empty(4 * nfields, np.int32)
    rf._check(be.lib.ldd_vbi_decode(rf._h, be.ptr(demod_buf), int(n), be.ptr(d_base) if d_base is not None else None,
                                    be.ptr(d_winlen) if d_winlen is not None else None, be.ptr(d_linelocs), LL_STRIDE,
                                    int(nfields), arr, len(lines), be.ptr(d_codes), be.stream()))
    return d_codes


def process_philips(rf, linecode):
    """processphilipscode (lddecode_core.py:836-884)."""
    vbi = {'minutes': None, 'seconds': None, 'clvframe': None, 'framenr': None, 'statuscode': None, 'status': None,
           'isclv': False}
    for l in rf.SysParams['philips_codelines']:
        lc = linecode.get(l)
        if lc is None:
            continue
        if lc[0] == 15 and lc[2] == 13:
            vbi['minutes'] = 60 * lc[1] + lc[4] * 10 + lc[5]
            vbi['isclv'] = True
        elif lc[0] == 15:
            vbi['framenr'] = (lc[1] & 7) * 10000 + lc[2] * 1000 + lc[3] * 100 + lc[4] * 10 + lc[5]
        else:
            h = (lc[0] << 20) | (lc[1] << 16) | (lc[2] << 12) | (lc[3] << 8) | (lc[4] << 4) | lc[5]
            if lc[2] == 0xE:
                vbi['seconds'] = (lc[1] - 10) * 10 + lc[3]
                vbi['clvframe'] = lc[4] * 10 + lc[5]
                vbi['isclv'] = True
            htop = h >> 12
            if htop == 0x8dc or htop == 0x8ba:
                vbi['status'] = h
            if h == 0x87ffff:
                vbi['isclv'] = True
    return vbi


def downscale_audio(audio, lineinfo, rf, linecount, timeoffset=0, freq=48000.0, scale=64):
    """lddecode_core.downscale_audio (lddecode_core.py:431-484) on the device (ldd_downscale_audio).

    audio: the phase-2 audio of the field's window -- a dict of device buffers (DeviceDemod.audio) or the host record
    array RFDecode.demod returns.  Returns (int16 interleaved L/R samples, time offset to carry into the next field).
    Like the reference it assumes a total decimation of `scale` (64: right from 32 MSPS up, twice the true factor at
    8fsc NTSC, where the reference therefore plays the audio at half speed); scale=None uses the decoder's true factor."""
    be = rf._be
    if scale is None:
        scale = rf.audio_decimation
    frametime = (rf.SysParams['line_period'] * linecount) / 1000000
    soundgap = 1 / freq
    arange = np.arange(timeoffset, frametime + soundgap, soundgap, dtype=np.double)
    n = len(arange) - 1
    if n < 1:
        return np.zeros(0, dtype=np.int16), (arange[-1] - frametime if len(arange) else timeoffset)
    if isinstance(audio, dict):
        al, ar = audio['audio_left'], audio['audio_right']
    else:
        al = be.to_device(np.ascontiguousarray(audio['audio_left'], dtype=np.float64))
        ar = be.to_device(np.ascontiguousarray(audio['audio_right'], dtype=np.float64))
    ll = np.zeros(LL_STRIDE, dtype=np.float64)
    nll = min(len(lineinfo), LL_STRIDE)
    ll[:nll] = np.asarray(lineinfo, dtype=np.float64)[:nll]
    out = be.empty(2 * n, np.int16)
    st = be.zeros(1, np.int32)
    SP = rf.SysParams
    # the small per-field arguments (kept referenced until the call has been enqueued)
    d = [be.to_device(x) for x in (ll, np.array([nll], dtype=np.int32), arange[0:1].copy(), arange[1:2].copy(),
                                   np.array([n], dtype=np.int32), np.zeros(1, dtype=np.int64))]
    rf._check(be.lib.ldd_downscale_audio(
        rf._h, be.ptr(al), be.ptr(ar), len(al), None, be.ptr(d[0]), LL_STRIDE, be.ptr(d[1]), be.ptr(d[2]), be.ptr(d[3]),
        be.ptr(d[4]), be.ptr(d[5]), 1, n, 0.0, float(scale), float(SP['line_period']), float(SP['audio_lfreq']),
        float(SP['audio_rfreq']), be.ptr(out), be.ptr(st), be.stream()))
    be.synchronize()
    if be.to_host(st)[0] & 16:
        raise IndexError("downscale_audio: index outside the line table / audio array")
    return be.to_host(out), arange[-1] - frametime


# ---- drop-in classes -----------------------------------------------------------------------------
class Field:
    """Field(rf, rawdecode, start, audio_offset=0): rawdecode is what RFDecode.demod_device returned
    (device resident, preferred) or the (video, audio) record arrays of RFDecode.demod."""

    full = False

    def __init__(self, rf, rawdecode, start, audio_offset=0, keepraw=True, colorlevel=1.45, colorphase=91.5):
        if rawdecode is None:
            return
        self.rf = rf
        self.data = rawdecode
        self.start = start
        self.inlinelen = rf.linelen
        self.outlinelen = rf.SysParams['outlinelen']
        self.valid = False
        self.dspicture = None
        self.dsaudio = None
        self.audio_rec = None
        self.audio_next_offset = audio_offset
        self.colorlevel, self.colorphase = colorlevel, colorphase
        self.burstlevel = None
        self._planes, self._n = self._device_planes(rawdecode)
        # phase-2 audio of the window: device buffers (DeviceDemod) or the host record array of RFDecode.demod
        self._audio = rawdecode.audio if isinstance(rawdecode, DeviceDemod) else rawdecode[1]
        be = rf._be

        self.peaklist, self._peakvals = self.get_syncpeaks(with_values=True)
        info, ll1, bad = locate(rf, self.peaklist, self._peakvals, self._n, start)
        self._info = info
        self.vsyncs = [[info.vsyncs[i][q] for q in range(3)] for i in range(min(info.nvsyncs, 4))]
        if info.stage == _lib.FIELD_CRASH:
            raise TypeError("cannot unpack non-iterable NoneType object")      # what the reference does here
        self.nextfieldoffset = int(info.nextfieldoffset)
        if len(self.peaklist) >= 200:          # determine_vsyncs calls get_hsync_median from there on (lddecode_core.py:595-598)
            self.med_hsync, self.hsync_tolerance = info.med_hsync, info.hsync_tolerance
        if info.stage in (_lib.FIELD_NOVSYNC, _lib.FIELD_SHORT):
            return
        self.istop = info.istop
        self.linecount = info.linecount
        if info.stage == _lib.FIELD_BADLINES:
            print('unable to decode frame')
            return
        nll = self.linecount + 4
        self.linelocs1 = list(ll1[:nll])
        batch = FieldBatch(rf, 1)
        batch.linecount[0] = self.linecount
        batch.winlen[0] = self._n
        batch.linelocs1[0] = ll1
        batch.linebad[0] = bad
        self._batch = batch
        ref = refine_and_tbc(rf, self._planes, self._n, batch, colorlevel, colorphase) if self.full else \
            self._refine_hsync_only(batch)
        self.linebad = [bool(x) for x in ref.linebad[0][:nll]]
        self.linelocs2 = list(ref.linelocs2[0][:nll])
        if ref.status[0] & 2:
            print('unable to decode frame')
            return
        self.linelocs = self.linelocs2
        self._decode_vbi()
        self.valid = True
        self.tbcstart = int(info.tbcstart)
        if self.full:
            self._finish(ref, nll)

    # -- plumbing
    def _device_planes(self, rawdecode):
        rf, be = self.rf, self.rf._be
        if isinstance(rawdecode, DeviceDemod):
            return rawdecode.planes, rawdecode.length
        video = rawdecode[0]
        ire0 = rf.SysParams['ire0']
        planes = {}
        for name in video.dtype.names:
            v = np.asarray(video[name], dtype=np.float64)
            if name == 'demod_sync':
                planes[name] = be.to_device(v)
            else:
                planes[name] = be.to_device((v - ire0 if name in ('demod', 'demod_05') else v).astype(np.float32))
        return planes, len(video)

    def _refine_hsync_only(self, batch):
        rf, be, lib = self.rf, self.rf._be, self.rf._be.lib
        out = RefinedBatch()
        d = {k: be.to_device(getattr(batch, k).reshape(-1)) for k in ('base', 'winlen', 'linecount', 'linelocs1', 'linebad')}
        d_l2 = be.empty(LL_STRIDE, np.float64)
        d_bad2 = be.empty(LL_STRIDE, np.uint8)
        d_status = be.zeros(1, np.int32)
        rf._check(lib.ldd_refine_hsync(rf._h, be.ptr(self._planes['demod_05']), int(self._n), be.ptr(d['base']),
                                       be.ptr(d['winlen']), be.ptr(d['linecount']), 1, LL_STRIDE, be.ptr(d['linelocs1']),
                                       be.ptr(d['linebad']), be.ptr(d_l2), be.ptr(d_bad2), be.ptr(d_status), be.stream()))
        be.synchronize()
        out.status = be.to_host(d_status)
        out.linelocs2 = be.to_host(d_l2).reshape(1, LL_STRIDE)
        out.linebad = be.to_host(d_bad2).reshape(1, LL_STRIDE)
        return out

    def _plane_slice_host(self, name, a, b):
        """Host float64 copy of plane[a:b] in the reference's units."""
        be = self.rf._be
        a, b = max(int(a), 0), min(int(b), self._n)
        v = be.to_host(self._planes[name][a:b]).astype(np.float64)
        if name in ('demod', 'demod_05'):
            v += self.rf.SysParams['ire0']
        return v

    def _decode_vbi(self):
        rf, be = self.rf, self.rf._be
        self.isclv = False
        self.framenr = None
        ll = np.zeros(LL_STRIDE, dtype=np.float64)
        ll[:len(self.linelocs)] = self.linelocs
        d_ll = be.to_device(ll)
        codes = be.to_host(vbi_decode_device(rf, self._planes['demod'], self._n, d_ll, 1))
        self.linecode = {l: code_nibbles(codes[i]) for i, l in enumerate(rf.SysParams['philips_codelines'])}
        self.vbi = process_philips(rf, self.linecode)

    # -- reference API
    def usectoinpx(self, x):
        return x * self.rf.freq

    def inpxtousec(self, x):
        return x / self.rf.freq

    def get_syncpeaks(self, with_values=False):
        pk, vl = sync_peaks_device(self.rf, self._planes['demod_sync'], self._n, self.start)
        return (list(pk), vl) if with_values else list(pk)

    # -- the reference's per-step methods (lddecode_core.py:518-787, 814-884), for callers that walk through a field by
    # hand as the reference's notebooks do.  Each one re-runs its step through the library on this field's data.
    def _locate(self):
        return locate(self.rf, self.peaklist, self._peakvals, self._n, self.start)

    def get_hsync_median(self):
        info, _, _ = self._locate()
        self.med_hsync, self.hsync_tolerance = info.med_hsync, info.hsync_tolerance
        return self.med_hsync, self.hsync_tolerance

    def is_regular_hsync(self, peaknum):
        if peaknum >= len(self.peaklist) or self.peaklist[peaknum] > self._n:
            return False
        return bool(self.med_hsync - self.hsync_tolerance <= self._peakvals[peaknum] <= self.med_hsync + self.hsync_tolerance)

    def determine_field(self, peaknum):
        if peaknum < 11:
            return None
        pk = np.ascontiguousarray(self.peaklist, dtype=np.int64)
        vl = np.ascontiguousarray(self._peakvals, dtype=np.float64)
        line0, vote = C.c_int(-1), C.c_int(0)
        self.rf._check(self.rf._be.lib.ldd_field_vote(self.rf._h, _h(pk), _h(vl), len(pk), int(self._n), float(self.med_hsync),
                                                      float(self.hsync_tolerance), int(peaknum), C.byref(line0), C.byref(vote)))
        return (line0.value if line0.value >= 0 else None), vote.value

    def determine_vsyncs(self):
        info, _, _ = self._locate()
        if len(self.peaklist) >= 200:
            self.med_hsync, self.hsync_tolerance = info.med_hsync, info.hsync_tolerance
        return [(info.vsyncs[i][0], info.vsyncs[i][1], bool(info.vsyncs[i][2])) for i in range(min(info.nvsyncs, 4))]

    def compute_linelocs(self):
        info, ll1, bad = self._locate()
        if info.stage != _lib.FIELD_LOCATED:
            raise ValueError("compute_linelocs: the field's line grid cannot be built (stage %d)" % info.stage)
        nll = info.linecount + 4
        return list(ll1[:nll]), [bool(x) for x in bad[:nll]]

    def refine_linelocs_hsync(self):
        """lddecode_core.py:715-787 on self.linelocs1 / self.linebad; like the reference it updates self.linebad."""
        nll = self.linecount + 4
        batch = FieldBatch(self.rf, 1)
        batch.linecount[0], batch.winlen[0] = self.linecount, self._n
        batch.linelocs1[0][:nll] = self.linelocs1
        batch.linebad[0][:nll] = np.asarray(self.linebad, dtype=np.uint8)
        ref = self._refine_hsync_only(batch)
        self.linebad = [bool(x) for x in ref.linebad[0][:nll]]
        return list(ref.linelocs2[0][:nll])

    def decodephillipscode(self, linenum):
        """lddecode_core.py:814-834 for any line of self.linelocs: list of six nibbles or None."""
        rf, be = self.rf, self.rf._be
        ll = np.zeros(LL_STRIDE, dtype=np.float64)
        ll[:len(self.linelocs)] = self.linelocs
        d_codes = be.empty(4, np.int32)
        arr = (C.c_int * 1)(int(linenum))
        rf._check(be.lib.ldd_vbi_decode(rf._h, be.ptr(self._planes['demod']), int(self._n), None, None, be.ptr(be.to_device(ll)),
                                        LL_STRIDE, 1, arr, 1, be.ptr(d_codes), be.stream()))
        be.synchronize()
        return code_nibbles(be.to_host(d_codes)[0])

    def processphilipscode(self):
        self.vbi = process_philips(self.rf, self.linecode)

    def _line_tables_device(self, lineinfo):
        be = self.rf._be
        ll = np.zeros(LL_STRIDE, dtype=np.float64)
        ll[:len(lineinfo)] = lineinfo
        return (be.to_device(ll), be.to_device(np.zeros(1, dtype=np.int64)), be.to_device(np.array([self.linecount], dtype=np.int32)),
                be.zeros(1, np.int32))

    def downscale(self, lineoffset=1, lineinfo=None, outwidth=None, wow=True, channel='demod', audio=False):
        """Field.downscale (lddecode_core.py:789-812): float64 Hz, linecount * outwidth samples."""
        rf, be = self.rf, self.rf._be
        if lineinfo is None:
            lineinfo = self.linelocs
        if outwidth is None:
            outwidth = self.outlinelen
        ll = np.zeros(LL_STRIDE, dtype=np.float64)
        ll[:len(lineinfo)] = lineinfo
        d_ll = be.to_device(ll)
        d_lc = be.to_device(np.array([self.linecount], dtype=np.int32))
        d_st = be.zeros(1, np.int32)
        out = be.empty(self.linecount * outwidth, np.float64)
        add = float(rf.SysParams['ire0']) if channel in ('demod', 'demod_05') else 0.0
        rf._check(be.lib.ldd_tbc_fields(rf._h, be.ptr(self._planes[channel]), int(self._n), add, None, be.ptr(d_ll), LL_STRIDE,
                                        be.ptr(d_lc), 1, self.linecount, int(lineoffset), 0.0, int(outwidth), int(bool(wow)), 0,
                                        be.ptr(out), self.linecount * outwidth, None, 1.0, be.ptr(d_st), be.stream()))
        be.synchronize()
        if be.to_host(d_st)[0] & 1:
            raise ValueError("line window outside the decoded data")
        if audio:
            self._downscale_audio(lineinfo)
        return be.to_host(out), self.dsaudio

    def _downscale_audio(self, lineinfo):
        """The audio half of Field.downscale (lddecode_core.py:809-810)."""
        if self.rf.decode_analog_audio and self._audio is not None:
            self.dsaudio, self.audio_next_offset = downscale_audio(self._audio, lineinfo, self.rf, self.linecount,
                                                                   self.audio_next_offset)

    def _finish(self, ref, nll):
        pass


class FieldNTSC(Field):
    full = True

    def _finish(self, ref, nll):
        be = self.rf._be
        if ref.status[0] & (1 | 4):
            print("ERROR: Unable to decode frame, skipping")
            self.valid = False
            return
        self.linelocs3 = ref.linelocs3[0][:nll].copy()
        self.linelocs4 = ref.linelocs4[0][:nll].copy()
        self.burstlevel = ref.burstlevel[0][:nll].copy()
        self.linelocs = self.linelocs4 + ref.lineloc_add
        W = self.outlinelen
        self.dspicture = be.to_host(ref.d_pic)[:self.linecount * W].copy()
        try:
            self._downscale_audio(self.linelocs)          # downscale(final=True) -> audio=True (lddecode_core.py:1136, 1188)
        except IndexError:
            print("ERROR: Unable to decode frame, skipping")
            self.valid = False

    def downscale(self, lineoffset=1, final=False, *args, **kwargs):
        """FieldNTSC.downscale (lddecode_core.py:1135-1159); final=True returns the uint16 picture of the constructor."""
        if final:
            return self.dspicture, self.dsaudio
        return super().downscale(lineoffset=lineoffset, *args, **kwargs)

    def apply_offsets(self, linelocs, phaseoffset, picoffset=0):
        return np.array(linelocs) + picoffset + (phaseoffset * (self.rf.freq / (4 * 315 / 88)))

    def refine_linelocs_burst(self, linelocs2):
        """lddecode_core.py:1054-1133: one burst-phase pass over the given line table -> (linelocs, burstlevel)."""
        rf, be = self.rf, self.rf._be
        nll = self.linecount + 4
        d_ll, d_base, d_lc, d_st = self._line_tables_device(linelocs2)
        d_out, d_bl = be.empty(LL_STRIDE, np.float64), be.zeros(LL_STRIDE, np.float32)
        rf._check(be.lib.ldd_refine_burst(rf._h, be.ptr(self._planes['demod_burst']), int(self._n), be.ptr(d_base), be.ptr(d_lc),
                                          1, LL_STRIDE, be.ptr(d_ll), be.ptr(d_out), be.ptr(d_bl), be.ptr(d_st), be.stream()))
        be.synchronize()
        if be.to_host(d_st)[0] & (1 | 4):
            raise IndexError("refine_linelocs_burst: a burst window lies outside the decoded data")
        return be.to_host(d_out)[:nll].copy(), be.to_host(d_bl)[:nll].copy()


class FieldPAL(Field):
    full = True

    def refine_linelocs_pilot(self, linelocs=None):
        """lddecode_core.py:962-1021: pilot-phase alignment of a line table (default: self.linelocs2)."""
        rf, be = self.rf, self.rf._be
        nll = self.linecount + 4
        d_ll, d_base, d_lc, d_st = self._line_tables_device(self.linelocs2 if linelocs is None else linelocs)
        d_out = be.empty(LL_STRIDE, np.float64)
        rf._check(be.lib.ldd_refine_pilot(rf._h, be.ptr(self._planes['demod']), be.ptr(self._planes['demod_05']), int(self._n),
                                          be.ptr(d_base), be.ptr(d_lc), 1, LL_STRIDE, be.ptr(d_ll), be.ptr(d_out), be.ptr(d_st),
                                          be.stream()))
        be.synchronize()
        if be.to_host(d_st)[0] & (1 | 8):
            raise IndexError("refine_linelocs_pilot: a pilot window lies outside the decoded data")
        return be.to_host(d_out)[:nll].copy()

    def downscale(self, final=False, *args, **kwargs):
        """FieldPAL.downscale (lddecode_core.py:1023-1035): lineoffset 3; final=True returns the constructor's picture."""
        if final:
            return self.dspicture, self.dsaudio
        kwargs.setdefault('lineoffset', 3)
        return super().downscale(*args, **kwargs)

    def _finish(self, ref, nll):
        be = self.rf._be
        if ref.status[0] & (1 | 8):
            print("ERROR: Unable to decode frame, skipping")
            self.valid = False
            return
        self.linelocs = ref.final[0][:nll].copy()
        W = self.outlinelen
        self.dspicture = be.to_host(ref.d_pic)[:self.linecount * W].copy()
        try:
            self._downscale_audio(self.linelocs)          # downscale(final=True) -> audio=True (lddecode_core.py:1024, 1045)
        except IndexError:
            print("ERROR: Unable to decode frame, skipping")
            self.valid = False

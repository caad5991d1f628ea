"""Whole-capture decode with everything resident in HBM: the throughput path behind bench.py.

The reference walks a capture field by field (Framer.readfield, lddecode_core.py:1194-1223): every
field re-reads and re-demodulates 1e6 samples (2.1x redundant) and the next read position comes out
of the previous field.  Here a range of the capture is demodulated ONCE on the fixed global block
grid (blocks start at multiples of blocklen-1056; they are independent given the 1024/32-sample
halos), one sync-peak chase covers the whole range, the field-to-field walk
runs on the host over the ~16 K peaks per second of video (ldd_field_chain), and all located fields
are refined and resampled in a handful of batched launches.  Only the peak list crosses to the host
in the middle.

A step is two calls into the library (include/ldd_b200.h, ldd_pipe_launch / ldd_pipe_finish): the
first enqueues demodulation + peak chase (+ audio), the second waits for the peak list, walks, and
enqueues refinement + VBI + TBC from preallocated tables.  This module only owns the buffers (torch
tensors) and hands their addresses over; no per-field Python runs in the steady state.

Ranges make the same code serve a single GPU (one range), captures larger than HBM (ranges one
after the other) and multi-GPU sharding (one range per rank, parallel.py): a range owns the
fields whose read position lies in [r0, r1); it starts walking 1.6 fields early so that its read
positions are those of the sequential walk by the time it reaches r0, and it demodulates one read
length beyond r1 so every owned field completes locally.  Because all ranges use the same global
block grid, a sharded decode is bit-identical to the single-range decode.

Differences from calling Field() per window, by construction: planes come from one block grid
instead of one grid per field, so demod_sync differs by the FPsync wrap term (< 3e-5 absolute);
peak indices, line tables and TBC output are compared against the reference flow in tests/.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from . import field as F

READLEN = 1000000        # Framer.readlen (lddecode_core.py:1319, 1324)
LL = F.LL_STRIDE

# numpy view of the library's ldd_field records
FIELD_DTYPE = np.dtype([('stage', 'i4'), ('istop', 'i4'), ('linecount', 'i4'), ('npeaks', 'i4'), ('nvsyncs', 'i4'),
                        ('vsyncs', 'i4', (4, 3)), ('_pad', 'i4'), ('nextfieldoffset', 'i8'), ('tbcstart', 'i8'),
                        ('med_hsync', 'f8'), ('hsync_tolerance', 'f8')])
assert FIELD_DTYPE.itemsize == C.sizeof(_lib.FieldInfo)


def _host_array(ptr, n, ctype, dtype):
    """numpy copy of n elements at a ctypes pointer (host memory owned by the library)."""
    if n <= 0 or not bool(ptr):
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((ctype * n).from_address(C.addressof(ptr.contents)), dtype=dtype).copy()


class Refined:
    """Line tables of the located fields of a range: device views, fetched to the host on first access."""

    def __init__(self, res):
        self._res = res
        self.lineloc_add = res.lineloc_add
        self.out_stride = res.out_stride
        self.d_pic, self.d_status = res.d_pic, res.d_status
        self._cache = {}

    def _fetch(self, name, ptr, dtype):
        if name not in self._cache:
            res = self._res
            res._live()
            slot, n = res.slot, len(res.located)
            be = slot.be
            be.synchronize()
            off = ptr - slot.tables_addr
            nbytes = n * LL * np.dtype(dtype).itemsize
            self._cache[name] = be.to_host(slot.d_tables[off:off + nbytes]).view(dtype).reshape(n, LL).copy()
        return self._cache[name]

    @property
    def status(self):
        be = self._res.slot.be
        be.synchronize()
        return be.to_host(self.d_status)[:len(self._res.located)]

    linelocs2 = property(lambda s: s._fetch('l2', s._res.pr.d_linelocs2, np.float64))
    linebad = property(lambda s: s._fetch('bad2', s._res.pr.d_linebad2, np.uint8))
    linelocs3 = property(lambda s: s._fetch('l3', s._res.pr.d_linelocs3, np.float64))
    linelocs4 = property(lambda s: s._fetch('l4', s._res.pr.d_linelocs4, np.float64))
    burstlevel = property(lambda s: s._fetch('bl', s._res.pr.d_burstlevel, np.float32))
    final = property(lambda s: s._fetch('final', s._res.pr.d_final, np.float64))


class RangeResult:
    """Fields owned by one range.  Device buffers: d_pic uint16 [nlocated][out_stride], d_status int32.

    Host tables (infos, readsamples, base, linelocs1, gpeaks) are copied out of the library's arrays when the
    result is built; device-side line tables come over on first access of `refined.*`."""

    # -- per owned window
    @property
    def infos(self):
        return self._infos

    def _live(self):
        if self.slot.generation != self._generation:
            raise RuntimeError("this RangeResult's workspace has been reused by a later decode")

    @property
    def gpeaks(self):
        if self._gpeaks is None:
            self._live()
            self._gpeaks = _host_array(self.pr.gpeaks, self._npeaks, C.c_longlong, np.int64)
        return self._gpeaks

    @property
    def linelocs1(self):
        if self._l1 is None:
            self._live()
            allrows = _host_array(self.pr.linelocs1, self._nwin_all * LL, C.c_double, np.float64).reshape(-1, LL)
            self._l1 = allrows[self._owned]
        return self._l1

    def vbi_codes(self):
        """int32 [nlocated][3]: the 24-bit Philips codes of the three code lines (-1 = None), decoded on the device."""
        slot, n = self.slot, len(self.located)
        be = slot.be
        be.synchronize()
        off = self.pr.d_vbi - slot.tables_addr
        return be.to_host(slot.d_tables[off:off + 16 * n]).view(np.int32).reshape(n, 4)[:, :3].copy()


class PipeSlot:
    """One plane workspace of a CaptureDecoder plus the library's ldd_pipe on it."""

    def __init__(self, cd, rf, plane_cap, max_fields):
        be = self.be = rf._be
        self.rf, self.plane_cap, self.max_fields = rf, int(plane_cap), int(max_fields)
        lib = be.lib
        from .rfdecode import VIDEO_FIELDS, _PLANE_OF
        self.planes = {n: be.empty(self.plane_cap, np.float64 if n == 'demod_sync' else np.float32) for n in VIDEO_FIELDS[rf.system]}
        b = _lib.PipeBufs()
        for n, t in self.planes.items():
            b.planes[_PLANE_OF[n]] = be.ptr(t).value
        b.plane_cap = self.plane_cap
        self.a1 = self.a2 = None
        if rf.decode_analog_audio:
            ds = rf.blocklen // len(rf.Filters['audio_lfilt'])
            acap = self.plane_cap // ds + 16
            self.a1 = (be.empty(acap, np.float64), be.empty(acap, np.float64))
            self.a2 = (be.empty(acap // 4 + 16, np.float64), be.empty(acap // 4 + 16, np.float64))
            b.audio1_l, b.audio1_r, b.audio1_cap = be.ptr(self.a1[0]), be.ptr(self.a1[1]), acap
            b.audio2_l, b.audio2_r = be.ptr(self.a2[0]), be.ptr(self.a2[1])
        cap = int(self.plane_cap // int(rf.linelen * .4)) + 64
        self.pk = (be.empty(cap, np.int64), be.empty(cap, np.float64), be.zeros(2, np.int32))
        self.h_pk = (be.pinned(cap, np.int64), be.pinned(cap, np.float64), be.pinned(2, np.int32))
        b.peaks, b.peak_vals, b.peak_count, b.peak_cap = be.ptr(self.pk[0]), be.ptr(self.pk[1]), be.ptr(self.pk[2]), cap
        b.h_peaks, b.h_peak_vals, b.h_peak_count = be.ptr(self.h_pk[0]), be.ptr(self.h_pk[1]), be.ptr(self.h_pk[2])
        up, dev = C.c_longlong(0), C.c_longlong(0)
        rf._check(lib.ldd_pipe_table_bytes(self.max_fields, C.byref(up), C.byref(dev)))
        self.d_tables = be.empty(dev.value, np.uint8)
        self.h_tables = be.pinned(up.value, np.uint8)
        self.tables_addr = be.ptr(self.d_tables).value
        b.field_tables, b.tables_bytes = self.tables_addr, dev.value
        b.h_tables, b.h_tables_bytes = be.ptr(self.h_tables), up.value
        pre = 40 * rf.linelen + 64
        self.h_prefix = be.pinned(pre, np.float64)
        b.h_prefix, b.prefix_cap = be.ptr(self.h_prefix), pre
        self.bufs = b
        self.out_stride = (rf.SysParams['frame_lines'] // 2 + 1) * rf.SysParams['outlinelen']
        self.d_pic = be.empty(self.max_fields * self.out_stride, np.uint16)
        self.d_status = be.zeros(self.max_fields, np.int32)
        h = C.c_void_p()
        rf._check(lib.ldd_pipe_create(rf._h, C.byref(b), self.max_fields, int(cd.field_samples), C.byref(h)))
        self.h = h
        self.result = _lib.PipeResult()
        self.generation = 0

    def close(self):
        if self.h is not None:
            self.be.lib.ldd_pipe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def split_pcm(pcm, off, status):
    """Per-field views of a range's PCM: None where the field was dropped (empty span) or an index left the tables."""
    return [pcm[off[k]:off[k + 1]] if off[k + 1] > off[k] and not (status[k] & 16) else None for k in range(len(off) - 1)]


class Pending:
    """A launched range: ldd_pipe_launch has been enqueued on `slot`."""

    def __init__(self, slot, rf, r0, r1, ncap_total):
        self.slot, self.rf, self.r0, self.r1, self.ncap_total = slot, rf, r0, r1, ncap_total


class HostStreamDecoder:
    """A sequence of host-resident captures (the chunks a file reader produces) through one
    CaptureDecoder, software-pipelined: while the host walks the fields of chunk k the GPU already
    demodulates chunk k+1 (same stream, second plane workspace), and the upload of chunk k+2 and the
    download of chunk k-1's fields run on their own streams.  Device input, planes and pinned output
    are double-buffered.  Results are those of CaptureDecoder.decode() of every chunk.

        for res, pics in sd.run(chunks): ...                    # or, step by step:
        t = sd.upload(buf, n); p = sd.launch(t); job = sd.finish(p); res, pics = sd.fetch(job)

    upload() takes a pinned host buffer (be.pinned) holding the chunk's bytes in the capture's own format
    (u8, s16/u16, packed .r30 / .lds: the unpack is fused into the demodulation's block load).  fetch()
    returns the RangeResult and a host view [nfields, out_stride] of its uint16 fields that stays valid
    until the second-next finish(); the RangeResult's device planes stay valid until the second-next
    launch().  With analog audio decoding on, the two channels are downloaded with the fields:
    res.audio_host = (left, right) float64 views, valid as long as the fields.  pcm=True (chunks that are consecutive
    ranges of ONE capture, as FileStreamDecoder feeds them): the 48 kHz PCM of every chunk's fields comes down as well,
    chained from chunk to chunk as Framer.readframe chains it: res.pcm_host = list of int16 views (None: dropped field)."""

    def __init__(self, cd, fmt, ncap_max, max_fields=96, np_dtype=np.uint8, nbytes_max=None, pcm=False):
        self.cd, self.fmt, self.max_fields = cd, fmt, max_fields
        self.want_pcm = bool(pcm) and bool(cd.rf.decode_analog_audio)
        self.pcm_state = (0.0, 2)           # (audio_offset, frame_state) carried from chunk to chunk
        self.h_pcm = [None, None]
        rf = cd.rf
        be = self.be = rf._be
        self.up, self.down = be.new_stream(), be.new_stream()
        self.out_stride = (rf.SysParams['frame_lines'] // 2 + 1) * rf.SysParams['outlinelen']
        nelem = ncap_max if nbytes_max is None else nbytes_max
        self.d_in = [be.empty(nelem, np_dtype) for _ in range(2)]
        self.h_out = [be.pinned(max_fields * self.out_stride, np.uint16) for _ in range(2)]
        self.h_status = [be.pinned(max_fields, np.int32) for _ in range(2)]
        self.h_audio = [None, None]         # pinned (left, right) float64 buffers, allocated on first use
        self.in_free = [None, None]         # event: the demodulation that read d_in[k] has finished
        self.nup = self.nlaunch = self.nfin = 0

    def upload(self, host_buf, n, nelem=None, window=None):
        """n: samples of the chunk; nelem: elements of host_buf to copy (defaults to n; packed formats differ).
        window = (cap_base, ncap_total, r0, r1): the chunk holds samples [cap_base, cap_base + n) of a longer capture and
        is decoded as the range [r0, r1) of it (decode_range's arguments); default: the chunk is a capture by itself."""
        be, k = self.be, self.nup % 2
        self.nup += 1
        ne = n if nelem is None else nelem
        if self.in_free[k] is not None:
            be.stream_wait_event(self.up, self.in_free[k])
        with be.stream_ctx(self.up):
            be.copy_async(self.d_in[k][:ne], host_buf[:ne])
            ev = be.record_event()
        return (k, int(n), ev, window)

    def launch(self, ticket):
        """Enqueue the demodulation and the sync-peak chase of an uploaded chunk."""
        be, cd = self.be, self.cd
        k, n, ev, window = ticket
        be.stream_wait_event(be.current_stream_obj(), ev)
        cap_base, ncap_total, r0, r1 = window if window is not None else (0, n, 0, n + 1)
        pend = cd._launch(cd.rf, cd._slot(self.nlaunch % 2, n), self.d_in[k], self.fmt, cap_base, n, ncap_total, r0, r1)
        self.nlaunch += 1
        self.in_free[k] = be.record_event()
        return pend

    def finish(self, pend):
        """Host walk, refinement and TBC of a launched chunk; starts the download of its fields."""
        be = self.be
        res = self.cd._finish(pend, side=self.cd._side_stream() if self.cd.overlap_refine else None)
        j = self.nfin % 2
        self.nfin += 1
        nloc = len(res.located)
        if nloc > self.max_fields:
            raise ValueError("max_fields too small")
        dev = None
        res.audio_host = None
        pcm_job = None
        if self.want_pcm and nloc and res.audio is not None and res.pr.audio2_len:
            d_pcm, off, d_pst, ao, fst = self.cd.pcm_enqueue(res, *self.pcm_state)
            self.pcm_state = (ao, fst)
            pcm_job = (d_pcm, off, d_pst)
        res._pcm_off = None
        if nloc or res.audio is not None:
            done = be.record_event()
            be.stream_wait_event(self.down, done)
            with be.stream_ctx(self.down):
                if nloc:
                    be.copy_async(self.h_out[j][:nloc * self.out_stride], res.d_pic[:nloc * self.out_stride])
                    be.copy_async(self.h_status[j][:nloc], res.d_status[:nloc])
                if res.audio is not None:
                    # the two analog audio channels ride along (pinned, double-buffered like the fields)
                    al, ar = res.audio['audio_left'], res.audio['audio_right']
                    na = len(al)
                    if self.h_audio[j] is None or len(self.h_audio[j][0]) < na:
                        self.h_audio[j] = (be.pinned(na, np.float64), be.pinned(na, np.float64))
                    be.copy_async(self.h_audio[j][0][:na], al)
                    be.copy_async(self.h_audio[j][1][:na], ar)
                    res.audio_host = (be.host_view(self.h_audio[j][0])[:na], be.host_view(self.h_audio[j][1])[:na])
                if pcm_job is not None:
                    d_pcm, off, d_pst = pcm_job
                    if self.h_pcm[j] is None:
                        self.h_pcm[j] = (be.pinned(len(d_pcm), np.int16), be.pinned(self.max_fields, np.int32))
                    npcm = max(off[-1], 1)
                    be.copy_async(self.h_pcm[j][0][:npcm], d_pcm[:npcm])
                    be.copy_async(self.h_pcm[j][1][:nloc], d_pst)
                    res._pcm_off = off
                dev = be.record_event()
            # the slot's picture buffer is rewritten by its next finish(): order that behind this download
            res.slot.pic_free = dev
        return (res, j, nloc, dev)

    def decode(self, ticket):
        return self.finish(self.launch(ticket))

    def fetch(self, job):
        res, j, nloc, dev = job
        if dev is not None:
            self.be.wait_event(dev)
        res.status_host = self.be.host_view(self.h_status[j])[:nloc]
        if nloc and np.any(res.status_host & _lib.ST_LINE_LONG):
            # rare: lines longer than the TBC's staging window; second pass in place, then those fields come down again
            st = self.cd.repair_long_lines(res, res.status_host)
            fixed = np.nonzero(st != res.status_host)[0]
            if len(fixed):
                pic = self.be.to_host(res.d_pic[:nloc * self.out_stride]).reshape(nloc, self.out_stride)
                hv = self.be.host_view(self.h_out[j])[:nloc * self.out_stride].reshape(nloc, self.out_stride)
                for k in fixed:
                    hv[k] = pic[k]
                res.status_host[:] = st
        res.pcm_host = [] if self.want_pcm else None
        if res._pcm_off is not None:
            res.pcm_host = split_pcm(self.be.host_view(self.h_pcm[j][0]), res._pcm_off, self.be.host_view(self.h_pcm[j][1]))
        return res, self.be.host_view(self.h_out[j])[:nloc * self.out_stride].reshape(nloc, self.out_stride)

    def run(self, chunks):
        """chunks: iterable of (pinned host buffer, samples[, elements]).  Yields (RangeResult, host pictures) per chunk."""
        it = iter(chunks)
        up = lambda: (lambda c: self.upload(*c) if c is not None else None)(next(it, None))
        t = up()
        if t is None:
            return
        pend, t = self.launch(t), up()
        prev = None
        while pend is not None:
            nxt = self.launch(t) if t is not None else None     # GPU demodulates chunk k+1 ...
            if t is not None:
                t = up()
            job = self.finish(pend)                              # ... while the host walks chunk k
            if prev is not None:
                yield self.fetch(prev)
            prev, pend = job, nxt
        yield self.fetch(prev)


# bytes per `group` samples of the capture formats (lddutils.py:131-229)
FILE_FORMATS = {_lib.FMT_U8: (1, 1, np.uint8), _lib.FMT_S16: (2, 1, np.int16), _lib.FMT_U16: (2, 1, np.uint16),
                _lib.FMT_R30: (4, 3, np.uint8), _lib.FMT_LDS40: (5, 4, np.uint8)}
FORMAT_OF_SUFFIX = {".lds": _lib.FMT_LDS40, ".r30": _lib.FMT_R30, ".r16": _lib.FMT_S16, ".raw": _lib.FMT_U8, ".u8": _lib.FMT_U8}


class FileStreamDecoder:
    """Decodes a capture FILE from start to end (the reference's loaders, lddutils.py:131-229, and
    the frame loop of lddecode.py:88-98): a reader thread fills page-locked buffers with the file's own bytes (packed
    10-bit formats stay packed: the unpack is fused into the demodulation's block load), HostStreamDecoder uploads,
    decodes and downloads them as consecutive read-position ranges of ONE capture, so the fields that come out are
    exactly those of decoding the whole file at once, in order, whatever the chunk size.

        for res, pics in FileStreamDecoder(cd, "side1.lds"):          # pics: uint16 [nfields, out_stride] host view
            for k, j in enumerate(res.located): write(pics[k, :res.infos[j].linecount * outlinelen])

    pcm=True (analog audio on): res.pcm_host holds the int16 L/R PCM of every field (None: a field ahead of the first
    frame), chained over the whole file as Framer.readframe chains it -- what lddecode.py:98 writes to the .pcm file."""

    def __init__(self, cd, path, fmt=None, chunk_samples=None, first_sample=0, nsamples=None, max_fields=None, pcm=False):
        import os as _os
        self.cd, self.path = cd, path
        self.fmt = FORMAT_OF_SUFFIX[_os.path.splitext(path)[1].lower()] if fmt is None else fmt
        self.bpg, self.group, self.np_dtype = FILE_FORMATS[self.fmt]
        rf = cd.rf
        fsize = _os.path.getsize(path)
        total = fsize // self.bpg * self.group                         # samples in the file
        self.ncap = total if nsamples is None else min(total, first_sample + nsamples)
        self.first = first_sample
        # chunk = the read positions one range owns; ~1 s by default, a multiple of 12 samples (whole .r30 / .lds groups)
        C_ = int(chunk_samples or rf.freq_hz)
        self.chunk = max(C_ // 12 * 12, 12 * rf.blocklen)
        r0s = list(range(first_sample, max(self.ncap - 1, first_sample + 1), self.chunk))
        self.ranges = [(r0, min(r0 + self.chunk, self.ncap + 1) if i + 1 < len(r0s) else self.ncap + 1) for i, r0 in enumerate(r0s)]
        from . import parallel
        self.windows = []
        span = 0
        for r0, r1 in self.ranges:
            lo, hi = parallel.needed_window(cd, self.ncap, r0, r1)
            lo = lo // 12 * 12
            hi = min(self.ncap, -(-hi // 12) * 12)
            self.windows.append((lo, hi))
            span = max(span, hi - lo)
        self.span = span
        fields = max_fields or int(self.chunk / cd.field_samples) + 8
        elem = np.dtype(self.np_dtype).itemsize
        self.nelem_max = -(-span // self.group) * self.bpg // elem + 16
        self.sd = HostStreamDecoder(cd, self.fmt, span, fields, np_dtype=self.np_dtype, nbytes_max=self.nelem_max, pcm=pcm)
        self.be = rf._be
        self.nbuf = 4
        self.bufs = [self.be.pinned(self.nelem_max, self.np_dtype) for _ in range(self.nbuf)]

    def _reader(self, free_q, full_q):
        elem = np.dtype(self.np_dtype).itemsize
        try:
            with open(self.path, "rb", buffering=0) as f:
                for i, (lo, hi) in enumerate(self.windows):
                    b = free_q.get()
                    if b is None:
                        return
                    nbytes = -(-(hi - lo) // self.group) * self.bpg
                    f.seek(lo // self.group * self.bpg)
                    view = memoryview(self.be.host_view(self.bufs[b]).view(np.uint8))[:nbytes]
                    got = 0
                    while got < nbytes:
                        m = f.readinto(view[got:])
                        if not m:
                            break
                        got += m
                    full_q.put((i, b, got // elem))
        finally:
            full_q.put(None)

    def __iter__(self):
        import queue
        import threading
        free_q, full_q = queue.Queue(), queue.Queue()
        for b in range(self.nbuf):
            free_q.put(b)
        th = threading.Thread(target=self._reader, args=(free_q, full_q), daemon=True)
        th.start()
        sd = self.sd

        def next_launch():
            """Next chunk the reader has ready -> (launched range, its host buffer), or None at the end of the file."""
            item = full_q.get()
            if item is None:
                return None
            i, b, nelem = item
            (lo, hi), (r0, r1) = self.windows[i], self.ranges[i]
            return sd.launch(sd.upload(self.bufs[b], hi - lo, nelem, window=(lo, self.ncap, r0, r1))), b

        try:
            pend, prev = next_launch(), None
            while pend is not None:
                nxt = next_launch()                   # the GPU demodulates chunk k+1 ...
                job = sd.finish(pend[0])              # ... while the host walks chunk k (waits for its peak list, so its
                free_q.put(pend[1])                   # upload is over: the reader may refill the buffer)
                if prev is not None:
                    yield sd.fetch(prev)
                prev, pend = job, nxt
            if prev is not None:
                yield sd.fetch(prev)
        finally:
            free_q.put(None)
            # the reader may sit in full_q.put / free_q.get: both queues are unbounded and the sentinel ends it
            th.join(timeout=5)


class CaptureDecoder:
    def __init__(self, rf, readlen=READLEN, mtf_level=1, colorlevel=1.45, colorphase=91.5, max_fields=8192):
        self.rf = rf
        self.readlen = readlen
        self.mtf_level = mtf_level          # Framer starts at 1 (lddecode_core.py:1334)
        self.colorlevel, self.colorphase = colorlevel, colorphase
        self.max_fields = max_fields
        self.field_samples = int(rf.freq_hz / rf.SysParams['FPS'] / 2)
        # Plane workspaces (a 1-s PAL range needs 0.9 GB of planes) are kept and reused across calls; results of
        # decode_range therefore stay valid until the next call that uses the same slot (decode / decode_range:
        # slot 0; the streaming decoders alternate between slots 0 and 1).
        self._slots = {}
        self._lanes = None          # extra (RFDecode, stream, slot key) sets of decode_pipelined
        self._side = None
        self.overlap_refine = os.environ.get("LDD_NO_REFINE_OVERLAP") is None
        # CAV discs: follow the reference's per-frame MTF adaptation (Framer.readframe, lddecode_core.py:1300-1306)
        self.cav_follow = False
        self.cav_state = None       # (capture sample where a frame starts, mtf level the reference uses for that frame)
        self.cav_hold = None        # (capture sample where the run's first frame ends, the start-up level it is decoded with)
        self._cav_pending = []      # VBI codes of finished ranges on their way to the host
        self._cav_first = None      # where the first complete frame of the run ends

    # -- workspaces
    def _slot(self, key, ncap_window, rf=None):
        """Plane workspace `key`, big enough for a range over a capture window of ncap_window samples."""
        rf = rf or self.rf
        need = int(ncap_window) + 2 * rf.blocklen
        s = self._slots.get(key)
        if s is None or s.plane_cap < need or s.rf is not rf:
            if s is not None:
                rf._be.synchronize()
                s.close()
            cap = int(need * 1.02) + 16
            nf = min(self.max_fields, cap // (100 * rf.linelen) + 16)
            s = PipeSlot(self, rf, cap, nf)
            s.pic_free = None
            self._slots[key] = s
        return s

    def decode_stream(self, captures, sink=None):
        """Software-pipelined decode() of a sequence of device-resident captures [(cap_dev, fmt, ncap), ...] or of ranges
        of one resident capture window [(cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1), ...] (decode_range's
        arguments: how a shard that is longer than the plane workspace is worked through):
        capture k+1 is demodulated while the host walks capture k (one stream, two plane workspaces).
        Yields one RangeResult per capture; its device buffers stay valid until the second-next launch.
        sink (parallel.FieldGatherer): every capture's pictures are written straight into the sink's send buffer
        and gathered."""
        it = iter(captures)
        k = 0

        def launch(c):
            nonlocal k
            if len(c) == 3:                       # a whole capture as one range
                c = (c[0], c[1], 0, c[2], c[2], 0, c[2] + 1)
            cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1 = c
            # plane workspace for the range's own block grid (not for the whole resident window)
            span = min(cap_len, (r1 - r0) + self.readlen + int(1.6 * self.field_samples) + 4 * self.rf.blocklen)
            slot = self._slot(k % 2, span)
            k += 1
            return self._launch(self.rf, slot, cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1)

        c = next(it, None)
        pend = launch(c) if c is not None else None
        while pend is not None:
            c = next(it, None)
            nxt = launch(c) if c is not None else None
            side = self._side_stream() if self.overlap_refine else None
            if sink is not None:
                pic, status = sink.buffers(side)
                res = self._finish(pend, side=side, pic_out=pic, status_out=status)
                sink.gather(res)
            else:
                res = self._finish(pend, side=side)
            yield res
            pend = nxt

    @property
    def stride(self):
        rf = self.rf
        return rf.blocklen - rf.blockcut - rf.blockcut_end

    def plan_range(self, ncap_total, r0, r1):
        """Block grid for the range owning read positions [r0, r1): (first_block, nblocks, walk_start).
        (The library plans the same way inside ldd_pipe_launch; this copy serves callers that have to cut the
        capture window a range needs, parallel.needed_window.)"""
        rf = self.rf
        S, N, bc = self.stride, rf.blocklen, rf.blockcut
        walk_start = 0 if r0 <= 0 else max(0, r0 - int(1.6 * self.field_samples))
        first_block = (max(walk_start - bc, 0) // S) * S
        # last capture sample any owned window can need: the reads of rf.demod(r1 - 1, readlen)
        need_end = min(ncap_total, r1 + self.readlen + 2 * N + bc)
        nblocks = max(0, (need_end - first_block - N) // S + 1)
        return first_block, nblocks, walk_start

    def decode_range(self, cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1, want_tables=False, audio_phase2=True):
        """Decode the fields whose read position lies in [r0, r1).

        cap_dev holds capture samples [cap_base, cap_base + cap_len) in format fmt; ncap_total is the
        length of the whole capture (the reference stops when a read would pass its end)."""
        pend = self._launch(self.rf, self._slot(0, cap_len), cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1, audio_phase2)
        return self._finish(pend)

    def _launch(self, rf, slot, cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1, audio_phase2=True):
        """Stage 1 (asynchronous): demodulate the range's blocks, chase the sync peaks, start their copy
        to pinned host memory, run the second audio stage."""
        be = rf._be
        rf._set_mtf(self.mtf_level)           # uploads the tables on first use; the level itself is set by the library
        level = self.mtf_level
        if self.cav_follow:
            level = self._cav_apply(rf, max(r0, cap_base))
        rf._check(be.lib.ldd_pipe_launch(slot.h, be.ptr(cap_dev), int(fmt), int(cap_base), int(cap_len), int(ncap_total),
                                         int(r0), int(r1), int(self.readlen), float(level), int(bool(audio_phase2)),
                                         be.stream()))
        return Pending(slot, rf, r0, r1, ncap_total)

    def _side_stream(self):
        """High-priority stream for the refine + TBC kernels of a streaming decode: the next capture's demodulation is
        already enqueued on the main stream when they are launched; on their own stream they fill the SMs that the tail
        of its float32 pass and its float64 re-run (one CTA on 4 % of the blocks) leave idle."""
        if self._side is None:
            self._side = self.rf._be.new_stream(high_priority=True)
        return self._side

    def _finish(self, pend, side=None, pic_out=None, status_out=None, frame_mode=False, pic_stride=None, pic_cap=None):
        """Stage 2: ldd_pipe_finish (host walk over the peak list, then the batched refine + VBI + TBC launches, on
        `side` when given; the current stream is ordered behind them).  pic_out / status_out: device buffers that
        receive the pictures instead of the slot's own (e.g. a collective's send buffer)."""
        slot, rf = pend.slot, pend.rf
        be = rf._be
        main = be.stream()
        ref = C.c_void_p(side.cuda_stream) if side is not None else main
        if slot.pic_free is not None and pic_out is None:
            # a streaming download of this slot's previous pictures may still be running: the stream that rewrites
            # them (not the demodulation of the next capture) waits for it
            be.stream_wait_event(side if side is not None else be.current_stream_obj(), slot.pic_free)
            slot.pic_free = None
        d_pic = slot.d_pic if pic_out is None else pic_out
        d_status = slot.d_status if status_out is None else status_out
        stride = slot.out_stride if pic_stride is None else int(pic_stride)
        cap = (len(d_pic) // stride) if pic_cap is None else int(pic_cap)
        pr = slot.result
        rf._check(be.lib.ldd_pipe_finish(slot.h, float(self.colorlevel), float(self.colorphase), int(bool(frame_mode)), be.ptr(d_pic),
                                         stride, cap, be.ptr(d_status), ref, main, C.byref(pr)))
        res = RangeResult()
        res.slot, res.pr = slot, pr
        res.r0, res.r1, res.ncap_total = pend.r0, pend.r1, pend.ncap_total
        res.plane_origin, res.plane_len, res.walk_start = pr.plane_origin, pr.plane_len, pr.walk_start
        total = pr.plane_len
        res.planes = {n: t[:total] for n, t in slot.planes.items()}
        res.audio = None
        if slot.a1 is not None:
            if pr.audio2_len:
                res.audio = {'audio_left': slot.a2[0][:pr.audio2_len], 'audio_right': slot.a2[1][:pr.audio2_len]}
            else:
                res.audio = {'audio_left': slot.a1[0][:pr.audio1_len], 'audio_right': slot.a1[1][:pr.audio1_len]}
        nw = pr.nwindows
        res._nwin_all = nw
        fields = _host_array(pr.fields, nw, _lib.FieldInfo, np.uint8).view(FIELD_DTYPE) if nw else np.zeros(0, dtype=FIELD_DTYPE)
        owned = _host_array(pr.owned, pr.nowned, C.c_int, np.int32)
        res._owned = owned
        res._infos = fields[owned].view(np.recarray)
        res.readsamples = _host_array(pr.readsample, nw, C.c_longlong, np.int64)[owned]
        res.base = _host_array(pr.base, nw, C.c_longlong, np.int64)[owned]
        res.nwindows = int(pr.nowned)
        res.located = [int(j) for j in _host_array(pr.located, pr.nlocated, C.c_int, np.int32)]
        res.frame_of = _host_array(pr.frame_of, pr.nlocated, C.c_int, np.int32) if frame_mode else None
        res.nframes = int(pr.nframes)
        res.prefix_windows = int(pr.prefix_windows)
        res._gpeaks = res._l1 = None
        res._npeaks = int(pr.npeaks)
        slot.generation += 1
        res._generation = slot.generation
        res.lineloc_add = pr.lineloc_add
        res.out_stride = stride
        nloc = len(res.located)
        res.d_pic = d_pic[:(cap if frame_mode else nloc) * stride] if nloc else None
        res.d_status = d_status[:nloc] if nloc else None
        res.refined = Refined(res) if nloc else None
        if self.cav_follow and nloc:
            self._cav_collect(res, side)
        return res

    # -- CAV: per-frame MTF level (lddecode_core.py:1300-1306) predicted from the frame numbers already decoded
    def _cav_apply(self, rf, start_sample):
        """Sets the library's per-block level ramp for a range that starts at capture sample start_sample; returns the
        level to bake in (the level of the frame that contains start_sample).  Without a decoded frame number yet the
        ramp is off and the level is self.mtf_level (the reference's start value)."""
        self._cav_drain(block=len(self._cav_pending) >= 2)
        if self.cav_state is None:
            rf.set_mtf_ramp(0.0, 0.0, 0.0)
            return self.mtf_level
        pos, level = self.cav_state
        period = rf.freq_hz / rf.SysParams['FPS']
        n0 = np.floor((start_sample - pos) / period)
        hold = self.cav_hold if self.cav_hold is not None else (-1e300, 1.0)
        rf.set_mtf_ramp(pos + n0 * period, period, -1e-4, hold[0], hold[1])
        return max(level - 1e-4 * n0, 0.0)

    def _cav_collect(self, res, side):
        """Starts the copy of a finished range's VBI codes to page-locked memory (behind its refinement kernels)."""
        slot, be = res.slot, res.slot.be
        nloc = len(res.located)
        if getattr(slot, 'h_vbi', None) is None:
            slot.h_vbi = be.pinned(16 * slot.max_fields, np.uint8)
        off = res.pr.d_vbi - slot.tables_addr
        with be.stream_ctx(side):
            res.slot.rf._check(be.lib.ldd_copy_small(be.ptr(slot.h_vbi), be.ptr(slot.d_tables[off:off + 16 * nloc]), 16 * nloc, be.stream()))
            ev = be.record_event()
        loc = np.asarray(res.located, dtype=np.intp)
        self._cav_pending.append((ev, slot, nloc, res.readsamples[loc].copy(), res.infos[loc].copy()))

    def _cav_drain(self, block=False):
        be = self.rf._be
        while self._cav_pending:
            ev, slot, nloc, rs, infos = self._cav_pending[0]
            if ev is not None and not block and not ev.query():
                return
            if ev is not None:
                be.wait_event(ev)
            self._cav_pending.pop(0)
            block = False
            codes = be.host_view(slot.h_vbi)[:16 * nloc].view(np.int32).reshape(nloc, 4).copy()
            self._cav_update(rs, infos, codes)

    def _cav_update(self, readsamples, infos, codes):
        """Framer.readframe's rule on a range's fields: after a complete frame (first field istop == topfirst, then its
        partner) whose LAST field carries a CAV frame number, the next frame is decoded with max(1 - framenr/10000, 0)."""
        rf = self.rf
        top = int(bool(rf.SysParams['topfirst']))
        lines = rf.SysParams['philips_codelines']
        k = 0
        while k + 1 < len(infos):
            if int(infos[k].istop) == top and int(infos[k + 1].istop) != top:
                vbi = F.process_philips(rf, {l: F.code_nibbles(codes[k + 1][i]) for i, l in enumerate(lines)})
                if not vbi['isclv'] and vbi['framenr'] is not None:
                    # the next frame's first read starts 10 lines before its vertical interval (Field.nextfieldoffset,
                    # lddecode_core.py:926); this frame's last picture lines end there: switch the level at the interval
                    nxt = int(readsamples[k + 1]) + int(infos[k + 1].nextfieldoffset) + 10 * rf.linelen
                    self.cav_state = (nxt, max(1 - vbi['framenr'] / 10000, 0))
                    if self._cav_first is None:
                        self._cav_first = nxt
                k += 2
            else:
                k += 1

    def prime_cav(self, cap_dev, fmt, cap_base, cap_len, ncap_total, r0=0):
        """Decodes the first frame at r0 (synchronously) to learn the disc position, so that the ranges that follow get
        the reference's level for every frame.  Returns the state (frame start sample, level) or None (CLV / no code)."""
        self.cav_follow = True
        self._cav_first = None
        nfr = int(2.6 * 2 * self.field_samples)
        res = self.decode_range(cap_dev, fmt, cap_base, cap_len, ncap_total, r0, min(r0 + nfr, ncap_total + 1))
        self._cav_drain(block=True)
        if self.cav_state is not None:
            # the run's first frame keeps the start-up level (Framer.mtf_level = 1 until the first frame number is read)
            self.cav_hold = (self._cav_first, self.mtf_level)
        return self.cav_state

    def decode_pipelined(self, cap_dev, fmt, ncap, nranges=2):
        """The whole capture as `nranges` read-position ranges on their own streams and handles, so that
        the host walk of one range overlaps the demodulation / resampling of the others.  The union of
        the ranges' fields is bit-identical to decode() (same global block grid).  Returns the list of
        RangeResults in capture order."""
        from . import parallel, rfdecode
        rf0, be = self.rf, self.rf._be
        if self._lanes is None or len(self._lanes) < nranges:
            lanes = list(self._lanes or [])
            while len(lanes) < nranges:
                rfk = rf0 if not lanes else rfdecode.RFDecode(rf0.freq, rf0.system, rf0.blocklen, rf0.decode_analog_audio,
                                                              precision=rf0.precision, _backend=be)
                lanes.append((rfk, be.new_stream(), ('lane', len(lanes))))
            self._lanes = lanes
        main = be.current_stream_obj()
        bounds = parallel.shard_bounds(ncap, nranges)
        pend = []
        for (rfk, sk, key), (r0, r1) in zip(self._lanes, bounds):
            if sk is not None:
                be.wait_stream(sk, main)
            with be.stream_ctx(sk):
                pend.append(self._launch(rfk, self._slot(key, ncap, rfk), cap_dev, fmt, 0, ncap, ncap, r0, r1))
        out = []
        for (rfk, sk, key), p in zip(self._lanes, pend):
            with be.stream_ctx(sk):
                out.append(self._finish(p))
            if sk is not None:
                be.wait_stream(main, sk)
        return out

    def decode(self, cap_dev, fmt, ncap, want_tables=False, audio_phase2=True):
        """The whole capture as one range."""
        return self.decode_range(cap_dev, fmt, 0, ncap, ncap, 0, ncap + 1, want_tables, audio_phase2)

    def decode_frames(self, cap_dev, fmt, ncap):
        """The whole capture as interleaved uint16 frames written by the TBC kernel itself (Framer.formatoutput,
        lddecode_core.py:1238-1252; fields paired by parity as Framer.readframe does for CLV discs).  Returns
        (RangeResult, frames uint16 device view [nframes][frame_lines * outlinelen])."""
        rf = self.rf
        slot = self._slot(0, ncap)
        fstride = rf.SysParams['frame_lines'] * rf.SysParams['outlinelen']
        pend = self._launch(rf, slot, cap_dev, fmt, 0, ncap, ncap, 0, ncap + 1)
        cap = len(slot.d_pic) // fstride
        res = self._finish(pend, frame_mode=True, pic_stride=fstride, pic_cap=cap)
        return res, slot.d_pic[:res.nframes * fstride].reshape(-1, fstride)

    # -- 48 kHz PCM of a range (SURVEY section 8f-1 in pipeline mode)
    def pcm_enqueue(self, res, audio_offset=0.0, frame_state=2, chain='framer', freq=48000.0, scale=64):
        """Asynchronous part of pcm(): enqueues ldd_pipe_pcm for `res` on the current stream.  Returns
        (device int16 buffer, offsets [nlocated + 1] into it, device status [nlocated], audio_offset, frame_state)."""
        res._live()
        slot, rf = res.slot, res.slot.rf
        be = rf._be
        n = len(res.located)
        if n == 0:
            return None, [0], None, audio_offset, frame_state
        if res.audio is None or not res.pr.audio2_len:
            raise ValueError("pcm() needs a range decoded with the analog audio and its second stage on")
        SP = rf.SysParams
        if scale is None:                   # the decoder's true decimation instead of the reference's fixed 64
            scale = rf.audio_decimation
        per_field = int((SP['line_period'] * (SP['frame_lines'] // 2 + 1) / 1e6) * freq) + 4
        cap = 2 * per_field * slot.max_fields
        if getattr(slot, 'd_pcm', None) is None or len(slot.d_pcm) < cap:
            slot.d_pcm = be.empty(cap, np.int16)
            slot.d_pcm_status = be.zeros(slot.max_fields, np.int32)
        off = (C.c_longlong * (n + 1))()
        ao, fs = C.c_double(float(audio_offset)), C.c_int(int(frame_state))
        be.fill_zero(slot.d_pcm_status[:n])
        rf._check(be.lib.ldd_pipe_pcm(slot.h, float(freq), float(scale), float(SP['line_period']), float(SP['audio_lfreq']),
                                      float(SP['audio_rfreq']), _lib.PCM_CHAIN_FRAMER if chain == 'framer' else _lib.PCM_CHAIN_FIELDS,
                                      C.byref(ao), C.byref(fs), be.ptr(slot.d_pcm), cap, off, be.ptr(slot.d_pcm_status),
                                      be.stream()))
        return slot.d_pcm, [int(x) for x in off], slot.d_pcm_status[:n], ao.value, fs.value

    def pcm(self, res, audio_offset=0.0, frame_state=2, chain='framer', freq=48000.0, scale=64):
        """int16 L/R PCM of the located fields of `res` (the latest result of its workspace, decoded with the analog
        audio on): downscale_audio (lddecode_core.py:431-484) per field on the final line positions, on the device, from
        the range's phase-2 audio (ldd_pipe_pcm).  chain='framer': time offsets as Framer.readframe carries them for CLV
        discs (every field of a frame starts at the offset the frame started with, fields ahead of the first frame are
        dropped while bit 1 of frame_state is set); chain='fields': every field continues the previous one.
        scale: the reference's `scale` argument (64, whatever the sample rate); None = the decoder's true decimation.
        Returns (list of int16 arrays | None for a dropped or flagged field, audio_offset, frame_state) -- pass the two
        state values to the next range of the same capture."""
        be = res.slot.rf._be
        d_pcm, off, d_st, audio_offset, frame_state = self.pcm_enqueue(res, audio_offset, frame_state, chain, freq, scale)
        if d_pcm is None:
            return [], audio_offset, frame_state
        be.synchronize()
        return split_pcm(be.to_host(d_pcm[:max(off[-1], 1)]), off, be.to_host(d_st)), audio_offset, frame_state

    # -- lines too long for the TBC's staging window (fields whose line location partly failed)
    def repair_long_lines(self, res, st=None):
        """Fields of `res` whose status says 1 | ST_LINE_LONG without ST_LINE_BAD had lines longer than 1.25 x nominal,
        which the TBC pass leaves out; the reference resamples any span (lddutils.py:83-97).  Runs the second pass
        (ldd_pipe_long_lines: exact kernel, spans up to 4032 samples) over the range's pictures in place and clears the
        bits of the fields it completes, on the device and in the returned host copy.  Synchronous; a no-op (no launch)
        when no field is flagged, which is the normal case."""
        be = res.slot.rf._be
        if res.d_status is None:
            return np.zeros(0, dtype=np.int32)
        if st is None:
            be.synchronize()
            st = be.to_host(res.d_status)
        st = np.array(st[:len(res.located)], dtype=np.int32)
        need = ((st & _lib.ST_LINE_LONG) != 0) & ((st & _lib.ST_LINE_BAD) == 0)
        if not need.any() or res.slot.generation != res._generation:
            return st
        d2 = be.zeros(len(st), np.int32)
        res.slot.rf._check(be.lib.ldd_pipe_long_lines(res.slot.h, be.ptr(d2), be.stream()))
        be.synchronize()
        done = need & (be.to_host(d2) == 0)
        st[done] &= ~(1 | _lib.ST_LINE_LONG)
        be.copy_async(res.d_status, be.to_device(st))
        be.synchronize()
        return st

    # -- host copies
    def pictures(self, res):
        """uint16 TBC fields of the located windows -> list of (readsample, istop, array | None)."""
        rf, be = self.rf, self.rf._be
        if res.d_pic is None:
            return []
        be.synchronize()
        st = self.repair_long_lines(res)
        pic = be.to_host(res.d_pic).reshape(len(res.located), res.out_stride)
        W = rf.SysParams['outlinelen']
        out = []
        for k, j in enumerate(res.located):
            ok = (st[k] & (1 | 2 | 4 | 8)) == 0
            info = res.infos[j]
            out.append((int(res.readsamples[j]), int(info.istop), pic[k, :info.linecount * W].copy() if ok else None))
        return out

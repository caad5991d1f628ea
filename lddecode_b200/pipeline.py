"""Whole-capture decode with everything resident in HBM: the throughput path behind bench.py.

The reference walks a capture field by field (Framer.readfield, lddecode_core.py:1194-1223): every
field re-reads and re-demodulates 1e6 samples (2.1x redundant) and the next read position comes out
of the previous field.  Here a range of the capture is demodulated ONCE on the fixed global block
grid (blocks start at multiples of blocklen-1056; they are independent given the 1024/32-sample
halos), one sync-peak chase covers the whole range, the field-to-field walk
runs on the host over the ~16 K peaks per second of video (ldd_field_chain), and all located fields
are refined and resampled in a handful of batched launches.  Only the peak list crosses to the host
in the middle.

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


class RangeResult:
    """Fields owned by one range.  Device buffers: d_pic uint16 [nfields][out_stride], d_status."""
    pass


class HostStreamDecoder:
    """A sequence of host-resident captures (the chunks a file reader produces) through one
    CaptureDecoder, software-pipelined: while the host walks the fields of chunk k the GPU already
    demodulates chunk k+1 (same stream, second plane workspace), and the upload of chunk k+2 and the
    download of chunk k-1's fields run on their own streams.  Device input, planes and pinned output
    are double-buffered.  Results are those of CaptureDecoder.decode() of every chunk.

        for res, pics in sd.run(chunks): ...                    # or, step by step:
        t = sd.upload(buf, n); p = sd.launch(t); job = sd.finish(p); res, pics = sd.fetch(job)

    upload() takes a pinned host buffer (be.pinned).  fetch() returns the RangeResult and a host view
    [nfields, out_stride] of its uint16 fields that stays valid until the second-next finish(); the
    RangeResult's device planes stay valid until the second-next launch().  With analog audio decoding on, the two
    channels are downloaded with the fields: res.audio_host = (left, right) float64 views, valid as long as the fields."""

    def __init__(self, cd, fmt, ncap_max, max_fields=96, np_dtype=np.uint8):
        self.cd, self.fmt, self.max_fields = cd, fmt, max_fields
        rf = cd.rf
        be = self.be = rf._be
        self.up, self.down = be.new_stream(), be.new_stream()
        self.out_stride = (rf.SysParams['frame_lines'] // 2 + 1) * rf.SysParams['outlinelen']
        self.d_in = [be.empty(ncap_max, np_dtype) for _ in range(2)]
        self.h_out = [be.pinned(max_fields * self.out_stride, np.uint16) for _ in range(2)]
        self.h_status = [be.pinned(max_fields, np.int32) for _ in range(2)]
        self.h_audio = [None, None]         # pinned (left, right) float64 buffers, allocated on first use
        self.in_free = [None, None]         # event: the demodulation that read d_in[k] has finished
        self.nup = self.nlaunch = self.nfin = 0

    def upload(self, host_buf, n):
        be, k = self.be, self.nup % 2
        self.nup += 1
        if self.in_free[k] is not None:
            be.stream_wait_event(self.up, self.in_free[k])
        with be.stream_ctx(self.up):
            be.copy_async(self.d_in[k][:n], host_buf[:n])
            ev = be.record_event()
        return (k, int(n), ev)

    def launch(self, ticket):
        """Enqueue the demodulation and the sync-peak chase of an uploaded chunk."""
        be, cd = self.be, self.cd
        k, n, ev = ticket
        be.stream_wait_event(be.current_stream_obj(), ev)
        ws, staging = cd._lane(self.nlaunch % 2)
        self.nlaunch += 1
        pend = cd._launch_demod(cd.rf, ws, staging, self.d_in[k], self.fmt, 0, n, n, 0, n + 1)
        self.in_free[k] = be.record_event()
        return pend

    def finish(self, pend):
        """Host walk, refinement and TBC of a launched chunk; starts the download of its fields."""
        be = self.be
        res = self.cd._finish_range(self.cd.rf, pend, side=self.cd._side_stream() if self.cd.overlap_refine else None)
        j = self.nfin % 2
        self.nfin += 1
        nloc = len(res.located)
        if nloc > self.max_fields:
            raise ValueError("max_fields too small")
        dev = None
        res.audio_host = None
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
                dev = be.record_event()
        return (res, j, nloc, dev)

    def decode(self, ticket):
        return self.finish(self.launch(ticket))

    def fetch(self, job):
        res, j, nloc, dev = job
        if dev is not None:
            self.be.wait_event(dev)
        res.status_host = self.be.host_view(self.h_status[j])[:nloc]
        return res, self.be.host_view(self.h_out[j])[:nloc * self.out_stride].reshape(nloc, self.out_stride)

    def run(self, chunks):
        """chunks: iterable of (pinned host buffer, length).  Yields (RangeResult, host pictures) per chunk."""
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


class CaptureDecoder:
    def __init__(self, rf, readlen=READLEN, mtf_level=1, colorlevel=1.45, colorphase=91.5, max_fields=8192):
        self.rf = rf
        self.readlen = readlen
        self.mtf_level = mtf_level          # Framer starts at 1 (lddecode_core.py:1334)
        self.colorlevel, self.colorphase = colorlevel, colorphase
        self.max_fields = max_fields
        self.field_samples = int(rf.freq_hz / rf.SysParams['FPS'] / 2)
        # Plane / audio buffers are kept and reused across calls (a 1-s PAL range needs 0.9 GB of
        # planes; re-allocating that per call stalls on cudaMalloc).  Results of decode_range
        # therefore stay valid until the next call on the same CaptureDecoder.
        self._ws = {}
        self._staging = {}
        self._lanes = None          # extra (RFDecode, stream, workspace, staging) sets of decode_pipelined
        self._lane2 = None          # second workspace of decode_stream / HostStreamDecoder
        self._side = None
        self.overlap_refine = os.environ.get("LDD_NO_REFINE_OVERLAP") is None

    def _lane(self, i):
        """Workspace + staging set i (0 = the default one); a second set lets the demodulation of the
        next capture run while the fields of the current one are still being walked and resampled."""
        if i == 0:
            return self._ws, self._staging
        if self._lane2 is None:
            self._lane2 = ({}, {})
        return self._lane2

    def decode_stream(self, captures):
        """Software-pipelined decode() of a sequence of device-resident captures [(cap_dev, fmt, ncap), ...]:
        capture k+1 is demodulated while the host walks capture k (one stream, two plane workspaces).
        Yields one RangeResult per capture; its device buffers stay valid until the second-next launch."""
        it = iter(captures)
        k = 0

        def launch(c):
            nonlocal k
            ws, staging = self._lane(k % 2)
            k += 1
            return self._launch_demod(self.rf, ws, staging, c[0], c[1], 0, c[2], c[2], 0, c[2] + 1)

        c = next(it, None)
        pend = launch(c) if c is not None else None
        while pend is not None:
            c = next(it, None)
            nxt = launch(c) if c is not None else None
            yield self._finish_range(self.rf, pend, side=self._side_stream() if self.overlap_refine else None)
            pend = nxt

    def _buf(self, tag, n, dtype, ws=None):
        ws = self._ws if ws is None else ws
        key = (tag, np.dtype(dtype).str)
        b = ws.get(key)
        if b is None or len(b) < n:
            b = None
            ws.pop(key, None)
            b = self.rf._be.empty(int(n * 1.02) + 16, dtype)
            ws[key] = b
        return b[:n]

    @property
    def stride(self):
        rf = self.rf
        return rf.blocklen - rf.blockcut - rf.blockcut_end

    def plan_range(self, ncap_total, r0, r1):
        """Block grid for the range owning read positions [r0, r1): (first_block, nblocks, walk_start)."""
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
        pend = self._launch_demod(self.rf, self._ws, self._staging, cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1,
                                  audio_phase2)
        return self._finish_range(self.rf, pend, want_tables)

    def _launch_demod(self, rf, ws, staging, cap_dev, fmt, cap_base, cap_len, ncap_total, r0, r1, audio_phase2=True):
        """Stage 1 (asynchronous): demodulate the range's blocks, chase the sync peaks, start their copy
        to pinned host memory."""
        be = rf._be
        S, N = self.stride, rf.blocklen
        first_block, nblocks, walk_start = self.plan_range(ncap_total, r0, r1)
        avail_end = min(cap_base + cap_len, ncap_total)
        while nblocks > 0 and first_block + (nblocks - 1) * S + N > avail_end:
            nblocks -= 1
        if first_block < cap_base:
            raise ValueError("capture window does not cover the range's halo")
        total = nblocks * S
        res = RangeResult()
        res.r0, res.r1, res.plane_origin, res.plane_len = r0, r1, first_block, total
        res.ncap_total, res.walk_start = ncap_total, walk_start
        res.staging = staging
        rf._set_mtf(self.mtf_level)
        planes, parr = rf._alloc_planes(max(total, 1), alloc=lambda name, n, dt: self._buf("plane_" + name, n, dt, ws))
        a1l = a1r = None
        alen = 0
        if rf.decode_analog_audio:
            ds = N // len(rf.Filters['audio_lfilt'])
            alen = total // ds
            a1l, a1r = self._buf("a1l", max(alen, 1), np.float64, ws), self._buf("a1r", max(alen, 1), np.float64, ws)
        if nblocks:
            rf._check(be.lib.ldd_demod_blocks(rf._h, be.ptr(cap_dev), fmt, int(cap_base), int(cap_len), int(first_block),
                                              int(nblocks), int(total), parr, be.ptr(a1l) if a1l is not None else None,
                                              be.ptr(a1r) if a1r is not None else None, int(alen), be.stream()))
        res.planes = planes
        # sync-peak chase over the whole plane; its result is the only device->host hop of the path
        res.pending_peaks = F.sync_peaks_launch(rf, planes['demod_sync'], total, 0, staging)
        # the second audio stage does not depend on the walk: enqueue it behind the chase so that it
        # runs while the host walks the fields
        res.audio = None
        if rf.decode_analog_audio:
            if audio_phase2 and alen > rf.blocklen:
                res.audio = rf._audio_phase2_device(a1l, a1r, alen)
            else:
                res.audio = {'audio_left': a1l, 'audio_right': a1r}
        return res

    def _side_stream(self):
        """High-priority stream for the refine + TBC kernels of a streaming decode: the next capture's demodulation is
        already enqueued on the main stream when they are launched; on their own stream they fill the SMs that the tail
        of its float32 pass and its float64 re-run (one CTA on 4 % of the blocks) leave idle."""
        if self._side is None:
            self._side = self.rf._be.new_stream(high_priority=True)
        return self._side

    def _finish_range(self, rf, res, want_tables=False, side=None):
        """Stage 2: host walk over the peak list, then the batched refine + TBC launches (on `side` when given; the
        current stream is ordered behind them before this returns, so callers keep using the current stream)."""
        planes, total, r0, r1 = res.planes, res.plane_len, res.r0, res.r1
        ready_ev = res.pending_peaks.ev                      # recorded behind the demodulation and the peak chase
        gpk, gvl = res.pending_peaks.result()
        res.pending_peaks = None
        res.gpeaks = gpk
        batch, infos, readsamples = self._walk(rf, planes, total, res.plane_origin, res.ncap_total, res.walk_start, r1,
                                               r0 > 0, gpk, gvl, res.staging)
        n_all = len(infos)
        stages = np.fromiter((f.stage for f in infos), dtype=np.int32, count=n_all)
        owned = np.nonzero((readsamples >= r0) & (readsamples < r1))[0]
        res.infos = [infos[i] for i in owned]
        res.readsamples = readsamples[owned]
        res.base = batch.base[owned]
        res.linelocs1 = batch.linelocs1[owned]
        res.nwindows = len(owned)
        loc_mask = stages[owned] == _lib.FIELD_LOCATED
        located = np.nonzero(loc_mask)[0]
        res.located = [int(j) for j in located]
        res.refined = None
        res.d_pic = res.d_status = None
        res.out_stride = (rf.SysParams['frame_lines'] // 2 + 1) * rf.SysParams['outlinelen']
        if len(located):
            idx = owned[located]
            sub = F.FieldBatch.view(rf, batch, idx, np.fromiter((infos[i].linecount for i in idx), dtype=np.int32, count=len(idx)))
            be = rf._be
            if side is not None:
                be.stream_wait_event(side, ready_ev)
                with be.stream_ctx(side):
                    ref = F.refine_and_tbc(rf, planes, total, sub, self.colorlevel, self.colorphase, want_intermediates=want_tables,
                                           staging=res.staging)
                    done = be.record_event()
                be.stream_wait_event(be.current_stream_obj(), done)
            else:
                ref = F.refine_and_tbc(rf, planes, total, sub, self.colorlevel, self.colorphase, want_intermediates=want_tables,
                                       staging=res.staging)
            res.refined = ref
            res.d_pic, res.d_status = ref.d_pic, ref.d_status
        return res

    def decode_pipelined(self, cap_dev, fmt, ncap, nranges=2):
        """The whole capture as `nranges` read-position ranges on their own streams and handles, so that
        the host walk of one range overlaps the demodulation / resampling of the others.  The union of
        the ranges' fields is bit-identical to decode() (same global block grid).  Returns the list of
        RangeResults in capture order."""
        from . import parallel, rfdecode
        rf0, be = self.rf, self.rf._be
        if self._lanes is None or len(self._lanes) < nranges:
            lanes = [(rf0, None, self._ws, self._staging)] if self._lanes is None else self._lanes
            while len(lanes) < nranges:
                rfk = rfdecode.RFDecode(rf0.freq, rf0.system, rf0.blocklen, rf0.decode_analog_audio, precision=rf0.precision,
                                        _backend=be)
                lanes.append((rfk, be.new_stream(), {}, {}))
            if lanes[0][1] is None:
                lanes[0] = (rf0, be.new_stream(), self._ws, self._staging)
            self._lanes = lanes
        main = be.current_stream_obj()
        bounds = parallel.shard_bounds(ncap, nranges)
        pend = []
        for (rfk, sk, ws, stg), (r0, r1) in zip(self._lanes, bounds):
            if sk is not None:
                be.wait_stream(sk, main)
            with be.stream_ctx(sk):
                pend.append(self._launch_demod(rfk, ws, stg, cap_dev, fmt, 0, ncap, ncap, r0, r1))
        out = []
        for (rfk, sk, ws, stg), p in zip(self._lanes, pend):
            with be.stream_ctx(sk):
                out.append(self._finish_range(rfk, p))
            if sk is not None:
                be.wait_stream(main, sk)
        return out

    def decode(self, cap_dev, fmt, ncap, want_tables=False, audio_phase2=True):
        """The whole capture as one range."""
        return self.decode_range(cap_dev, fmt, 0, ncap, ncap, 0, ncap + 1, want_tables, audio_phase2)

    # -- host walk
    def _walk(self, rf, planes, total, plane_origin, ncap_total, first_readsample, stop_readsample, tolerant, gpk, gvl,
              staging=None):
        be = rf._be
        cb_staging = staging.setdefault('cb', {}) if staging is not None else {}
        mf = self.max_fields
        fields = (_lib.FieldInfo * mf)()
        batch = F.FieldBatch(rf, mf)
        readsample = np.zeros(mf, dtype=np.int64)
        nf = C.c_int(0)
        keep = {}
        gpk = np.ascontiguousarray(gpk, dtype=np.int64)
        gvl = np.ascontiguousarray(gvl, dtype=np.float64)

        L = rf.linelen
        half, skip = L // 2, int(L * .4)

        def window_peaks(ctx, b, wl, ppk, pvl, pn):
            # This window does not start on a peak of the range's chase (second read of a capture,
            # range starts).  Chase a short prefix by itself; from the first peak it shares with the
            # range's chase on, the two lists are the same, cut at the reference's loop bound.
            b, wl = int(b), int(wl)
            pk = vl = None
            npre = min(wl, 40 * L)
            spk, svl = F.sync_peaks_prefix_host(rf, planes['demod_sync'][b:b + npre], npre, cb_staging)
            pos = np.searchsorted(gpk, spk + b)
            common = np.nonzero(gpk[np.minimum(pos, len(gpk) - 1)] == spk + b)[0] if len(gpk) else np.zeros(0, dtype=np.int64)
            if len(common):
                k = int(common[0])
                m = int(pos[k])
                pk = np.concatenate([spk[:k], gpk[m:] - b])
                vl = np.concatenate([svl[:k], gvl[m:]])
                # a peak belongs to the window's list iff the step that found it started below the bound
                limit = wl - 2 * L
                i0 = np.concatenate([[0], pk[:-1] + skip])
                istep = i0 + ((pk - i0) // half) * half
                stop = np.nonzero(istep >= limit)[0]
                n = int(stop[0]) if len(stop) else len(pk)
                pk, vl = pk[:n], vl[:n]
            else:
                pk, vl = F.sync_peaks_device(rf, planes['demod_sync'][b:b + wl], wl, 0)
            keep['pk'], keep['vl'] = np.ascontiguousarray(pk, dtype=np.int64), np.ascontiguousarray(vl, dtype=np.float64)
            ppk[0] = keep['pk'].ctypes.data
            pvl[0] = keep['vl'].ctypes.data
            pn[0] = len(pk)
            return 0

        cb = _lib.WINDOW_PEAKS_FN(window_peaks)
        rf._check(be.lib.ldd_field_chain(rf._h, F._h(gpk), F._h(gvl), len(gpk), int(total), int(plane_origin), int(ncap_total),
                                         int(self.readlen), int(first_readsample), int(stop_readsample), int(bool(tolerant)), mf,
                                         C.cast(cb, C.c_void_p), None, C.cast(fields, C.c_void_p), F._h(batch.base),
                                         F._h(batch.winlen), F._h(readsample), F._h(batch.linelocs1.reshape(-1)),
                                         F._h(batch.linebad.reshape(-1)), F.LL_STRIDE, C.byref(nf)))
        n = nf.value
        infos = [fields[i] for i in range(n)]
        return batch, infos, readsample[:n].copy()

    # -- host copies
    def pictures(self, res):
        """uint16 TBC fields of the located windows -> list of (readsample, istop, array | None)."""
        rf, be = self.rf, self.rf._be
        if res.d_pic is None:
            return []
        be.synchronize()
        pic = be.to_host(res.d_pic).reshape(len(res.located), res.out_stride)
        st = be.to_host(res.d_status)
        W = rf.SysParams['outlinelen']
        out = []
        for k, j in enumerate(res.located):
            ok = (st[k] & (1 | 2 | 4 | 8)) == 0
            info = res.infos[j]
            out.append((int(res.readsamples[j]), int(info.istop), pic[k, :info.linecount * W].copy() if ok else None))
        return out

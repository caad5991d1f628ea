"""Multi-GPU decode: one process per GPU, the capture sharded by contiguous read-position ranges.

Every rank decodes the fields whose read position falls in its range (pipeline.CaptureDecoder.
decode_range: own block range + halos, no data-path collective); the only exchange is the gather
of the per-field outputs (uint16 TBC fields, read positions, parity, status) to rank 0, over NCCL
on NVLink when the backend is CUDA (gloo in the CPU tests).  SURVEY.md section 8e.
"""
import ctypes as C

import numpy as np

from . import _lib


def shard_bounds(ncap, world):
    """Read-position ranges [r0, r1) per rank; the last range is open-ended."""
    edges = [(g * ncap) // world for g in range(world)] + [ncap + 1]
    return [(edges[g], edges[g + 1]) for g in range(world)]


def needed_window(cd, ncap, r0, r1):
    """Capture samples [lo, hi) a rank must hold to decode its range (block range + halos)."""
    first_block, nblocks, _ = cd.plan_range(ncap, r0, r1)
    S, N = cd.stride, cd.rf.blocklen
    hi = min(ncap, first_block + max(nblocks - 1, 0) * S + N)
    return first_block, hi


class FieldGatherer:
    """Gathers the located fields of every rank on rank 0: ONE asynchronous collective per step.

    Send buffers are double-buffered and handed to the decoder as the TBC kernel's destination (buffers()), so a
    step's pictures are written straight into the buffer the collective sends -- no staging copy -- and the
    collective of step k only has to be over before step k+2 writes the same buffer: ranks are never
    stream-ordered behind the slowest rank's previous step.  Per-field metadata (read position, parity, line
    count) rides in the head of the same buffer.  to_host() turns the gathered buffers of the LAST gather() into a
    list of (readsample, istop, picture | None) ordered by read position."""

    NBUF = 2

    def __init__(self, cd, rank, world, max_fields, dist=None):
        import torch
        self.torch = torch
        self.cd, self.rank, self.world, self.max_fields, self.dist = cd, rank, world, max_fields, dist
        rf = cd.rf
        self.be = rf._be
        self.W = rf.SysParams['outlinelen']
        self.stride = (rf.SysParams['frame_lines'] // 2 + 1) * self.W
        self.cuda = self.be.name == "cuda"
        dev = self.be.device if self.cuda else "cpu"
        F = max_fields
        self.meta_bytes = F * 3 * 8                               # int64 [F][3]: readsample, istop, linecount
        self.status_off = self.meta_bytes                         # int32 [F]: the kernels' error bits
        self.pic_off = (self.status_off + 4 * F + 255) // 256 * 256
        nbytes = self.pic_off + F * self.stride * 2
        self.send = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(self.NBUF)]
        self.recv = None
        if rank == 0 and world > 1:
            self.recv = [[torch.empty_like(self.send[0]) for _ in range(world)] for _ in range(self.NBUF)]
        # host staging of the metadata: page-locked on CUDA (uploaded by a kernel, not the copy engine)
        self._hmeta = [self.be.pinned(F * 3, np.int64) for _ in range(self.NBUF)] if self.cuda else None
        self._n = 0
        self._work = [None] * self.NBUF
        self._last = None

    def _views(self, b):
        torch = self.torch
        F = self.max_fields
        meta = b[:self.meta_bytes].view(torch.int64).view(F, 3)
        status = b[self.status_off:self.status_off + 4 * F].view(torch.int32)
        pic = b[self.pic_off:].view(torch.uint16)
        return meta, status, pic

    def buffers(self, stream=None):
        """(pictures uint16 [max_fields * stride], status int32 [max_fields]) of the next gather(): pass them to the
        decoder as pic_out / status_out.  Orders `stream` (default: the current stream) -- the stream whose kernels will
        write the buffer -- behind the collective that last sent it."""
        k = self._n % self.NBUF
        w = self._work[k]
        if w is not None:
            if stream is not None and self.cuda:
                with self.be.stream_ctx(stream):
                    w.wait()
            else:
                w.wait()
            self._work[k] = None
        _, status, pic = self._views(self.send[k])
        return pic, status

    def gather(self, res):
        """res: the RangeResult whose pictures were written into buffers() (or any RangeResult: its pictures are then
        copied into the send buffer)."""
        torch = self.torch
        k = self._n % self.NBUF
        pic, status = self.buffers()
        self._n += 1
        buf = self.send[k]
        nloc = len(res.located)
        if nloc > self.max_fields:
            raise ValueError("max_fields too small")
        meta = self.be.host_view(self._hmeta[k]).reshape(self.max_fields, 3) if self.cuda else np.empty((self.max_fields, 3), dtype=np.int64)
        meta[:] = -1
        if nloc:
            loc = np.asarray(res.located, dtype=np.intp)
            infos = res.infos[loc]
            meta[:nloc, 0] = np.asarray(res.readsamples)[loc]
            meta[:nloc, 1] = infos['istop']
            meta[:nloc, 2] = infos['linecount']
            in_place = res.d_pic is not None and hasattr(res.d_pic, "data_ptr") and res.d_pic.data_ptr() == pic.data_ptr()
            if not in_place:
                # decoded into the decoder's own buffer: stage it
                n = nloc * self.stride
                if self.cuda:
                    pic.view(torch.uint8)[:2 * n] = res.d_pic[:n].view(torch.uint8)
                    status[:nloc] = res.d_status[:nloc]
                else:
                    pic[:n] = torch.from_numpy(np.ascontiguousarray(self.be.to_host(res.d_pic)[:n]).view(np.int16)).view(torch.uint16)
                    status[:nloc] = torch.from_numpy(np.asarray(self.be.to_host(res.d_status)[:nloc], dtype=np.int32))
        if self.cuda:
            be = self.be
            be.lib.ldd_copy_small(be.ptr(buf), be.ptr(self._hmeta[k]), self.meta_bytes, be.stream())
        else:
            buf[:self.meta_bytes] = torch.from_numpy(meta.reshape(-1).view(np.uint8))
        self._last = k
        if self.world > 1:
            # asynchronous: the collective runs on the process group's own stream behind the work enqueued above,
            # so the next range's demodulation overlaps it
            self._work[k] = self.dist.gather(buf, self.recv[k] if self.rank == 0 else None, dst=0, async_op=True)

    def wait(self):
        """Order the current stream behind every outstanding gather."""
        for k, w in enumerate(self._work):
            if w is not None:
                w.wait()
                self._work[k] = None

    def to_host(self):
        self.wait()
        if self.rank != 0 or self._last is None:
            return None
        bufs = self.recv[self._last] if self.recv is not None else [self.send[self._last]]
        out = []
        for b in bufs:
            meta, status, pic = self._views(b)
            m = meta.cpu().numpy()
            st = status.cpu().numpy()
            p = pic.cpu().view(self.torch.int16).numpy().view(np.uint16).reshape(self.max_fields, self.stride)
            for k in range(self.max_fields):
                if m[k, 0] < 0:
                    break
                ok = (st[k] & 15) == 0
                out.append((int(m[k, 0]), int(m[k, 1]), p[k, :m[k, 2] * self.W].copy() if ok else None))
        out.sort(key=lambda t: t[0])
        return out


class RawBuf:
    """A typed view of device memory that torch does not own (here: the root rank's gather buffer, mapped through
    CUDA IPC).  Just enough of a tensor's surface for the decoder: data_ptr(), len(), slicing, a host copy."""

    def __init__(self, lib, ptr, n, dtype):
        self.lib, self.ptr, self.n, self.dtype = lib, int(ptr), int(n), np.dtype(dtype)

    def data_ptr(self):
        return self.ptr

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        a, b, st = sl.indices(self.n)
        assert st == 1
        return RawBuf(self.lib, self.ptr + a * self.dtype.itemsize, max(b - a, 0), self.dtype)

    def to_host(self):
        out = np.empty(self.n, dtype=self.dtype)
        if self.n:
            rc = self.lib.ldd_peer_read(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), out.nbytes)
            if rc:
                raise RuntimeError("ldd_peer_read failed (%d)" % rc)
        return out

    def cpu(self):
        import torch
        return torch.from_numpy(self.to_host())


class PeerGatherer:
    """The gather of per-field outputs to rank 0 WITHOUT a collective: rank 0 publishes its receive buffer (CUDA IPC) and
    every rank maps it.  Two ways of filling it:

    * push=False ("p2p"): the decoder gets an address inside the mapping as the TBC kernel's destination (buffers()), so
      the resampling kernel of rank g stores its uint16 fields straight into rank 0's HBM over NVLink -- compute and
      transfer are one kernel and there is no staging copy.  Measured on 8 B200s this LOSES to the NCCL gather (170 vs
      185 Gsamples/s): seven ranks' remote stores keep their TBC CTAs resident ~3x longer, SMs the next range's
      demodulation wants.
    * push=True ("push"): the TBC kernel writes a local send buffer (double-buffered, as with NCCL) and ONE DMA transfer
      per step (ldd_peer_copy on a copy stream: copy engine over NVLink) moves the used part of it into the rank's slot
      of rank 0's buffer -- no SM of either GPU takes part in the transfer.

    Per-field metadata and a per-rank step flag follow on the same stream (ldd_copy_small, ldd_peer_signal).  Flow
    control: NBUF slots per rank; before slot k is written for step s the writer waits (on the device) until rank 0 has
    released step s - NBUF of that slot, which it does when it moves on to step s itself (or when its consumer calls
    to_host()).  Same interface as FieldGatherer."""

    NBUF = 2
    HEADER = 4096

    def __init__(self, cd, rank, world, max_fields, dist, push=False):
        import torch
        self.torch = torch
        self.cd, self.rank, self.world, self.max_fields, self.dist = cd, rank, world, max_fields, dist
        self.push = bool(push)
        rf = cd.rf
        self.be = be = rf._be
        self.lib = be.lib
        self.W = rf.SysParams['outlinelen']
        self.stride = (rf.SysParams['frame_lines'] // 2 + 1) * self.W
        F = max_fields
        self.meta_bytes = F * 3 * 8
        self.status_off = self.meta_bytes
        self.pic_off = (self.status_off + 4 * F + 255) // 256 * 256
        self.slot_bytes = (self.pic_off + F * self.stride * 2 + 255) // 256 * 256
        total = self.HEADER + self.NBUF * world * self.slot_bytes
        handle = (C.c_ubyte * 64)()
        base = C.c_void_p()
        rc = 0
        if rank == 0:
            rc = self.lib.ldd_peer_alloc(total, C.byref(base), handle)
        box = [bytes(handle) if rc == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            rc = -1
            if box[0] is not None:
                hb = (C.c_ubyte * 64).from_buffer_copy(box[0])
                rc = self.lib.ldd_peer_open(hb, C.byref(base))
        # every rank learns whether every rank has the mapping (so that nobody is left waiting in a collective)
        ok = torch.tensor([1 if rc == 0 else 0], device=be.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.base = base.value if rc == 0 else 0
        if int(ok[0]) != 1:
            self.close(collective=False)
            raise RuntimeError("peer mapping of rank 0's gather buffer failed on some rank")
        self._hmeta = [be.pinned(F * 3, np.int64) for _ in range(self.NBUF)]
        self._step = 0
        self._last = None
        self._cache = None
        if self.push:
            self.local = [torch.zeros(self.slot_bytes, dtype=torch.uint8, device=be.device) for _ in range(self.NBUF)]
            self.copy_stream = be.new_stream()
            self.copy_done = [None] * self.NBUF

    # header: ack[k] at 4 k ; flags[k][r] at 256 + 4 (k world + r)
    def _ack(self, k):
        return self.base + 4 * k

    def _flag(self, k, r):
        return self.base + 256 + 4 * (k * self.world + r)

    def _slot(self, k, r):
        return self.base + self.HEADER + (k * self.world + r) * self.slot_bytes

    def buffers(self, stream=None):
        """(pictures, status) of the next gather() inside rank 0's buffer; orders `stream` behind rank 0's release of
        the slot's previous contents."""
        be = self.be
        s = self._step + 1
        if self._cache is not None and self._cache[0] == s:
            return self._cache[1], self._cache[2]
        k = s % self.NBUF
        st = C.c_void_p(stream.cuda_stream) if stream is not None else be.stream()
        if self.rank == 0:
            if s > self.NBUF:
                self.lib.ldd_peer_signal(C.c_void_p(self._ack(k)), s - self.NBUF, st)       # slot k's old contents may go
        elif s > self.NBUF and not self.push:
            self.lib.ldd_peer_wait(C.c_void_p(self._ack(k)), 1, 1, s - self.NBUF, st)
        if self.push:
            # the kernels write the local buffer; it is free again when its previous transfer is over
            if self.copy_done[k] is not None:
                be.stream_wait_event(stream if stream is not None else be.current_stream_obj(), self.copy_done[k])
                self.copy_done[k] = None
            b = self.local[k]
            pic = b[self.pic_off:self.pic_off + 2 * self.max_fields * self.stride].view(self.torch.uint16)
            status = b[self.status_off:self.status_off + 4 * self.max_fields].view(self.torch.int32)
            self._cache = (s, pic, status)
            return pic, status
        slot = self._slot(k, self.rank)
        pic = RawBuf(self.lib, slot + self.pic_off, self.max_fields * self.stride, np.uint16)
        status = RawBuf(self.lib, slot + self.status_off, self.max_fields, np.int32)
        self._cache = (s, pic, status)
        return pic, status

    def gather(self, res):
        be = self.be
        pic, status = self.buffers()
        self._step += 1
        s = self._step
        k = s % self.NBUF
        nloc = len(res.located)
        if nloc > self.max_fields:
            raise ValueError("max_fields too small")
        meta = be.host_view(self._hmeta[k]).reshape(self.max_fields, 3)
        meta[:] = -1
        if nloc:
            loc = np.asarray(res.located, dtype=np.intp)
            infos = res.infos[loc]
            meta[:nloc, 0] = np.asarray(res.readsamples)[loc]
            meta[:nloc, 1] = infos['istop']
            meta[:nloc, 2] = infos['linecount']
            if res.d_pic.data_ptr() != pic.data_ptr():
                # decoded into the decoder's own buffer: copy it over (the decode_stream(sink=...) path never does)
                n = nloc * self.stride
                self.lib.ldd_copy_small(be.ptr(pic), be.ptr(res.d_pic), 2 * n, be.stream())
                self.lib.ldd_copy_small(be.ptr(status), be.ptr(res.d_status), 4 * nloc, be.stream())
        slot = self._slot(k, self.rank)
        if self.push:
            self.lib.ldd_copy_small(be.ptr(self.local[k]), be.ptr(self._hmeta[k]), self.meta_bytes, be.stream())
            ready = be.record_event()
            be.stream_wait_event(self.copy_stream, ready)
            cs = C.c_void_p(self.copy_stream.cuda_stream)
            if self.rank != 0 and s > self.NBUF:
                self.lib.ldd_peer_wait(C.c_void_p(self._ack(k)), 1, 1, s - self.NBUF, cs)
            rc = self.lib.ldd_peer_copy(C.c_void_p(slot), be.ptr(self.local[k]), self.pic_off + 2 * nloc * self.stride, cs)
            if rc:
                raise RuntimeError("ldd_peer_copy failed (%d)" % rc)
            self.lib.ldd_peer_signal(C.c_void_p(self._flag(k, self.rank)), s, cs)
            with be.stream_ctx(self.copy_stream):
                self.copy_done[k] = be.record_event()
        else:
            self.lib.ldd_copy_small(C.c_void_p(slot), be.ptr(self._hmeta[k]), self.meta_bytes, be.stream())
            self.lib.ldd_peer_signal(C.c_void_p(self._flag(k, self.rank)), s, be.stream())
        self._last = (k, s)

    def wait(self):
        """Rank 0: orders the current stream behind the arrival of every rank's last step."""
        if self.rank == 0 and self._last is not None:
            k, s = self._last
            self.lib.ldd_peer_wait(C.c_void_p(self._flag(k, 0)), self.world, 1, s, self.be.stream())

    def to_host(self):
        """Collective over the ranks: rank 0 returns the fields of the last gather(), ordered by read position."""
        self.wait()
        self.be.synchronize()
        out = None
        if self.rank == 0 and self._last is not None:
            k, s = self._last
            out = []
            F = self.max_fields
            for r in range(self.world):
                raw = RawBuf(self.lib, self._slot(k, r), self.slot_bytes, np.uint8).to_host()
                m = raw[:self.meta_bytes].view(np.int64).reshape(F, 3)
                st = raw[self.status_off:self.status_off + 4 * F].view(np.int32)
                p = raw[self.pic_off:self.pic_off + 2 * F * self.stride].view(np.uint16).reshape(F, self.stride)
                for i in range(F):
                    if m[i, 0] < 0:
                        break
                    ok = (st[i] & 15) == 0
                    out.append((int(m[i, 0]), int(m[i, 1]), p[i, :m[i, 2] * self.W].copy() if ok else None))
            out.sort(key=lambda t: t[0])
        self.dist.barrier()
        return out

    def close(self, collective=True):
        if collective:
            self.dist.barrier()
        if self.base:
            if self.rank == 0:
                self.be.synchronize()
                self.lib.ldd_peer_free(C.c_void_p(self.base))
            else:
                self.be.synchronize()
                self.lib.ldd_peer_close(C.c_void_p(self.base))
            self.base = 0


DEFAULT_GATHER = "push"      # measured on 8 B200s (PAL 1 s per GPU): push 225, NCCL 201, kernel stores through the mapping 170 Gsamples/s


def make_gatherer(cd, rank, world, max_fields, dist, mode=None):
    """FieldGatherer (NCCL gather) or PeerGatherer (rank 0's buffer mapped by every rank; CUDA only).
    mode: 'nccl' | 'push' (one DMA transfer per step into the mapping) | 'p2p' (TBC kernels store through the mapping) |
    None = environment LDD_GATHER, default DEFAULT_GATHER; the peer modes fall back to NCCL when the mapping fails."""
    import os
    mode = mode or os.environ.get("LDD_GATHER", DEFAULT_GATHER)
    if mode in ("p2p", "push") and cd.rf._be.name == "cuda" and world > 1:
        try:
            return PeerGatherer(cd, rank, world, max_fields, dist, push=(mode == "push"))   # fails on every rank together or on none
        except RuntimeError:
            pass
    return FieldGatherer(cd, rank, world, max_fields, dist)


def gather_fields(cd, res, rank, world, max_fields, dist=None):
    """One-shot convenience wrapper: gather and return the host list on rank 0 (None elsewhere)."""
    g = FieldGatherer(cd, rank, world, max_fields, dist)
    g.gather(res)
    return g.to_host()


def pcm_chain(rf, linecounts, istops, audio_offset=0.0, frame_state=2, chain='framer', freq=48000.0):
    """ldd_pcm_chain: run the PCM time-offset chain over fields given by (linecount, istop) -> (samples per field,
    audio_offset, frame_state) after the last of them."""
    n = len(linecounts)
    lc = (C.c_int * max(n, 1))(*[int(x) for x in linecounts])
    tp = (C.c_int * max(n, 1))(*[int(x) for x in istops])
    nout = (C.c_int * max(n, 1))()
    ao, fs = C.c_double(float(audio_offset)), C.c_int(int(frame_state))
    rf._check(rf._be.lib.ldd_pcm_chain(_lib.SYSTEM[rf.system], float(freq),
                                       float(rf.SysParams['line_period']), _lib.PCM_CHAIN_FRAMER if chain == 'framer' else _lib.PCM_CHAIN_FIELDS,
                                       n, lc, tp, C.byref(ao), C.byref(fs), nout))
    return [int(nout[k]) for k in range(n)], ao.value, fs.value


def sharded_pcm(cd, res, rank, world, dist, chain='framer', freq=48000.0, scale=64):
    """48 kHz PCM of a capture sharded over the ranks (one read-position range each, `res` = this rank's decode_range with
    the analog audio on).  The time offset a field starts with is a prefix over the line counts and parities of all fields
    before it, so the ranks first exchange those few integers (all_gather_object), every rank
    runs the chain over the fields of the ranks before it (ldd_pcm_chain, host) and resamples its own fields from its
    own audio (ldd_pipe_pcm); the samples are then gathered on rank 0.  Returns on rank 0 the list of
    (readsample, int16 L/R array | None) ordered by read position, None elsewhere."""
    rf = cd.rf
    mine = [(int(res.infos[j].linecount), int(res.infos[j].istop)) for j in res.located]
    every = [None] * world
    if world > 1:
        dist.all_gather_object(every, mine)
    else:
        every[0] = mine
    before = [f for r in range(rank) for f in every[r]]
    _, offset, state = pcm_chain(rf, [f[0] for f in before], [f[1] for f in before], 0.0, 2, chain, freq)
    pcm, _, _ = cd.pcm(res, offset, state, chain=chain, freq=freq, scale=scale)
    out = [(int(res.readsamples[j]), p) for j, p in zip(res.located, pcm)]
    if world == 1:
        return out
    parts = [None] * world if rank == 0 else None
    dist.gather_object(out, parts, dst=0)
    if rank != 0:
        return None
    return sorted([x for part in parts for x in part], key=lambda x: x[0])

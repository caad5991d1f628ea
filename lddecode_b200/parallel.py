"""Multi-GPU decode: one process per GPU, the capture sharded by contiguous read-position ranges.

Every rank decodes the fields whose read position falls in its range (pipeline.CaptureDecoder.
decode_range: own block range + halos, no data-path collective); the only exchange is the gather
of the per-field outputs (uint16 TBC fields, read positions, parity, status) to rank 0, over NCCL
on NVLink when the backend is CUDA (gloo in the CPU tests).  SURVEY.md section 8e.
"""
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


def gather_fields(cd, res, rank, world, max_fields, dist=None):
    """One-shot convenience wrapper: gather and return the host list on rank 0 (None elsewhere)."""
    g = FieldGatherer(cd, rank, world, max_fields, dist)
    g.gather(res)
    return g.to_host()

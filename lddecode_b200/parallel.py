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
    """Gathers the located fields of every rank on rank 0.  Buffers are allocated once; gather()
    only enqueues device work (a copy into the send buffer and ONE asynchronous collective: the per-field
    metadata rides in the head of the same buffer as the pictures), to_host()
    turns the gathered buffers into a list of (readsample, istop, picture | None) ordered by read
    position; call it before the next gather() overwrites them."""

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
        self.meta_bytes = max_fields * 4 * 8                     # int64 [max_fields][4]: readsample, istop, linecount, status
        self.buf = torch.zeros(self.meta_bytes + max_fields * self.stride * 2, dtype=torch.uint8, device=dev)

        def split(b):
            return b[:self.meta_bytes].view(torch.int64).view(max_fields, 4), b[self.meta_bytes:]      # (meta, uint16 pictures as bytes)
        self.meta, self.pic = split(self.buf)
        if rank == 0 and world > 1:
            self.bufs = [torch.empty_like(self.buf) for _ in range(world)]
            parts = [split(b) for b in self.bufs]
            self.metas, self.pics = [p[0] for p in parts], [p[1] for p in parts]
        else:
            self.bufs = None
            self.pics, self.metas = [self.pic], [self.meta]
        # host staging of the metadata: page-locked and double-buffered on CUDA (uploaded by a kernel, not the copy engine)
        self._hmeta = [self.be.pinned(max_fields * 4, np.int64) for _ in range(2)] if self.cuda else None
        self._n = 0
        self._work = []

    def gather(self, results):
        """results: one RangeResult or the list decode_pipelined returns."""
        torch = self.torch
        if not isinstance(results, (list, tuple)):
            results = [results]
        if self.cuda:
            hm = self._hmeta[self._n % 2]
            self._n += 1
            meta = self.be.host_view(hm).reshape(self.max_fields, 4)
            meta[:] = -1
        else:
            meta = np.full((self.max_fields, 4), -1, dtype=np.int64)
        k0 = 0
        spans = []
        for res in results:
            nloc = len(res.located)
            if k0 + nloc > self.max_fields:
                raise ValueError("max_fields too small")
            if nloc:
                loc = np.asarray(res.located, dtype=np.intp)
                meta[k0:k0 + nloc, 0] = np.asarray(res.readsamples)[loc]
                meta[k0:k0 + nloc, 1] = np.fromiter((res.infos[j].istop for j in res.located), dtype=np.int64, count=nloc)
                meta[k0:k0 + nloc, 2] = np.fromiter((res.infos[j].linecount for j in res.located), dtype=np.int64, count=nloc)
                meta[k0:k0 + nloc, 3] = 0
            spans.append((k0, nloc, res))
            k0 += nloc
        self.wait()
        if self.cuda:
            be = self.be
            be.lib.ldd_copy_small(be.ptr(self.buf), be.ptr(hm), self.meta_bytes, be.stream())
            for k0, nloc, res in spans:
                if nloc:
                    self.meta[k0:k0 + nloc, 3] = res.d_status[:nloc].to(torch.int64)
                    self.pic[k0 * self.stride * 2:(k0 + nloc) * self.stride * 2] = res.d_pic[:nloc * self.stride].view(torch.uint8)
        else:
            tm = torch.from_numpy(meta)
            for k0, nloc, res in spans:
                if nloc:
                    tm[k0:k0 + nloc, 3] = torch.from_numpy(np.asarray(self.be.to_host(res.d_status)[:nloc], dtype=np.int64))
                    self.pic[k0 * self.stride * 2:(k0 + nloc) * self.stride * 2] = \
                        torch.from_numpy(self.be.to_host(res.d_pic)[:nloc * self.stride].view(np.uint8))
            self.meta.copy_(tm)
        if self.world > 1:
            # asynchronous: the collective runs on the process group's own stream behind the copies
            # above, so the next range's demodulation overlaps it; wait() orders the caller's
            # stream (not the host) behind it
            self._work = [self.dist.gather(self.buf, self.bufs if self.rank == 0 else None, dst=0, async_op=True)]

    def wait(self):
        """Order the current stream behind the last gather (called before the send buffers are rewritten
        and before the gathered buffers are read)."""
        for w in self._work:
            w.wait()
        self._work = []

    def to_host(self):
        self.wait()
        if self.rank != 0:
            return None
        out = []
        for g in range(len(self.pics)):
            m = self.metas[g].cpu().numpy()
            p = self.pics[g].cpu().numpy().view(np.uint16).reshape(self.max_fields, self.stride)
            for k in range(self.max_fields):
                if m[k, 0] < 0:
                    break
                ok = (m[k, 3] & 15) == 0
                out.append((int(m[k, 0]), int(m[k, 1]), p[k, :m[k, 2] * self.W].copy() if ok else None))
        out.sort(key=lambda t: t[0])
        return out


def gather_fields(cd, res, rank, world, max_fields, dist=None):
    """One-shot convenience wrapper: gather and return the host list on rank 0 (None elsewhere)."""
    g = FieldGatherer(cd, rank, world, max_fields, dist)
    g.gather(res)
    return g.to_host()

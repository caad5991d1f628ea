"""Multi-GPU decode: one process per GPU, the capture sharded by contiguous read-position ranges.

Every rank decodes the fields whose read position falls in its range (pipeline.CaptureDecoder.
decode_range: own block range + halos, no data-path collective); the only exchange is the gather
of the per-field outputs (uint16 TBC fields, read positions, parity) to rank 0, over NCCL on
NVLink when the backend is CUDA (gloo in the CPU tests).  SURVEY.md section 8e.
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


def gather_fields(cd, res, rank, world, max_fields, dist=None):
    """Gather (readsample, istop, linecount, ok) and the uint16 pictures of every rank on rank 0.

    Returns on rank 0 a list of (readsample, istop, picture) ordered by read position; None elsewhere."""
    rf, be = cd.rf, cd.rf._be
    torch = None
    W = rf.SysParams['outlinelen']
    stride = res.out_stride
    nloc = len(res.located)
    if nloc > max_fields:
        raise ValueError("max_fields too small")
    meta = np.zeros((max_fields, 4), dtype=np.int64)
    st = be.to_host(res.d_status) if nloc else np.zeros(0, dtype=np.int32)
    for k, j in enumerate(res.located):
        info = res.infos[j]
        meta[k] = (res.readsamples[j], info.istop, info.linecount, int((st[k] & 15) == 0))
    meta[nloc:, 0] = -1
    import torch
    if be.name == "cuda":
        dev = be.device
        pic = torch.zeros(max_fields * stride, dtype=torch.uint16, device=dev)
        if nloc:
            pic[:nloc * stride] = res.d_pic[:nloc * stride]
        tmeta = torch.from_numpy(meta).to(dev)
    else:
        pic = torch.zeros(max_fields * stride, dtype=torch.int16)
        if nloc:
            pic[:nloc * stride] = torch.from_numpy(be.to_host(res.d_pic)[:nloc * stride].view(np.int16))
        tmeta = torch.from_numpy(meta)
    if world == 1:
        pics, metas = [pic], [tmeta]
    else:
        pv = pic.view(torch.uint8)             # bytes: NCCL and gloo have no uint16
        pics = [torch.empty_like(pv) for _ in range(world)] if rank == 0 else None
        metas = [torch.empty_like(tmeta) for _ in range(world)] if rank == 0 else None
        dist.gather(pv, pics, dst=0)
        dist.gather(tmeta, metas, dst=0)
    if rank != 0:
        return None
    out = []
    for g in range(world):
        m = metas[g].cpu().numpy()
        p = pics[g].cpu().contiguous().view(torch.uint8).numpy().view(np.uint16).reshape(max_fields, stride)
        for k in range(max_fields):
            if m[k, 0] < 0:
                break
            out.append((int(m[k, 0]), int(m[k, 1]), p[k, :m[k, 2] * W].copy() if m[k, 3] else None))
    out.sort(key=lambda t: t[0])
    return out

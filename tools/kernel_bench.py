#!/usr/bin/env python3
"""Per-kernel timing and algorithmic HBM bandwidth on one GPU (CUDA events, inputs larger than L2),
for the roofline table in DESIGN.md / profiles.  Also the command that the per-kernel ncu captures wrap."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from lddecode_b200 import _lib, field as F, pipeline, rfdecode, synth  # noqa: E402


def timeit(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    peak = 6650.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    system = os.environ.get("SYSTEM", "PAL")
    out = []

    def rec(name, ms, nbytes, note=""):
        r = dict(kernel=name, ms=round(ms, 4), algorithmic_bytes=int(nbytes), gbs=round(nbytes / ms / 1e6, 1),
                 frac_of_measured_hbm=round(nbytes / ms / 1e6 / peak, 4), note=note)
        print(json.dumps(r), flush=True)
        out.append(r)

    # ---- kernel (1): unpack, 300 M samples of .r30 / .lds
    nwords = 100_000_000
    words = torch.randint(0, 2**31 - 1, (nwords,), dtype=torch.int32, device="cuda")
    o16 = torch.empty(nwords * 3, dtype=torch.int16, device="cuda")
    lib = _lib.load()
    st = lambda: torch.cuda.current_stream().cuda_stream
    ms = timeit(lambda: lib.ldd_unpack_r30_ddunpack(words.data_ptr(), nwords, o16.data_ptr(), st()))
    rec("unpack_r30_dd_kernel (.r30 -> int16, ddunpack.c)", ms, nwords * 4 + nwords * 6, "4 B in + 6 B out per 3 samples")
    n = 300_000_000
    of = torch.empty(n, dtype=torch.float32, device="cuda")
    ms = timeit(lambda: lib.ldd_unpack_f32(words.data_ptr(), _lib.FMT_R30, 0, n, of.data_ptr(), st()))
    rec("unpack_f32_kernel (.r30 -> float32)", ms, n * 4 // 3 + n * 4, "1.33 B in + 4 B out per sample")
    lds = torch.randint(0, 255, (n * 5 // 4,), dtype=torch.uint8, device="cuda")
    ou = torch.empty(n, dtype=torch.int16, device="cuda")
    ms = timeit(lambda: lib.ldd_unpack_raw(lds.data_ptr(), _lib.FMT_LDS40, 0, n, ou.data_ptr(), st()))
    rec("unpack_raw_kernel (.lds -> uint16, load_packed_data_4_40)", ms, n * 5 // 4 + n * 2, "1.25 B in + 2 B out per sample")
    del words, o16, of, lds, ou

    # ---- the pipeline kernels on a 1 s capture
    ncap = bench.one_second(system) + bench.TAIL
    cap = bench.synth_capture(system, ncap, 1)
    cap_dev = torch.from_numpy(cap).cuda()
    for prec in ("f64", "f32", "mixed"):
        rf = rfdecode.RFDecode(bench.FS[system], system, 16384, decode_analog_audio=False, precision=prec)
        cd = pipeline.CaptureDecoder(rf, max_fields=256)
        total = bench.demod_only(cd, cap_dev, _lib.FMT_U8, ncap)
        ms = timeit(lambda: bench.demod_only(cd, cap_dev, _lib.FMT_U8, ncap))
        nplanes32 = 4 if system == "PAL" else 3
        bps = 16384 / 15328 + 4 * nplanes32 + 8
        rec("demod_kernel<%s> (fused block demodulation, N=16384)" % prec, ms, bps * total, "%.2f B per sample; %.0f Msamples/s" % (bps, total / ms / 1e3))
        if prec != "f64":
            del rf, cd
    rf = rfdecode.RFDecode(bench.FS[system], system, 16384, decode_analog_audio=False)
    cd = pipeline.CaptureDecoder(rf, max_fields=256)
    res = cd.decode(cap_dev, _lib.FMT_U8, ncap, want_tables=False)
    torch.cuda.synchronize()
    planes, total = res.planes, res.plane_len
    be = rf._be
    stg = {}
    ms = timeit(lambda: F.sync_peaks_launch(rf, planes['demod_sync'], total, 0, stg))
    rec("peak chase (peaks_phase1 + merge + scan + copy + list to pinned host memory)", ms, total * 8, "reads the float64 sync plane once")
    # refine + tbc on the located fields
    n = len(res.located)
    W = rf.SysParams['outlinelen']
    lines = sum(res.infos[j].linecount for j in res.located)
    sub = F.FieldBatch(rf, n)
    for k, j in enumerate(res.located):
        sub.base[k], sub.winlen[k], sub.linecount[k] = res.base[j], 1001026, res.infos[j].linecount
        sub.linelocs1[k] = res.linelocs1[j]
    d = {k: be.to_device(getattr(sub, k).reshape(-1)) for k in ("base", "winlen", "linecount", "linelocs1", "linebad")}
    d_l2 = be.empty(n * F.LL_STRIDE, np.float64); d_b2 = be.empty(n * F.LL_STRIDE, np.uint8); d_st = be.zeros(n, np.int32)
    ms = timeit(lambda: lib.ldd_refine_hsync(rf._h, be.ptr(planes['demod_05']), total, be.ptr(d['base']), be.ptr(d['winlen']), be.ptr(d['linecount']),
                                             n, F.LL_STRIDE, be.ptr(d['linelocs1']), be.ptr(d['linebad']), be.ptr(d_l2), be.ptr(d_b2), be.ptr(d_st), be.stream()))
    rec("refine_hsync_kernel + fix-up", ms, lines * 4 * 650, "~650 demod_05 samples per line")
    d_l3 = be.empty(n * F.LL_STRIDE, np.float64)
    if system == "PAL":
        ms = timeit(lambda: lib.ldd_refine_pilot(rf._h, be.ptr(planes['demod']), be.ptr(planes['demod_05']), total, be.ptr(d['base']), be.ptr(d['linecount']),
                                                 n, F.LL_STRIDE, be.ptr(d_l2), be.ptr(d_l3), be.ptr(d_st), be.stream()))
        rec("pilot_lines_kernel + pilot_median_kernel", ms, lines * 8 * 167, "167 samples of two planes per line")
    else:
        d_bl = be.empty(n * F.LL_STRIDE, np.float32)
        ms = timeit(lambda: lib.ldd_refine_burst(rf._h, be.ptr(planes['demod_burst']), total, be.ptr(d['base']), be.ptr(d['linecount']),
                                                 n, F.LL_STRIDE, be.ptr(d_l2), be.ptr(d_l3), be.ptr(d_bl), be.ptr(d_st), be.stream()))
        rec("refine_burst_kernel (one pass)", ms, lines * 4 * 150, "~150 burst-plane samples per line")
    d_pic = be.empty(n * res.out_stride, np.uint16)
    ms = timeit(lambda: lib.ldd_tbc_fields(rf._h, be.ptr(planes['demod']), total, float(rf.SysParams['ire0']), be.ptr(d['base']), be.ptr(d_l3), F.LL_STRIDE,
                                           be.ptr(d['linecount']), n, int(sub.linecount.max()), 3 if system == "PAL" else 1, 0.0, W, 1, 1, be.ptr(d_pic),
                                           res.out_stride, None, 1.45, be.ptr(d_st), be.stream()))
    import ctypes
    rec("tbc_f32_kernel (not-a-knot spline resample -> uint16, bulk-copy staged, default lane)", ms, lines * (rf.linelen * 4 + W * 2), "4 B per input sample of the line + 2 B per output sample")
    os.environ["LDD_TBC_F64"] = "1"
    ms = timeit(lambda: lib.ldd_tbc_fields(rf._h, be.ptr(planes['demod']), total, float(rf.SysParams['ire0']), be.ptr(d['base']), be.ptr(d_l3), F.LL_STRIDE,
                                           be.ptr(d['linecount']), n, int(sub.linecount.max()), 3 if system == "PAL" else 1, 0.0, W, 1, 1, be.ptr(d_pic),
                                           res.out_stride, None, 1.45, be.ptr(d_st), be.stream()))
    del os.environ["LDD_TBC_F64"]
    rec("tbc_kernel (float64 spline, exact lane)", ms, lines * (rf.linelen * 4 + W * 2), "4 B per input sample of the line + 2 B per output sample")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_bench_%s.json" % system), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()

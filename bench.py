#!/usr/bin/env python3
"""bench.py -- RF Msamples/s demodulated + TBC on B200 (BASELINE.json metric), and the CPU arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--system PAL|NTSC] [--audio]

A step is one pass of the hot path (demodulate -> sync peaks -> locate fields -> refine -> TBC to
uint16) over one synthetic capture.  Workload at N=1: BASELINE.json configs[1], "PAL synthetic
8-bit RF 1 s (8fsc PAL) video demod + TBC on 1xB200"; for N>1 every rank decodes its own one-second
shard (weak scaling, no data-path collective) and the per-field outputs are gathered to rank 0 over
NCCL inside the timed region.

Printed JSON (one line, rank 0): value = whole-job Msamples/s with the capture resident in HBM;
e2e = the same through the public API from pinned host memory with H2D of the capture and D2H of the
uint16 fields inside the timed region; roofline = the dominant kernel (fused block demodulation)
against the measured HBM peak; cpu_baseline = the oracle port of the reference timed on this box.
--impl reference times the reference's CPU algorithm (oracle port: same numpy/scipy calls) on all
host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = {"NTSC": 8 * 315 / 88, "PAL": 35.46895}
BLOCKLEN = 16384                     # the reference's default blocklen_ (lddecode_core.py:120)
TAIL = 1100000                       # so the last 1e6-sample read succeeds


def one_second(system):
    return int(round(FS[system] * 1e6))


def synth_capture(system, n, seed, bits=8):
    """Seeded synthetic capture, cached under .bench_cache/ (generation is ~1.2 s per Msample)."""
    from lddecode_b200 import synth
    cdir = os.path.join(ROOT, ".bench_cache")
    path = os.path.join(cdir, "%s_%d_%d%s.npy" % (system, n, seed, "" if bits == 8 else "_%dbit" % bits))
    if os.path.exists(path):
        try:
            return np.load(path)
        except Exception:
            pass
    cap = synth.SynthRF(system, FS[system], seed=seed, bits=bits).generate(n)
    try:
        os.makedirs(cdir, exist_ok=True)
        tmp = path + ".%d.tmp.npy" % os.getpid()
        np.save(tmp, cap)
        os.replace(tmp, path)
    except Exception:
        pass
    return cap


# ---- CPU arm: the oracle port of the reference ------------------------------------------------------
def _cpu_decode_fields(args):
    """Framer.readfield's loop with the oracle on one capture: returns (samples consumed, seconds)."""
    system, audio, cap, nfields = args
    from oracle import ldd_oracle as O
    dec = O.Decoder(FS[system], system, BLOCKLEN, analog_audio=audio)
    ld = lambda s, n: cap[s:s + n] if s + n <= len(cap) else None
    t0 = time.perf_counter()
    readsample, done = 0, 0
    while done < nfields:
        d = O.demod(dec, ld, readsample, 1000000, 1)
        if d is None:
            break
        f = O.decode_field(dec, d[0], 0)
        readsample += f.nextfieldoffset
        done += 1
    return readsample, time.perf_counter() - t0


def cpu_baseline(system, audio, budget_s=12.0):
    """Single-core oracle port on a bounded sample (whole fields until ~budget_s of CPU work)."""
    nf = 6
    field = one_second(system) // (60 if system == "NTSC" else 50)
    cap = synth_capture(system, field * (nf + 3), 0)
    consumed, secs, fields = 0, 0.0, 0
    readsample = 0
    from oracle import ldd_oracle as O
    dec = O.Decoder(FS[system], system, BLOCKLEN, analog_audio=audio)
    ld = lambda s, n: cap[s:s + n] if s + n <= len(cap) else None
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        d = O.demod(dec, ld, readsample, 1000000, 1)
        if d is None:
            readsample = 0
            continue
        f = O.decode_field(dec, d[0], 0)
        consumed += f.nextfieldoffset
        readsample += f.nextfieldoffset
        fields += 1
    secs = time.perf_counter() - t0
    return dict(value=consumed / secs / 1e6, unit="Msamples/s", cores=1, kind="port",
                sample="%d %s fields (Framer.readfield loop: demod 1e6 + Field decode each) in %.1f s, oracle port of "
                       "lddecode_core on 1 core" % (fields, system, secs))


def run_reference(a):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    system, audio = a.system, a.audio
    cores = os.cpu_count() or 1
    field = one_second(system) // (60 if system == "NTSC" else 50)
    nfields = 2
    cap = synth_capture(system, field * (nfields + 3), 0)
    ctx = mp.get_context("fork")
    times = []
    consumed = 0
    with ctx.Pool(cores) as pool:
        for step in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            outs = pool.map(_cpu_decode_fields, [(system, audio, cap, nfields)] * cores)
            dt = time.perf_counter() - t0
            if step >= a.warmup:
                times.append(dt)
                consumed = sum(o[0] for o in outs)
    ms = 1e3 * sum(times) / len(times)
    value = consumed / (ms / 1e3) / 1e6
    sample = "%d processes x %d %s fields per step (identical captures; demod 1e6 + Field decode per field)" % (cores, nfields, system)
    line = dict(impl="reference", metric="rf_msamples_per_s_demod_tbc", value=value, unit="Msamples/s", n_gpus=a.gpus,
                steps=a.steps, warmup=a.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", config=workload_config(system, audio, a.gpus, a.fmt),
                cpu_baseline=dict(value=value, unit="Msamples/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="Msamples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                realtime_x=value / FS[system])
    print(json.dumps(line), flush=True)


def workload_config(system, audio, gpus, fmt="u8"):
    return dict(workload="%s synthetic %s RF, 1 s at 8fsc (%.3f MSPS) per GPU, %s demod + sync + TBC to uint16 4fsc"
                         % (system, FMT_NAMES[fmt], FS[system], "video + both analog audio channels" if audio else "video"),
                blocklen=BLOCKLEN, readlen=1000000, parallelism="block-range shards, one 1-s shard per GPU (x%d)" % gpus,
                l2="per step ~30-46 MB in + ~0.7-0.9 GB of planes written: working set exceeds the 126 MB L2, no flush needed")


# ---- clocks -----------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML queries every few ms (an
    `nvidia-smi -lms` child stalls the driver for milliseconds per sample, which is visible in a 2 ms step); falls back
    to nvidia-smi when pynvml is missing."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period_s=0.02):
        self.rows = []          # (time, sm_mhz, max_mhz, reasons bitmask)   [nvml]  or (time, csv line) [nvidia-smi]
        self.proc = None
        self.index = index
        self.period = period_s
        self.nvml = None
        self._stop = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(hi)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        h = hi
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml, self.h = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), sm, self.mx, r))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken between wall-clock times t0 and t1 (the timed region)."""
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml is not None:
            self._stop = True
            n = self.nvml
            bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                    n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
            rows = [r for r in self.rows if (t0 is None or r[0] >= t0 - 0.01) and (t1 is None or r[0] <= t1 + 0.01)]
            reasons = sorted({nm for r in rows for nm, bit in zip(names, bits) if r[3] & bit})
            sm = [r[1] for r in rows]
            return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=self.mx, reasons=reasons, samples=len(sm),
                        source="nvml")
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0 - 0.11) and (t1 is None or t <= t1 + 0.11)]
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")


# ---- our arm ----------------------------------------------------------------------------------------
FMT_NAMES = {"u8": "8-bit", "r30": "ddpack 10-bit packed (.r30, 3 samples / 4 bytes)", "lds": "10-bit packed (.lds, 4 samples / 5 bytes)",
             "u16": "10-bit unpacked uint16"}


def make_capture(system, fmt, ncap, seed):
    """(sample array for checks, bytes as the capture file holds them, library format id, samples)."""
    from lddecode_b200 import _lib, synth
    if fmt == "u8":
        cap = synth_capture(system, ncap, seed)
        return cap, cap, _lib.FMT_U8, ncap
    ncap = ncap // 12 * 12                                        # whole .r30 words and .lds groups
    s10 = synth_capture(system, ncap, seed, bits=10)
    if fmt == "u16":
        return s10, s10, _lib.FMT_U16, ncap
    if fmt == "r30":
        return s10, synth.pack_r30(s10).view(np.uint8), _lib.FMT_R30, ncap
    return s10, synth.pack_lds(s10), _lib.FMT_LDS40, ncap


def golden_field(system, fmt, seed):
    """The reference's own first field of this capture (tests/golden/make_bench_golden.py), or None."""
    key = {("PAL", "u8", 1): "PAL_u8_seed1", ("NTSC", "u8", 0): "NTSC_u8_seed0"}.get((system, fmt, seed))
    if key is None and system == "NTSC" and fmt in ("r30", "lds", "u16") and seed == 0:
        key = "NTSC_10bit_seed0"
    if key is None:
        return None
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "bench_fields.npz"))
        return g[key + "_pic"], int(g[key + "_next"]), int(g[key + "_istop"])
    except Exception:
        return None


def check_result(cd, res, system, fmt, seed, what):
    """Outside the timed region: every located field of the step must be clean (no error bits from any kernel), the
    field cadence must be the system's, and the first field must be the reference's own decode of these bytes (+-1 LSB)."""
    be = cd.rf._be
    be.synchronize()
    nloc = len(res.located)
    if nloc == 0:
        raise SystemExit("bench self-check failed (%s): no field located" % what)
    st = be.to_host(res.d_status)[:nloc]
    if np.any(st & 15):
        raise SystemExit("bench self-check failed (%s): field status bits %s" % (what, sorted(set(int(x) for x in st if x & 15))))
    infos = res.infos[np.asarray(res.located)]
    tops = infos['istop']
    if np.any(tops[1:] == tops[:-1]):
        raise SystemExit("bench self-check failed (%s): field parity does not alternate" % what)
    out = dict(fields=int(nloc), status_clean=True, golden="none")
    g = golden_field(system, fmt, seed)
    if g is not None:
        gpic, gnext, gtop = g
        j = res.located[0]
        W = cd.rf.SysParams['outlinelen']
        n = int(res.infos[j].linecount) * W
        pic = be.to_host(res.d_pic[:res.out_stride])[:n].astype(np.int64)
        if int(res.readsamples[j]) != 0 or int(res.infos[j].nextfieldoffset) != gnext or int(res.infos[j].istop) != gtop or n != len(gpic):
            raise SystemExit("bench self-check failed (%s): first field geometry differs from the reference's" % what)
        d = np.abs(pic - gpic.astype(np.int64))
        if d.max() > 1:
            raise SystemExit("bench self-check failed (%s): first field differs from the reference by %d LSB" % (what, int(d.max())))
        out["golden"] = "first field == reference's FieldPAL/FieldNTSC output within +-1 LSB (%.2f %% of samples differ by 1)" % (100.0 * float(np.mean(d > 0)))
    return out


def pin_rank_to_gpu_numa(local):
    """Run this rank on the CPUs next to its GPU (NVML's ideal CPU set intersected with what the cgroup allows)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = ideal & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
        return sorted(use)[:2] + ["..."] + sorted(use)[-1:] if use else None
    except Exception:
        return None


def measure(a, system, audio, fmt, rank, world, local, dist, clocks_rank0=None, roofline=True):
    """One workload through the resident and the end-to-end path.  Returns the result dict on rank 0."""
    import torch
    from lddecode_b200 import parallel, pipeline, rfdecode

    ncap = one_second(system) + TAIL
    seed = (1 + rank) if system == "PAL" else rank
    chk, raw, fmt_id, ncap = make_capture(system, fmt, ncap, seed)
    rf = rfdecode.RFDecode(FS[system], system, BLOCKLEN, decode_analog_audio=audio, device=local, precision=a.precision)
    cd = pipeline.CaptureDecoder(rf, max_fields=256)
    be = rf._be
    cap_dev = torch.from_numpy(raw).cuda()
    cap_pin = torch.from_numpy(raw).pin_memory()
    nraw = len(raw)
    max_fields = 64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # (LDD_BENCH_NO_GATHER: diagnostic runs that separate the collective's cost from the rest of a multi-GPU step)
    gatherer = parallel.make_gatherer(cd, rank, world, max_fields, dist) if world > 1 and not os.environ.get("LDD_BENCH_NO_GATHER") else None

    def run_resident(nsteps):
        # K decodes of the HBM-resident capture through CaptureDecoder.decode_stream: the demodulation of step k+1 is
        # enqueued before the host walks the fields of step k (same results as decode()); with several ranks every
        # step's fields are written into the gather's send buffer and collected on rank 0 over NCCL
        res = None
        for res in cd.decode_stream(((cap_dev, fmt_id, ncap) for _ in range(nsteps)), sink=gatherer):
            pass
        return res

    # end to end: the public host-buffer API (pipeline.HostStreamDecoder).  Every step uploads its capture -- the
    # bytes of the capture file, packed formats stay packed -- from pinned host memory and downloads its uint16 fields
    # (and audio) into pinned host memory; the upload of step k+1 and the download of step k-1 overlap the decode of
    # step k.  Each rank delivers its own fields to its host buffers (no gather on top: that would deliver them twice).
    sd = pipeline.HostStreamDecoder(cd, fmt_id, ncap, max_fields, np_dtype=raw.dtype, nbytes_max=nraw)

    def run_e2e(nsteps):
        pend = sd.launch(sd.upload(cap_pin, ncap, nraw))            # H2D of the first step's input
        t = sd.upload(cap_pin, ncap, nraw) if nsteps > 1 else None
        prev = None
        for i in range(nsteps):
            nxt = sd.launch(t) if t is not None else None           # demodulation of step i+1 ...
            t = sd.upload(cap_pin, ncap, nraw) if i + 2 < nsteps else None
            job = sd.finish(pend)                                   # ... under the host walk of step i; D2H of its fields
            pend = nxt
            if prev is not None:
                sd.fetch(prev)                                      # host reads the previous step's result
            prev = job
        res, pics = sd.fetch(prev)
        na = 0 if res.audio_host is None else 16 * len(res.audio_host[0])      # two float64 channels
        return res, pics, na

    # warm-up
    res = run_resident(max(a.warmup, 3))
    torch.cuda.synchronize()
    nfields = len(res.located)
    verify = check_result(cd, res, system, fmt, seed, "%s resident" % system)
    # samples of the capture demodulated and decoded per step (each counted once: neither the block
    # overlaps nor the halos that neighbouring ranges demodulate twice are counted)
    S_ = cd.stride
    consumed = ((ncap - BLOCKLEN) // S_ + 1) * S_

    # resident timing
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record()
    res = run_resident(a.steps)
    if gatherer is not None:
        gatherer.wait()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    check_result(cd, res, system, fmt, seed, "%s resident, timed run" % system)

    # end-to-end timing (host buffers)
    run_e2e(3)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    res, pics, naudio = run_e2e(a.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    tw1 = time.time()
    npic = pics.size
    if np.any(res.status_host & 15):
        raise SystemExit("bench self-check failed (%s e2e): field status bits" % system)
    g = golden_field(system, fmt, seed)
    if g is not None:
        d = np.abs(pics[0, :len(g[0])].astype(np.int64) - g[0].astype(np.int64))
        if d.max() > 1:
            raise SystemExit("bench self-check failed (%s e2e): downloaded field differs from the reference by %d LSB" % (system, int(d.max())))

    out = None
    k_ms = planes_total = None
    if roofline:
        # dominant kernel alone: the fused block demodulation (both passes of the mixed lane)
        planes_total = demod_only(cd, cap_dev, fmt_id, ncap)
        planes_total = demod_only(cd, cap_dev, fmt_id, ncap)
        torch.cuda.synchronize()
        kt = []
        for _ in range(max(a.steps, 5)):
            e0.record()
            planes_total = demod_only(cd, cap_dev, fmt_id, ncap)
            e1.record()
            torch.cuda.synchronize()
            kt.append(e0.elapsed_time(e1))
        k_ms = float(np.mean(kt))

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])
        c = torch.tensor([float(consumed)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        consumed_all = float(c[0])
    else:
        consumed_all = float(consumed)

    if rank == 0:
        ms_step = ms_total / a.steps
        value = consumed_all / (ms_step / 1e3) / 1e6
        e2e_val = consumed_all / (ms_e2e / a.steps / 1e3) / 1e6
        # kernels of this library launched per step: demodulation (+ the float64 re-run of the mixed lane), 5 of the peak
        # chase + the peak list's copy to pinned memory, audio phase 2, table upload, hsync refinement + fix-up, VBI
        # decode, pilot (per-line + per-field) or 2 x (burst lines + vote), TBC, + the gather's metadata upload
        # (the mixed lane at the default block length is ONE launch: float64 re-runs happen inside the fused kernel)
        launches_per_step = (2 if a.precision == "mixed" and BLOCKLEN != 16384 else 1) + 5 + 1 + (1 if audio else 0) + 1 + 2 + 1 + \
            (2 if system == "PAL" else 4) + 1 + (1 if world > 1 else 0)
        out = dict(value=value, ms_per_step=ms_step, realtime_x=value / FS[system] / world, fields_per_step=nfields * world,
                   gather=(None if gatherer is None else "NCCL gather" if type(gatherer).__name__ != "PeerGatherer" else
                           "one DMA transfer per step over NVLink into rank 0's mapped buffer (no collective kernel)" if gatherer.push
                           else "TBC kernels store over NVLink into rank 0's mapped buffer (no collective kernel)"),
                   e2e=dict(value=e2e_val, unit="Msamples/s", h2d_bytes_per_step=int(nraw * raw.dtype.itemsize),
                            d2h_bytes_per_step=int(npic * 2 + naudio), wall_ms_per_step=wall_e2e / a.steps),
                   gpu_launches=launches_per_step * a.steps, self_check=verify, window=(tw0, tw1))
        if roofline:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            peak = float(peaks.get("hbm_gbs", 6650.0))
            nplanes32 = 4 if system == "PAL" else 3
            N, S = BLOCKLEN, BLOCKLEN - 1056
            b_in = {"u8": 1.0, "u16": 2.0, "r30": 4 / 3, "lds": 1.25}[fmt]
            # algorithmic bytes per RF sample of this kernel: overlapped input, float32 planes, the float64
            # sync plane, phase-1 audio (two float64 channels at fs/8 NTSC, fs/16 PAL)
            bytes_per_sample = b_in * N / S + 4 * nplanes32 + 8 + (2 * 8 / (16 if system == "PAL" else 8) if audio else 0)
            achieved = bytes_per_sample * planes_total / (k_ms / 1e3) / 1e9
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "demod_traffic.json")))
                key = "%s_%s_%s%s" % (system, fmt, a.precision, "_audio" if audio else "")
                if key in tj:
                    traffic = tj[key]["dram_bytes_per_launch"]
            except Exception:
                pass
            out["roofline"] = dict(bound="hbm", kernel="demod_kernel (fused unpack+FFT+filter+IFFT+FM discriminator+post filters+sync scan)",
                                   achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                                   bytes_per_sample=bytes_per_sample, kernel_ms=k_ms, kernel_msamples_per_s=planes_total / k_ms / 1e3,
                                   peak_source="MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s")
    del sd, cd, rf, cap_dev, cap_pin, gatherer
    torch.cuda.empty_cache()
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (a.gpus, a.gpus))
    torch.cuda.set_device(local)
    numa = pin_rank_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.scaling == "strong":
        out = run_strong(a, rank, world, local, dist)
        if rank == 0:
            out.update(warmup=2, vs_baseline=None, dtype={"mixed": "f32+f64", "f64": "f64", "f32": "f32"}[a.precision], data="synthetic")
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    system, audio, fmt = a.system, a.audio, a.fmt
    # the sampler is started before the warm-up (NVML initialisation, or nvidia-smi's start-up in the fallback, stalls
    # the driver for ~100 ms)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.5)
    m = measure(a, system, audio, fmt, rank, world, local, dist)
    extra = []
    if a.extra and world == 1:
        # BASELINE.json configs[2]: NTSC, both analog audio channels, 10-bit packed input -- the heaviest configuration
        # of the path (burst refinement x2, two audio stages), measured the same way in the same run
        for xs, xa, xf in (("NTSC", True, "lds"),):
            if (xs, xa, xf) == (system, audio, fmt):
                continue
            x = measure(a, xs, xa, xf, rank, world, local, dist)
            x.pop("window", None)
            extra.append(dict(config=workload_config(xs, xa, world, xf), metric="rf_msamples_per_s_demod_tbc", unit="Msamples/s", **x))
    strong = None
    if world > 1 and a.extra:
        # BASELINE.json configs[3] in the same run: ONE NTSC CLV capture (1 s per GPU, so every rank still has a
        # second of work) sharded by block range with halos -- the north-star split, next to the weak-scaling headline
        strong = run_strong(a, rank, world, local, dist, seconds=float(world), steps=max(3, a.steps // 4))
    if rank == 0:
        tw0, tw1 = m.pop("window")
        clk = clocks.stop(tw0, tw1)
        line = dict(metric="rf_msamples_per_s_demod_tbc", value=m.pop("value"), unit="Msamples/s", n_gpus=world, steps=a.steps,
                    warmup=max(a.warmup, 3), ms_per_step=m.pop("ms_per_step"), higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"mixed": "f32+f64", "f64": "f64", "f32": "f32"}[a.precision], data="synthetic",
                    config=workload_config(system, audio, world, fmt), precision=a.precision, **m)
        line["clocks"] = clk
        if numa:
            line["cpu_affinity"] = numa
        if strong is not None:
            extra.append(strong)
        if extra:
            line["other_workloads"] = extra
        if world == 1 and not a.skip_cpu:
            line["cpu_baseline"] = cpu_baseline(system, audio)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_strong(a, rank, world, local, dist, seconds=None, steps=None):
    """BASELINE.json configs[3]: ONE NTSC CLV capture of `seconds` sharded by contiguous read-position ranges (block
    ranges + halos) over the ranks; only the per-field outputs cross NVLink (NCCL gather to rank 0).  The capture is the
    tiled synthetic one (lddecode_b200.synth.TiledCapture: a pure function of the sample index), every rank generates
    the window its shard needs in its own HBM (untimed) and decodes it chunk by chunk (~1 s of capture per chunk, two
    plane workspaces, chunk k+1 demodulated under the host walk of chunk k).  A step = the whole capture once."""
    import torch
    from lddecode_b200 import _lib, parallel, pipeline, rfdecode, synth
    seconds = a.seconds if seconds is None else seconds
    steps = a.steps if steps is None else steps
    fs = FS["NTSC"]
    T = int(round(seconds * fs * 1e6))
    ncap = T + TAIL
    rf = rfdecode.RFDecode(fs, "NTSC", BLOCKLEN, decode_analog_audio=a.audio, device=local, precision=a.precision)
    cd = pipeline.CaptureDecoder(rf, max_fields=256)
    R0, R1 = parallel.shard_bounds(ncap, world)[rank]
    lo, hi = parallel.needed_window(cd, ncap, R0, R1)
    tc = synth.TiledCapture(seed=2, device="cuda")
    cap_dev = tc.generate(lo, hi - lo)
    torch.cuda.synchronize()
    # the same number of chunks on every rank (one gather per chunk): ~1 s of capture each, at most 1.25 s
    shard = ncap // world + 1
    nch = max(1, -(-shard // int(1.25 * one_second("NTSC"))))
    chunk = -(-shard // nch)
    edges = [R0 + i * chunk for i in range(nch)] + [R1]
    ranges = [(cap_dev, _lib.FMT_U8, lo, hi - lo, ncap, edges[i], edges[i + 1]) for i in range(nch)]
    max_fields = 80
    gatherer = parallel.make_gatherer(cd, rank, world, max_fields, dist) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_once():
        nf, last = 0, None
        for res in cd.decode_stream(iter(ranges), sink=gatherer):
            nf += len(res.located)
            last = res
        return nf, last

    nf, last = run_once()
    run_once()
    torch.cuda.synchronize()
    st = rf._be.to_host(last.d_status)[:len(last.located)] if last.located else np.zeros(0, dtype=np.int32)
    if np.any(st & 15):
        raise SystemExit("bench self-check failed (strong): field status bits")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run_once()
    if world > 1:
        gatherer.wait()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    nft = torch.tensor([float(nf)], device="cuda", dtype=torch.float64)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        dist.all_reduce(nft, op=dist.ReduceOp.SUM)
    out = None
    if rank == 0:
        ms_step = ms / steps
        # every sample of the capture counted once (halos that neighbouring shards / chunks demodulate twice are not)
        value = T / (ms_step / 1e3) / 1e6
        out = dict(metric="rf_msamples_per_s_demod_tbc", value=value, unit="Msamples/s", n_gpus=world, steps=steps,
                   ms_per_step=ms_step, higher_is_better=True, scaling="strong", realtime_x=value / fs,
                   fields_per_step=int(nft[0]), capture_seconds=seconds,
                   config=dict(workload="NTSC CLV synthetic 8-bit RF, ONE capture of %.1f s (%.2f Gsamples, tiled 2-frame template, "
                                        "running carrier phase), sharded by block range with halos over %d GPU(s), %s demod + sync + TBC"
                                        % (seconds, T / 1e9, world, "video + both analog audio channels" if a.audio else "video"),
                               blocklen=BLOCKLEN, readlen=1000000, chunk_samples=chunk, chunks_per_rank=len(ranges),
                               parallelism="read-position ranges [g T/G, (g+1) T/G) per rank, walk starts 1.6 fields early, "
                                           "demodulates one read length past the end; NCCL gather of uint16 fields to rank 0"))
    del cap_dev, gatherer, cd, rf
    torch.cuda.empty_cache()
    return out


def demod_only(cd, cap_dev, fmt_id, ncap):
    """One launch of the demodulation kernel over the whole capture (what the roofline entry times)."""
    rf, be = cd.rf, cd.rf._be
    rf._set_mtf(cd.mtf_level)
    S, N = cd.stride, rf.blocklen
    first_block, nblocks, _ = cd.plan_range(ncap, 0, ncap + 1)
    while nblocks > 0 and first_block + (nblocks - 1) * S + N > ncap:
        nblocks -= 1
    total = nblocks * S
    if not hasattr(cd, "_bench_planes") or cd._bench_planes[1] != total:
        cd._bench_planes = (rf._alloc_planes(total), total)
    (planes, parr), _ = cd._bench_planes
    a1 = None
    alen = 0
    if rf.decode_analog_audio:
        ds = N // len(rf.Filters['audio_lfilt'])
        alen = total // ds
        if not hasattr(cd, "_bench_audio"):
            cd._bench_audio = (be.empty(alen, np.float64), be.empty(alen, np.float64))
        a1 = cd._bench_audio
    rf._check(be.lib.ldd_demod_blocks(rf._h, be.ptr(cap_dev), fmt_id, 0, int(ncap), 0, int(nblocks), int(total), parr,
                                      be.ptr(a1[0]) if a1 else None, be.ptr(a1[1]) if a1 else None, int(alen), be.stream()))
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--system", default="PAL", choices=["PAL", "NTSC"])
    ap.add_argument("--audio", action="store_true", help="also demodulate the two analog FM audio channels")
    ap.add_argument("--precision", default="mixed", choices=["f64", "f32", "mixed"],
                    help="demodulation lane (DESIGN.md 3.1); 'mixed' is the library default")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg (profiling runs)")
    ap.add_argument("--fmt", default="u8", choices=["u8", "u16", "r30", "lds"], help="capture sample format fed to the decoder")
    ap.add_argument("--no-extra", dest="extra", action="store_false",
                    help="N=1: skip the second workload (NTSC + both audio channels from packed .lds, BASELINE configs[2])")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: one 1-s shard per GPU; strong: ONE capture of --seconds sharded by block range (configs[3])")
    ap.add_argument("--seconds", type=float, default=8.0, help="--scaling strong: length of the capture")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

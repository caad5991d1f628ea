// Kernel (4): sync-pulse peak list.  Bit-exact restatement of Field.get_syncpeaks
// (lddecode_core.py:497-516):
//     i = start
//     while i < len(ds) - 2*linelen:
//         p = argmax(ds[i : i + linelen//2]);  if ds[i+p] > .2: append(i+p); i += p + int(.4*linelen)
//         else: i += linelen//2
// The chase is sequential, but its state after appending a peak is a function of that peak only,
// so chains started at different places coincide from their first common peak on.  Phase 1 runs
// one warp per segment of the plane (window arg-max by warp shuffles, first index wins ties like
// numpy); phase 2 walks the segments in order with one warp, splices each chain into the true
// one at the first common peak and, if a chain has not merged inside the overlap, simply keeps
// chasing itself -- so the result is the sequential algorithm's list in every case.
#include "ldd_internal.h"

namespace ldd {

struct ArgMax {
    double v;
    int idx;
};

// arg-max of ds[i .. i+len), first index among equal values; all 32 lanes get the result.
// The chase is a chain of dependent windows, so what matters is the number of memory round trips per
// window: with len <= 32*PER all of a lane's loads are issued before the first comparison (one round
// trip per window instead of one per four loads).
template <int PER>
__device__ inline ArgMax warp_argmax_batched(const double* __restrict__ ds, long long i, int len, int lane) {
    double v[PER];
    LDD_UNROLL
    for (int q = 0; q < PER; ++q) {
        int k = lane + 32 * q;
        v[q] = k < len ? ds[i + k] : -1e300;
    }
    ArgMax b;
    b.v = -1e300;
    b.idx = 0x7fffffff;
    LDD_UNROLL
    for (int q = 0; q < PER; ++q)
        if (v[q] > b.v) { b.v = v[q]; b.idx = lane + 32 * q; }
    for (int d = 16; d > 0; d >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, b.v, d);
        int oi = __shfl_xor_sync(0xffffffffu, b.idx, d);
        if (ov > b.v || (ov == b.v && oi < b.idx)) { b.v = ov; b.idx = oi; }
    }
    return b;
}

__device__ inline ArgMax warp_argmax(const double* __restrict__ ds, long long i, int len, int lane) {
    if (len <= 32 * 30) return warp_argmax_batched<30>(ds, i, len, lane);      // NTSC at 8fsc: 910
    if (len <= 32 * 40) return warp_argmax_batched<40>(ds, i, len, lane);      // PAL at 8fsc: 1135, 40 MSPS: 1280
    ArgMax b;
    b.v = -1e300;
    b.idx = 0x7fffffff;
    int k = lane;
    for (; k + 96 < len; k += 128) {            // four independent loads in flight per lane
        double v0 = ds[i + k], v1 = ds[i + k + 32], v2 = ds[i + k + 64], v3 = ds[i + k + 96];
        if (v0 > b.v) { b.v = v0; b.idx = k; }
        if (v1 > b.v) { b.v = v1; b.idx = k + 32; }
        if (v2 > b.v) { b.v = v2; b.idx = k + 64; }
        if (v3 > b.v) { b.v = v3; b.idx = k + 96; }
    }
    for (; k < len; k += 32) {
        double v = ds[i + k];
        if (v > b.v) { b.v = v; b.idx = k; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, b.v, d);
        int oi = __shfl_xor_sync(0xffffffffu, b.idx, d);
        if (ov > b.v || (ov == b.v && oi < b.idx)) { b.v = ov; b.idx = oi; }
    }
    return b;
}

struct PeakGeom {
    long long n, start, limit;     // plane length, first i, loop bound (n - 2*linelen)
    int half, skip;                // linelen//2, int(.4*linelen)
    long long seg, ov;             // segment length and overlap (samples)
    int nseg, cap_seg;             // segments, capacity of each segment's list
};

// Phase 1: segment s chases from start + s*seg until i passes the next segment's start + ov.
__global__ void __launch_bounds__(128) peaks_phase1(const double* __restrict__ ds, PeakGeom g, long long* __restrict__ pos,
                                                    double* __restrict__ val, int* __restrict__ cnt,
                                                    long long* __restrict__ iend) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= g.nseg) return;
    long long i = g.start + (long long)s * g.seg;
    long long stop = i + g.seg + g.ov;
    if (s == g.nseg - 1 || stop > g.limit) stop = g.limit;
    int c = 0;
    long long* mypos = pos + (size_t)s * g.cap_seg;
    double* myval = val + (size_t)s * g.cap_seg;
    while (i < stop) {
        ArgMax m = warp_argmax(ds, i, g.half, lane);
        if (m.v > .2) {
            if (lane == 0 && c < g.cap_seg) { mypos[c] = i + m.idx; myval[c] = m.v; }
            ++c;
            i += m.idx + g.skip;
        } else {
            i += g.half;
        }
    }
    if (lane == 0) { cnt[s] = c < g.cap_seg ? c : g.cap_seg; iend[s] = i; }
}

// Phase 2a: one thread per segment boundary finds where the chain of segment s and the chain of
// segment s+1 first share a peak (they are identical from there on).  hi[s] = number of peaks of
// segment s that belong to the true chain, lo[s+1] = index of the shared peak in segment s+1's list;
// merged[s] = 0 when the two chains did not meet inside the overlap.
__global__ void __launch_bounds__(128) peaks_merge(PeakGeom g, const long long* __restrict__ pos, const int* __restrict__ cnt,
                                                   int* __restrict__ hi, int* __restrict__ lo, int* __restrict__ merged,
                                                   int* __restrict__ any_unmerged) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) lo[0] = 0;
    if (s >= g.nseg) return;
    const long long* p = pos + (size_t)s * g.cap_seg;
    const int c = cnt[s];
    if (s == g.nseg - 1) { hi[s] = c; merged[s] = 1; return; }
    const long long* q = pos + (size_t)(s + 1) * g.cap_seg;
    const int cq = cnt[s + 1];
    const long long xt = g.start + (long long)(s + 1) * g.seg;
    // first peak of this chain at or beyond the next segment's start (binary search)
    int a0 = 0, a1 = c;
    while (a0 < a1) { int m = (a0 + a1) >> 1; if (p[m] < xt) a0 = m + 1; else a1 = m; }
    int a = a0, b = 0, ok = 0;
    while (a < c && b < cq) {
        long long pa = p[a], qb = q[b];
        if (pa == qb) { ok = 1; break; }
        if (pa < qb) ++a; else ++b;
    }
    hi[s] = ok ? a : c;
    lo[s + 1] = ok ? b : 0;
    merged[s] = ok;
    if (!ok) atomicOr(any_unmerged, 1);
}

// Phase 2b (all chains merged -- the normal case): exclusive prefix sum of the per-segment counts by one
// CTA (1024 segments per sweep), then a grid-wide copy with one warp per segment.
__global__ void __launch_bounds__(1024) peaks_scan(PeakGeom g, const int* __restrict__ hi, const int* __restrict__ lo,
                                                   const int* __restrict__ any_unmerged, int* __restrict__ offs,
                                                   int* __restrict__ out_count) {
    if (*any_unmerged) return;                  // the sequential kernel below produces the list instead
    __shared__ int s_warp[32];
    __shared__ int s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < g.nseg; base += 1024) {
        const int s = base + tid;
        int n = 0;
        if (s < g.nseg) { n = hi[s] - lo[s]; n = n > 0 ? n : 0; }
        int incl = n;
        for (int d = 1; d < 32; d <<= 1) {
            int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                int up = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += up;
            }
            s_warp[lane] = wi - w;              // exclusive over warps
        }
        __syncthreads();
        const int run = s_run;
        if (s < g.nseg) offs[s] = run + s_warp[warp] + incl - n;
        __syncthreads();
        if (tid == 1023) s_run = run + s_warp[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) { out_count[0] = s_run; out_count[1] = 0; }
}

__global__ void __launch_bounds__(256) peaks_copy(PeakGeom g, const long long* __restrict__ pos, const double* __restrict__ val,
                                                  const int* __restrict__ hi, const int* __restrict__ lo,
                                                  const int* __restrict__ any_unmerged, const int* __restrict__ offs,
                                                  long long* __restrict__ out_pos, double* __restrict__ out_val, int cap) {
    if (*any_unmerged) return;
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= g.nseg) return;
    const long long* p = pos + (size_t)s * g.cap_seg;
    const double* v = val + (size_t)s * g.cap_seg;
    const int l = lo[s], h = hi[s], o = offs[s];
    for (int k = l + lane; k < h; k += 32) {
        int d = o + (k - l);
        if (d < cap) { out_pos[d] = p[k]; out_val[d] = v[k]; }
    }
}

// Phase 2 (general): one warp, warp-uniform control flow; runs only when some chain did not merge
// inside its overlap.  out_count[0] = number of peaks (may exceed cap: then only the first cap are
// stored), out_count[1] = number of chase steps taken here.
__global__ void __launch_bounds__(32) peaks_phase2(const double* __restrict__ ds, PeakGeom g, const long long* __restrict__ pos,
                                                   const double* __restrict__ val, const int* __restrict__ cnt,
                                                   const long long* __restrict__ iend, const int* __restrict__ any_unmerged,
                                                   long long* __restrict__ out_pos,
                                                   double* __restrict__ out_val, int cap, int* __restrict__ out_count) {
    if (!*any_unmerged) return;
    const int lane = threadIdx.x;
    int n_out = 0, extra = 0;
    int s = 0, lo = 0;
    while (s < g.nseg) {
        const long long* p = pos + (size_t)s * g.cap_seg;
        const double* v = val + (size_t)s * g.cap_seg;
        const int c = cnt[s];
        if (s == g.nseg - 1) {
            for (int k = lo + lane; k < c; k += 32)
                if (n_out + (k - lo) < cap) { out_pos[n_out + (k - lo)] = p[k]; out_val[n_out + (k - lo)] = v[k]; }
            n_out += (c > lo ? c - lo : 0);
            break;
        }
        // try to merge with the next segment inside the overlap
        int t = s + 1;
        long long xt = g.start + (long long)t * g.seg;
        int a = lo;
        while (a < c && p[a] < xt) ++a;
        int b = 0;
        bool merged = false;
        {
            const long long* q = pos + (size_t)t * g.cap_seg;
            int cq = cnt[t];
            while (a < c && b < cq) {
                if (p[a] == q[b]) { merged = true; break; }
                if (p[a] < q[b]) ++a; else ++b;
            }
        }
        int hi = merged ? a : c;
        for (int k = lo + lane; k < hi; k += 32)
            if (n_out + (k - lo) < cap) { out_pos[n_out + (k - lo)] = p[k]; out_val[n_out + (k - lo)] = v[k]; }
        n_out += (hi > lo ? hi - lo : 0);
        if (merged) { s = t; lo = b; continue; }
        // not merged: keep chasing from where this chain stopped until it lands on a peak that a
        // later segment's chain also found
        long long i = iend[s];
        b = 0;
        bool done = false;
        while (!done) {
            if (i >= g.limit) { s = g.nseg; break; }            // reached the end of the plane
            ArgMax m = warp_argmax(ds, i, g.half, lane);
            ++extra;
            if (m.v > .2) {
                long long pk = i + m.idx;
                // advance to the segment whose chain could contain pk
                while (t < g.nseg - 1 && pk >= g.start + (long long)(t + 1) * g.seg) { ++t; b = 0; }
                const long long* q = pos + (size_t)t * g.cap_seg;
                int cq = cnt[t];
                while (b < cq && q[b] < pk) ++b;
                if (b < cq && q[b] == pk) { s = t; lo = b; done = true; break; }
                if (lane == 0 && n_out < cap) { out_pos[n_out] = pk; out_val[n_out] = m.v; }
                ++n_out;
                i += m.idx + g.skip;
            } else {
                i += g.half;
            }
        }
    }
    if (lane == 0) { out_count[0] = n_out; out_count[1] = extra; }
}

}  // namespace ldd

using namespace ldd;

extern "C" int ldd_sync_peaks(ldd_handle* h, const double* sync_dev, long long n, long long start,
                              long long* peaks_dev, double* vals_dev, int cap, int* count_dev, void* stream) {
    if (!h || !sync_dev || !peaks_dev || !vals_dev || !count_dev || n < 0 || start < 0 || cap < 0) return LDD_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = h->cfg.linelen;
    PeakGeom g;
    g.n = n;
    g.start = start;
    g.limit = n - 2LL * L;
    g.half = L / 2;
    g.skip = (int)(L * .4);
    if (g.limit <= start) {
        cudaMemsetAsync(count_dev, 0, 2 * sizeof(int), st);
        return LDD_OK;
    }
    const char* env = getenv("LDD_PEAK_SEG_LINES");
    long long seg_lines = env ? atoll(env) : 12;
    if (seg_lines < 4) seg_lines = 4;
    g.seg = seg_lines * L;
    g.ov = 6LL * L;
    if (g.ov > g.seg) g.ov = g.seg;      // a chain must merge before the next boundary for the boundaries to be independent
    long long span = g.limit - start;
    g.nseg = (int)((span + g.seg - 1) / g.seg);
    if (g.nseg < 1) g.nseg = 1;
    g.cap_seg = (int)((g.seg + g.ov) / (g.skip > 0 ? g.skip : 1)) + 4;
    size_t need = (size_t)g.nseg * g.cap_seg * (sizeof(long long) + sizeof(double)) + (size_t)g.nseg * (5 * sizeof(int) + sizeof(long long)) + 128;
    if (need > h->peak_ws_bytes) {
        if (h->peak_ws) { cudaStreamSynchronize(st); cudaFree(h->peak_ws); h->peak_ws = nullptr; }
        size_t grow = need + need / 2;
        if (cudaMalloc(&h->peak_ws, grow) != cudaSuccess) { h->err = "cudaMalloc peak workspace"; h->peak_ws_bytes = 0; return LDD_ENOMEM; }
        h->peak_ws_bytes = grow;
    }
    char* w = (char*)h->peak_ws;
    long long* pos = (long long*)w;                 w += (size_t)g.nseg * g.cap_seg * sizeof(long long);
    double* val = (double*)w;                       w += (size_t)g.nseg * g.cap_seg * sizeof(double);
    long long* iend = (long long*)w;                w += (size_t)g.nseg * sizeof(long long);
    int* cnt = (int*)w;                             w += (size_t)g.nseg * sizeof(int);
    int* hi = (int*)w;                              w += (size_t)g.nseg * sizeof(int);
    int* lo = (int*)w;                              w += (size_t)g.nseg * sizeof(int);
    int* merged = (int*)w;                          w += (size_t)g.nseg * sizeof(int);
    int* offs = (int*)w;                            w += (size_t)g.nseg * sizeof(int);
    int* any_unmerged = (int*)w;
    const int warps = 4;
    cudaMemsetAsync(any_unmerged, 0, sizeof(int), st);
    LDD_LAUNCH(peaks_phase1, dim3((g.nseg + warps - 1) / warps), dim3(32 * warps), 0, st, sync_dev, g, pos, val, cnt, iend);
    LDD_LAUNCH(peaks_merge, dim3((g.nseg + 127) / 128), dim3(128), 0, st, g, (const long long*)pos, (const int*)cnt, hi, lo, merged,
               any_unmerged);
    LDD_LAUNCH(peaks_scan, dim3(1), dim3(1024), 0, st, g, (const int*)hi, (const int*)lo, (const int*)any_unmerged, offs, count_dev);
    LDD_LAUNCH(peaks_copy, dim3((g.nseg + 7) / 8), dim3(256), 0, st, g, (const long long*)pos, (const double*)val, (const int*)hi,
               (const int*)lo, (const int*)any_unmerged, (const int*)offs, peaks_dev, vals_dev, cap);
    LDD_LAUNCH(peaks_phase2, dim3(1), dim3(32), 0, st, sync_dev, g, (const long long*)pos, (const double*)val,
               (const int*)cnt, (const long long*)iend, (const int*)any_unmerged, peaks_dev, vals_dev, cap, count_dev);
    return launch_status(h, "peaks_phase1/2");
}

// ---- small transfers by kernel (see include/ldd_b200.h) ------------------------------------------
namespace ldd {
__global__ void __launch_bounds__(256) copy_small_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t n) {
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16, step = (size_t)gridDim.x * blockDim.x * 16;
    if (((((uintptr_t)dst) | ((uintptr_t)src)) & 15) == 0) {
        for (size_t i = i0; i + 16 <= n; i += step) *(uint4*)(dst + i) = *(const uint4*)(src + i);
        const size_t tail = n & ~(size_t)15;
        if (blockIdx.x == 0 && threadIdx.x < (n & 15)) dst[tail + threadIdx.x] = src[tail + threadIdx.x];
    } else {
        for (size_t i = i0; i < n; i += step)
            for (size_t k = i; k < i + 16 && k < n; ++k) dst[k] = src[k];
    }
}

__global__ void __launch_bounds__(256) peaks_to_host_kernel(const long long* __restrict__ pk, const double* __restrict__ vl,
                                                            const int* __restrict__ cnt, int cap, long long* __restrict__ hpk,
                                                            double* __restrict__ hvl, int* __restrict__ hcnt) {
    const int c = cnt[0] < cap ? cnt[0] : cap;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { hcnt[0] = cnt[0]; hcnt[1] = cnt[1]; }
    for (int i = t; i < c; i += gridDim.x * blockDim.x) { hpk[i] = pk[i]; hvl[i] = vl[i]; }
}
}  // namespace ldd

extern "C" int ldd_copy_small(void* dst, const void* src, size_t nbytes, void* stream) {
    if (!dst || !src) return LDD_EINVAL;
    if (nbytes == 0) return LDD_OK;
    size_t blocks = (nbytes / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    LDD_LAUNCH(copy_small_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (unsigned char*)dst, (const unsigned char*)src, nbytes);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

extern "C" int ldd_peaks_to_host(const long long* peaks_dev, const double* vals_dev, const int* count_dev, int cap,
                                 long long* peaks_host, double* vals_host, int* count_host, void* stream) {
    if (!peaks_dev || !vals_dev || !count_dev || !peaks_host || !vals_host || !count_host || cap < 0) return LDD_EINVAL;
    LDD_LAUNCH(peaks_to_host_kernel, dim3(32), dim3(256), 0, (cudaStream_t)stream, peaks_dev, vals_dev, count_dev, cap, peaks_host,
               vals_host, count_host);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

// HOST restatement of the same chase for short windows whose samples are already in host memory (the
// field walk's off-peak window starts, pipeline.py): comparisons only, so it returns exactly the list
// the kernels above return for the same samples.
extern "C" int ldd_sync_peaks_host(ldd_handle* h, const double* sync_host, long long n, long long start,
                                   long long* peaks, double* vals, int cap, int* count) {
    if (!h || !sync_host || !peaks || !vals || !count || n < 0 || start < 0 || cap < 0) return LDD_EINVAL;
    const int L = h->cfg.linelen;
    const long long limit = n - 2LL * L;
    const int half = L / 2, skip = (int)(L * .4);
    long long i = start;
    int c = 0;
    while (i < limit) {
        const double* w = sync_host + i;
        double best = -1e300;
        int at = 0;
        for (int k = 0; k < half; ++k)
            if (w[k] > best) { best = w[k]; at = k; }
        if (best > .2) {
            if (c < cap) { peaks[c] = i + at; vals[c] = best; }
            ++c;
            i += at + skip;
        } else {
            i += half;
        }
    }
    *count = c;
    return LDD_OK;
}


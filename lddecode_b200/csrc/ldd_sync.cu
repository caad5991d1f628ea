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

// arg-max of ds[i .. i+len), first index among equal values; all 32 lanes get the result
__device__ inline ArgMax warp_argmax(const double* __restrict__ ds, long long i, int len, int lane) {
    ArgMax b;
    b.v = -1e300;
    b.idx = 0x7fffffff;
    for (int k = lane; k < len; k += 32) {
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

// Phase 2: one warp, warp-uniform control flow.  out_count[0] = number of peaks (may exceed cap:
// then only the first cap are stored), out_count[1] = number of chase steps taken here (0 when
// every chain merged inside its overlap).
__global__ void __launch_bounds__(32) peaks_phase2(const double* __restrict__ ds, PeakGeom g, const long long* __restrict__ pos,
                                                   const double* __restrict__ val, const int* __restrict__ cnt,
                                                   const long long* __restrict__ iend, long long* __restrict__ out_pos,
                                                   double* __restrict__ out_val, int cap, int* __restrict__ out_count) {
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
        int a_first = a, b = 0;
        bool merged = false;
        {
            const long long* q = pos + (size_t)t * g.cap_seg;
            int cq = cnt[t];
            while (a < c && b < cq) {
                if (p[a] == q[b]) { merged = true; break; }
                if (p[a] < q[b]) ++a; else ++b;
            }
        }
        (void)a_first;
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
    long long seg_lines = env ? atoll(env) : 48;
    if (seg_lines < 4) seg_lines = 4;
    g.seg = seg_lines * L;
    g.ov = 6LL * L;
    long long span = g.limit - start;
    g.nseg = (int)((span + g.seg - 1) / g.seg);
    if (g.nseg < 1) g.nseg = 1;
    g.cap_seg = (int)((g.seg + g.ov) / (g.skip > 0 ? g.skip : 1)) + 4;
    size_t need = (size_t)g.nseg * g.cap_seg * (sizeof(long long) + sizeof(double)) + (size_t)g.nseg * (sizeof(int) + sizeof(long long)) + 64;
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
    int* cnt = (int*)w;
    const int warps = 4;
    LDD_LAUNCH(peaks_phase1, dim3((g.nseg + warps - 1) / warps), dim3(32 * warps), 0, st, sync_dev, g, pos, val, cnt, iend);
    LDD_LAUNCH(peaks_phase2, dim3(1), dim3(32), 0, st, sync_dev, g, (const long long*)pos, (const double*)val,
               (const int*)cnt, (const long long*)iend, peaks_dev, vals_dev, cap, count_dev);
    return launch_status(h, "peaks_phase1/2");
}

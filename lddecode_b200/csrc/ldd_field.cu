// Field location: what Field.__init__ / FieldNTSC / FieldPAL do between the peak list and the
// final resample (lddecode_core.py:518-787, 889-957, 962-1021, 1054-1133).
//
//   ldd_field_locate      HOST.  vsync detection, field parity, integer line numbering, gap
//                         interpolation (get_hsync_median .. compute_linelocs, Field.__init__
//                         early-outs).  Scalar decisions over <= ~600 peaks per field: SURVEY.md
//                         section 8 (A7, A12) keeps this on the host; it reads only peak indices and
//                         their demod_sync values.
//   ldd_refine_hsync      DEVICE, one CTA per field: refine_linelocs_hsync (A8).
//   ldd_refine_burst      DEVICE, one CTA per field: FieldNTSC.refine_linelocs_burst (A9), including
//                         the spline resample of the burst plane at the 40 samples it looks at.
//   ldd_refine_pilot      DEVICE, one CTA per field: FieldPAL.refine_linelocs_pilot (A10).
//
// numpy details that decide thresholds are reproduced: pairwise summation order of np.mean/np.std,
// round-half-even of np.round, float32 burst levels.
#include "ldd_internal.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <vector>

namespace ldd {

// numpy's pairwise summation (umath loops, PW_BLOCKSIZE = 128), which np.mean / np.std use.
// Device code only ever sums <= 128 values, so the device-callable part has no recursion (a
// recursive device function hides its stack need from the launch and overflows the frame).
LDD_HD inline double np_sum_block(const double* a, int n) {       // n <= 128
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

inline double np_sum_host(const double* a, int n) {              // any n (host)
    if (n <= 128) return np_sum_block(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum_host(a, n2) + np_sum_host(a + n2, n - n2);
}

LDD_HD inline double np_sum(const double* a, int n) { return np_sum_block(a, n); }

LDD_HD inline double np_mean(const double* a, int n) { return np_sum(a, n) / (double)n; }

// lddutils.calczc on a window accessor: data(k) for k in [0, len).  Returns false for None.
template <class F>
LDD_HD inline bool calczc(F data, long long len, long long start, double target, int count, double* out) {
    if (start < 0 || start >= len) return false;
    long long end = start + (long long)count + 1;
    if (end > len) end = len;
    bool rising = data(start) < target;
    long long x = -1;
    for (long long k = start; k < end; ++k) {
        double v = data(k);
        if (rising ? (v >= target) : (v <= target)) { x = k; break; }
    }
    if (x < 0 || x == 0) return false;
    double a = data(x - 1) - target, b = data(x) - target;
    double y = -a / (-a + b);
    *out = (double)(x - 1) + y;
    return true;
}

// ---------------------------------------------------------------------------------------------
// A8: refine_linelocs_hsync (lddecode_core.py:715-787).  Threads own lines; thread 0 finishes the
// sequential fix-ups.
struct HsyncParams {
    const float* d05;          // demod_05 plane (relative to ire0)
    long long n;               // plane length
    double ire0, hz_ire, freq; // freq in MHz
    int linelen;
    const long long* base;     // [nfields] plane index of the field window's sample 0
    const long long* winlen;   // [nfields] window length (len(ds) of the reference)
    const int* linecount;      // [nfields]
    int ll_stride;
    const double* linelocs1;   // [nfields][ll_stride]
    const unsigned char* linebad_in;
    double* linelocs2;
    unsigned char* linebad_out;
    int* status;               // bit 1: the reference would have raised (field invalid)
    int nfields;
};

constexpr int HS_WARPS = 8;
constexpr int HS_WIN = 192;        // samples in 4 us (zc - 1 us .. zc + 3 us) at <= 40 MSPS, + slack

// One warp per line: the scans of the reference (first crossing of -20 IRE within 400 samples, min / max
// over two windows) run 32 samples at a time with coalesced loads; the window around the edge is staged
// in shared memory for the scalar tail (means, half-level crossing) that lane 0 finishes.
__global__ void __launch_bounds__(32 * HS_WARPS) refine_hsync_kernel(const HsyncParams p) {
    __shared__ double s_win[HS_WARPS][HS_WIN];
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nll = p.linecount[f] + 4;
    const int i = blockIdx.x * HS_WARPS + warp;
    if (i >= nll) return;
    const long long base = p.base[f], wlen = p.winlen[f];
    const double* l1 = p.linelocs1 + (size_t)f * p.ll_stride;
    const unsigned char* bad_in = p.linebad_in + (size_t)f * p.ll_stride;
    double* l2 = p.linelocs2 + (size_t)f * p.ll_stride;
    unsigned char* bad = p.linebad_out + (size_t)f * p.ll_stride;
    const double fq = p.freq;
    auto hz = [&](double ire) { return p.ire0 + p.hz_ire * ire; };
    auto d = [&](long long k) -> double { return (double)p.d05[base + k] + p.ire0; };
    // window length visible to the reference = min(wlen, plane end)
    long long len = wlen;
    if (base + len > p.n) len = p.n - base;
    const unsigned FULL = 0xffffffffu;

    double ll = l1[i];
    if (i < 9) ll -= 200;
    const double ll1 = ll;
    bool isbad = bad_in[i] != 0;
    // calczc(d, len, (long long)ll, hz(-20), 400): first sample at or beyond the target, 32 at a time
    double zc = 0.0;
    bool have = false;
    {
        const long long start = (long long)ll;
        const double target = hz(-20);
        if (start >= 0 && start < len) {
            long long end = start + 401;
            if (end > len) end = len;
            const bool rising = d(start) < target;
            long long x = -1;
            // all 401 samples of the scan are requested before the first comparison (one memory round trip
            // instead of one per 32 samples)
            constexpr int NR = 13;                          // 13 * 32 >= 401
            float v32[NR];
            LDD_UNROLL
            for (int r = 0; r < NR; ++r) {
                long long k = start + 32 * r + lane;
                v32[r] = k < end ? p.d05[base + k] : 0.f;
            }
            LDD_UNROLL
            for (int r = 0; r < NR; ++r) {
                if (x < 0) {
                    long long k = start + 32 * r + lane;
                    bool hit = false;
                    if (k < end) { double v = (double)v32[r] + p.ire0; hit = rising ? (v >= target) : (v <= target); }
                    unsigned m = __ballot_sync(FULL, hit);
                    if (m) x = start + 32 * r + (__ffs(m) - 1);
                }
            }
            if (x > 0) {
                double a = d(x - 1) - target, b = d(x) - target;
                zc = (double)(x - 1) + (-a / (-a + b));
                have = true;
            }
        }
    }
    double out = ll;
    if (have && !isbad) {
        out = zc;
        if (i >= 10) {
            long long a1 = (long long)(ll1 - fq * 2), b1 = (long long)(ll1 + fq * 2);
            long long a = (long long)(zc - fq * 1), b = (long long)(zc + fq * 3);
            long long ab = (long long)(zc + fq * 1);
            if (a1 < 0 || a < 0) { if (lane == 0) atomicOr(&p.status[f], 2); a1 = a1 < 0 ? 0 : a1; a = a < 0 ? 0 : a; }
            if (b1 > len) b1 = len;
            if (b > len) b = len;
            double mn = 1e300, mx = -1e300, mn1 = 1e300, mx1 = -1e300, mnb = 1e300, mxb = -1e300;
            const bool staged = (b - a) <= HS_WIN;
            for (long long k = a + lane; k < b; k += 32) {
                double v = d(k);
                if (staged) s_win[warp][k - a] = v;
                mn = v < mn ? v : mn; mx = v > mx ? v : mx;
                if (k >= ab) { mnb = v < mnb ? v : mnb; mxb = v > mxb ? v : mxb; }
            }
            for (long long k = a1 + lane; k < b1; k += 32) { double v = d(k); mn1 = v < mn1 ? v : mn1; mx1 = v > mx1 ? v : mx1; }
            for (int sh = 16; sh > 0; sh >>= 1) {
                double t;
                t = __shfl_xor_sync(FULL, mn, sh); mn = t < mn ? t : mn;
                t = __shfl_xor_sync(FULL, mx, sh); mx = t > mx ? t : mx;
                t = __shfl_xor_sync(FULL, mn1, sh); mn1 = t < mn1 ? t : mn1;
                t = __shfl_xor_sync(FULL, mx1, sh); mx1 = t > mx1 ? t : mx1;
                t = __shfl_xor_sync(FULL, mnb, sh); mnb = t < mnb ? t : mnb;
                t = __shfl_xor_sync(FULL, mxb, sh); mxb = t > mxb ? t : mxb;
            }
            __syncwarp();
            if ((mn < hz(-60) || mx > hz(20)) || (mn1 < hz(-60) || mx1 > hz(100)) || (mnb < hz(-10) || mxb > hz(10))) {
                isbad = true;
            } else if (lane == 0) {
                const long long wl = b - a;
                auto w = [&](long long k) -> double { return staged ? s_win[warp][k] : d(a + k); };
                double tmp[20];
                int n0 = (int)(wl < 20 ? wl : 20);
                for (int k = 0; k < n0; ++k) tmp[k] = w(k);
                double low = np_mean(tmp, n0);
                int n1 = (int)(wl < 100 ? 0 : (wl < 120 ? wl - 100 : 20));
                for (int k = 0; k < n1; ++k) tmp[k] = w(100 + k);
                double high = np_mean(tmp, n1);
                double zc2;
                if (!calczc(w, wl, 0, (low + high) / 2, (int)wl, &zc2)) {
                    atomicOr(&p.status[f], 2);          // reference: TypeError -> field invalid
                    isbad = true;
                } else {
                    zc2 += (double)(long long)zc - fq * 1;
                    if (fabs(zc2 - zc) < fq / 4) out = zc2; else isbad = true;
                }
            }
        }
    } else {
        isbad = true;
    }
    if (lane == 0) {
        if (i < 10) out += fq * 4.72;
        l2[i] = out;
        bad[i] = isbad ? 1 : 0;
    }
}

// The sequential fix-ups of refine_linelocs_hsync (lddecode_core.py:769-787): one warp per field; the tables
// are staged in shared memory with coalesced accesses, lane 0 walks them.
__global__ void __launch_bounds__(32) refine_hsync_fixup_kernel(const HsyncParams p) {
    __shared__ double s_l2[320 + 8];
    __shared__ unsigned char s_bad[320 + 8];
    const int f = blockIdx.x, lane = threadIdx.x;
    if (f >= p.nfields) return;
    const int nll = p.linecount[f] + 4;
    double* l2 = p.linelocs2 + (size_t)f * p.ll_stride;
    const unsigned char* bad = p.linebad_out + (size_t)f * p.ll_stride;
    for (int i = lane; i < nll; i += 32) { s_l2[i] = l2[i]; s_bad[i] = bad[i]; }
    __syncwarp();
    if (lane == 0) {
        const double fq = p.freq;
        for (int i = 11; i < nll; ++i)
            if (s_bad[i]) { double gap = s_l2[i - 1] - s_l2[i - 2]; s_l2[i] = s_l2[i - 1] + gap; }
        const double lo = p.linelen - fq * .2, hi = p.linelen + fq * .2;
        for (int i = 9; i >= 0; --i) {
            double gap = s_l2[i + 1] - s_l2[i];
            if (!(gap >= lo && gap <= hi)) gap = p.linelen;
            s_l2[i] = s_l2[i + 1] - gap;
        }
        for (int i = nll - 10; i < nll; ++i) {
            double gap = s_l2[i] - s_l2[i - 1];
            if (!(gap >= lo && gap <= hi)) gap = p.linelen;
            s_l2[i] = s_l2[i - 1] + gap;
        }
    }
    __syncwarp();
    for (int i = lane; i < nll; i += 32) l2[i] = s_l2[i];
}

// ---------------------------------------------------------------------------------------------
// A9: FieldNTSC.refine_linelocs_burst (lddecode_core.py:1054-1133).
struct BurstParams {
    const float* burst;        // demod_burst plane (absolute Hz)
    long long n;
    double freq;
    int linelen, outwidth;
    const long long* base;
    const int* linecount;
    int ll_stride;
    const double* linelocs_in;   // linelocs2 (first pass) or linelocs3 (second pass)
    double* linelocs_out;
    float* burstlevel;           // [nfields][ll_stride]
    int* status;
};

constexpr int BURST_K = 32;          // same Green's function reach as the TBC kernel
constexpr int BURST_FIRST = 20, BURST_N = 40;
constexpr int BURST_WIN = 192;       // input samples staged per line (>= 2*(FIRST+N)*max step + K + margin)
constexpr int BURST_WARPS = 8;       // warps per CTA of the per-line kernel
constexpr int BURST_LPW = 4;         // lines per warp: the CTA covers 32 lines, one per lane of warp 0 in the scalar phase
constexpr int BURST_LPC = BURST_WARPS * BURST_LPW;
constexpr int BURST_PITCH = BURST_N + 1;   // row pitch of the per-line sample arrays (odd: lanes on different lines do not collide)
constexpr int BURST_VOTE_THREADS = 512;

// w[m] = 6 (g[m-1] - 2 g[m] + g[m+1]), g[k] = r^|k| / (2 sqrt 3) truncated to |k| <= BURST_K: set once per process.
__constant__ double c_burst_taps[2 * BURST_K + 3];

// Per line: the resampled burst, its level, and the two phase candidates of the line (lddecode_core.py:1061-1110).
// ws_phase: [nfields][2][ll_stride].  Two phases per CTA of 32 lines: (A) warp-cooperative, four lines per warp -- stage
// the line's window, second derivatives of the spline where the 40 output samples fall, the 40 samples; (B) the
// scalar statistics and the zero-crossing walk of the reference, one LINE PER LANE of warp 0 (they are sequential per
// line but independent between lines; one lane per warp doing them was most of this kernel's time).
__global__ void __launch_bounds__(32 * BURST_WARPS) burst_lines_kernel(const BurstParams p, double* ws_phase) {
    __shared__ double ys[BURST_WARPS][BURST_WIN];
    __shared__ double Ms[BURST_WARPS][BURST_WIN];
    __shared__ double bas[BURST_LPC][BURST_PITCH];
    __shared__ double sq[BURST_LPC][BURST_PITCH];
    __shared__ unsigned char s_ok[BURST_LPC];
    const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int linecount = p.linecount[f], nll = linecount + 4;
    const long long base = p.base[f];
    const double* lin = p.linelocs_in + (size_t)f * p.ll_stride;
    const int W = p.outwidth;
    // ---- phase A
    for (int q = 0; q < BURST_LPW; ++q) {
        const int slot = warp * BURST_LPW + q;
        const int l = blockIdx.x * BURST_LPC + slot;
        bool ok = false;
        if (l < linecount) {
            const double b = lin[l], e = lin[l + 1];
            const long long ib = (long long)b, ie = (long long)e;
            const int dist = (int)(ie - ib);
            const double fb = b - (double)ib;
            const double step = (((e - b) + fb) - fb) / (double)W;
            const double wowf = (e - b) / (double)p.linelen;
            // input samples needed: x in [fb + 20 step, fb + 59 step] -> indices i0..i1+1, M needs +-(K+1) more
            const int i0 = (int)(fb + BURST_FIRST * step), i1 = (int)(fb + (BURST_FIRST + BURST_N - 1) * step) + 1;
            ok = (b >= 0.0) && dist >= 3 && i1 <= dist && i0 >= 16 && ib + dist + 1 <= p.n;      // i0: |r|^16 = 7e-10 of the line-start boundary term
            if (!ok) {
                // geometry the fast path does not cover (or the reference would raise): flag, leave the line "no burst"
                if (lane == 0) atomicOr(&p.status[f], 4);
            } else {
                // The 40 samples are produced in passes over the staging window: one pass for any line up to ~1.6 x
                // nominal, several (jn outputs each) for a longer one -- the reference resamples whatever span the line
                // table gives it.  A pass over outputs [j0, j1) stages inputs c0 - (K+1) .. c1 + (K+1).
                int jn = BURST_N;
                if ((i1 - i0 + 1) + 2 * (BURST_K + 1) > BURST_WIN) {
                    jn = (int)((double)(BURST_WIN - 2 * (BURST_K + 1) - 3) / step) + 1;
                    if (jn > BURST_N) jn = BURST_N;
                }
                for (int j0 = 0; j0 < BURST_N; j0 += jn) {
                    const int j1 = j0 + jn < BURST_N ? j0 + jn : BURST_N;
                    const int c0 = (int)(fb + (BURST_FIRST + j0) * step), c1 = (int)(fb + (BURST_FIRST + j1 - 1) * step) + 1;
                    const int s0 = c0 - (BURST_K + 1);             // first staged sample (line-relative, may be < 0)
                    const int ns = (c1 - c0 + 1) + 2 * (BURST_K + 1);
                    for (int k = lane; k < ns; k += 32) {
                        long long s = base + ib + s0 + k;
                        s = s < 0 ? 0 : (s >= p.n ? p.n - 1 : s);
                        ys[warp][k] = (double)p.burst[s];
                    }
                    __syncwarp();
                    const int nm = c1 - c0 + 1;                     // M[c0 .. c1]
                    for (int k = lane; k < nm; k += 32) {
                        double acc = 0.0;
                        const double* y = &ys[warp][k];             // y[(c0+k) - (K+1)] is ys[k]
                        for (int m = 0; m <= 2 * (BURST_K + 1); ++m) acc += c_burst_taps[m] * y[m];
                        Ms[warp][k] = acc;
                    }
                    __syncwarp();
                    for (int j = j0 + lane; j < j1; j += 32) {
                        double x = (double)(BURST_FIRST + j) * step + fb;
                        int i = (int)x;
                        double t = x - (double)i, u = 1.0 - t;
                        double Mi = Ms[warp][i - c0], Mj = Ms[warp][i + 1 - c0];
                        double yi = ys[warp][i - s0], yj = ys[warp][i + 1 - s0];
                        double S = Mi * u * u * u / 6.0 + Mj * t * t * t / 6.0 + (yi - Mi / 6.0) * u + (yj - Mj / 6.0) * t;
                        bas[slot][j] = S * wowf;
                    }
                    __syncwarp();
                }
            }
        }
        if (lane == 0) s_ok[slot] = ok ? 1 : 0;
    }
    __syncthreads();
    // ---- phase B: lane <-> line
    if (warp != 0) return;
    const int l = blockIdx.x * BURST_LPC + lane;
    if (l >= nll) return;
    float* level = p.burstlevel + (size_t)f * p.ll_stride;
    double* ph0 = ws_phase + (size_t)f * 2 * p.ll_stride;
    double* ph1 = ph0 + p.ll_stride;
    float lev_out = 0.f;
    double p0 = 0.0, p1 = 0.0;
    const double hz_ire = 1700000.0 / 140.0;
    if (l < linecount && s_ok[lane]) {
        double* ba = bas[lane];
        double* dev = sq[lane];
        double mean = np_mean(ba, BURST_N);
        for (int k = 0; k < BURST_N; ++k) ba[k] -= mean;
        double mx = 0.0;
        for (int k = 0; k < BURST_N; ++k) mx = fabs(ba[k]) > mx ? fabs(ba[k]) : mx;
        float lev = (float)mx;
        // np.std(ba): mean again, deviations, pairwise sum of squares
        double m2 = np_mean(ba, BURST_N);
        for (int k = 0; k < BURST_N; ++k) { double q = ba[k] - m2; dev[k] = q * q; }
        double sd = sqrt(np_sum(dev, BURST_N) / (double)BURST_N);
        const float hz_ire32 = (float)hz_ire;
        if (!((lev / hz_ire32) > 30.f || (sd / hz_ire) < 3)) {
            lev_out = lev;
            const double thr = (double)(lev * 0.6f);
            // phase offsets of the falling / rising crossings, kept in the (now free) deviation row: F from the
            // front, T from the back
            double* offF = dev;
            double offT[BURST_N / 2 + 1];
            int nF = 0, nT = 0;
            auto dat = [&](long long k) -> double { return ba[k]; };
            int bi = 0;
            while (bi < BURST_N) {
                if (fabs(ba[bi]) > thr) {
                    double zc;
                    if (calczc(dat, BURST_N, bi, 0.0, 10, &zc)) {
                        double off = zc - ((floor(zc / 4) * 4) - 1);
                        if (off > 3.5) off -= 4;
                        if (ba[bi] > 0) { if (nT < BURST_N / 2 + 1) offT[nT] = off; ++nT; } else { if (nF < BURST_N) offF[nF] = off; ++nF; }
                        bi = (int)zc;
                    }
                }
                ++bi;
            }
            if (nT > BURST_N / 2 + 1) nT = BURST_N / 2 + 1;
            if (nF >= 3 && nT >= 3) {
                double mF = np_mean(offF + 1, nF - 2), mT = np_mean(offT + 1, nT - 2);
                if (l % 2) { p0 = 2 - mT; p1 = 2 - mF; }
                else { p0 = 2 - mF; p1 = 2 - mT; }
            }
        }
    }
    level[l] = lev_out; ph0[l] = p0; ph1[l] = p1;
}

// Per field: the phase group vote (medians of both candidate columns over the lines that produced one,
// lddecode_core.py:1112-1117), the shifted line positions and the interpolation of lines without burst.
__global__ void __launch_bounds__(BURST_VOTE_THREADS) burst_vote_kernel(const BurstParams p, const double* ws_phase) {
    __shared__ double phase[2][512];
    __shared__ double col[2][512];
    __shared__ double srt[2][512];
    __shared__ int s_group, s_nc;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int linecount = p.linecount[f], nll = linecount + 4;
    const double* lin = p.linelocs_in + (size_t)f * p.ll_stride;
    double* lout = p.linelocs_out + (size_t)f * p.ll_stride;
    float* level = p.burstlevel + (size_t)f * p.ll_stride;
    const double* ph = ws_phase + (size_t)f * 2 * p.ll_stride;
    for (int l = tid; l < nll; l += blockDim.x) { phase[0][l] = ph[l]; phase[1][l] = ph[p.ll_stride + l]; }
    __syncthreads();
    if (tid == 0) {
        int nc = 0;
        for (int l = 0; l < nll; ++l)
            if (phase[0][l] != 0 || phase[1][l] != 0) { col[0][nc] = phase[0][l]; col[1][nc] = phase[1][l]; ++nc; }
        s_nc = nc;
    }
    __syncthreads();
    const int nc = s_nc;
    // rank-sort both columns in parallel (each thread places one element), thread 0 reads the middles
    for (int t = tid; t < 2 * nc; t += blockDim.x) {
        const double* c = t < nc ? col[0] : col[1];
        double* d = t < nc ? srt[0] : srt[1];
        const int me = t < nc ? t : t - nc;
        const double v = c[me];
        int rank = 0;
        for (int i = 0; i < nc; ++i) rank += (c[i] < v) || (c[i] == v && i < me);
        d[rank] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int group = 1;
        if (nc > 0) {
            double m0 = (nc & 1) ? srt[0][nc / 2] : (srt[0][nc / 2 - 1] + srt[0][nc / 2]) / 2.0;
            double m1 = (nc & 1) ? srt[1][nc / 2] : (srt[1][nc / 2 - 1] + srt[1][nc / 2]) / 2.0;
            group = fabs(m0) < fabs(m1) ? 0 : 1;
        }
        s_group = group;
    }
    __syncthreads();
    const int group = s_group;
    const double k4 = p.freq / (4.0 * 315.0 / 88.0);
    // the shifted positions go through shared memory (col[0] is free) for the sequential interpolation pass
    double* lo = col[0];
    float* lv_s = (float*)col[1];
    for (int l = tid; l < nll; l += blockDim.x) {
        float lv = level[l];
        if ((l & 1) == (group & 1)) lv = -lv;          // burstlevel[phasegroup::2] = -burstlevel[phasegroup::2]
        double adj = phase[group][l];
        double v = lin[l];
        if (fabs(adj) > 2) lv = 0.f; else v -= adj * k4 * 1;
        lv_s[l] = lv;
        lo[l] = v;
    }
    __syncthreads();
    if (tid == 0) {
        for (int l = 2; l < nll - 1; ++l)
            if (lv_s[l] == 0.f) lo[l] = (lo[l - 1] + lo[l + 1]) / 2;
    }
    __syncthreads();
    for (int l = tid; l < nll; l += blockDim.x) { level[l] = lv_s[l]; lout[l] = lo[l]; }
}

// ---------------------------------------------------------------------------------------------
// A10: FieldPAL.refine_linelocs_pilot (lddecode_core.py:962-1021).
struct PilotParams {
    const float* demod;        // relative to ire0
    const float* d05;          // relative to ire0
    long long n;
    double freq;
    int linelen;
    const long long* base;
    const int* linecount;
    int ll_stride;
    const double* linelocs_in;
    double* linelocs_out;
    int* status;
};

constexpr int PILOT_MAXOFF = 48;     // zero crossings kept per line (4.7 us of a 3.75 MHz pilot: ~17)
constexpr int PILOT_MAXLEN = 192;    // samples in 4.7 us (<= 40 MSPS)

constexpr int PILOT_WARPS = 8;       // lines per CTA of the per-line kernel
constexpr int PILOT_MED_THREADS = 1024;

// Per line (one warp each, grid over all lines of all fields): the zero crossings of the pilot in the
// 4.7 us before the line location, their phase offsets (sorted, for the line's median) and the count.
__global__ void __launch_bounds__(32 * PILOT_WARPS) pilot_lines_kernel(const PilotParams p, double* ws_offsets /*[nfields][ll_stride][MAXOFF]*/,
                                                                      int* ws_count /*[nfields][ll_stride]*/) {
    __shared__ double s_win[PILOT_WARPS][PILOT_MAXLEN];     // pilot = flip(demod - demod_05) of this warp's line
    __shared__ double s_off[PILOT_WARPS][PILOT_MAXOFF];
    __shared__ int s_cnt[PILOT_WARPS];
    const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.x * PILOT_WARPS + warp;
    const int nll = p.linecount[f] + 4;
    if (l >= nll) return;
    double* s_pil = s_win[warp];
    double* so = s_off[warp];
    const long long base = p.base[f];
    const double* lin = p.linelocs_in + (size_t)f * p.ll_stride;
    double* my = ws_offsets + ((size_t)f * p.ll_stride + l) * PILOT_MAXOFF;
    int* cnt = ws_count + (size_t)f * p.ll_stride;
    const double fq = p.freq;
    long long a = (long long)(lin[l] - fq * 4.7), b = (long long)lin[l];
    int len = (int)(b - a);
    if (len <= 0 || len > PILOT_MAXLEN || base + a < 0 || base + b > p.n) {
        if (lane == 0) { atomicOr(&p.status[f], 8); cnt[l] = 0; }
        return;
    }
    // coalesced staging; the two ire0 offsets cancel, the subtraction is done in float64 like the reference
    for (int i = lane; i < len; i += 32) {
        long long s = base + b - 1 - i;
        s_pil[i] = ((double)p.demod[s]) - ((double)p.d05[s]);
    }
    __syncwarp();
    double adjfreq = fq;
    if (l > 1) adjfreq /= (lin[l] - lin[l - 1]) / p.linelen;
    // The reference walks the window sample by sample: at a sample in -300k..-100k Hz it looks for the first
    // sample >= 0 among the next 10, records that zero crossing and continues behind it.  Every run of negative
    // samples therefore yields its closing crossing x exactly when one of the (up to 10) run samples before x
    // lies in that band -- a per-crossing test, done here by all lanes at once; the ballot keeps the order.
    // (If an interpolated crossing lands exactly on x the reference also skips sample x+1: lane 0 then redoes
    // the line with the sequential walk.)
    const unsigned FULL = 0xffffffffu;
    int nfound = 0;
    bool redo = false;
    for (int x0 = 0; x0 < len; x0 += 32) {
        const int x = x0 + lane;
        bool hit = false;
        double off = 0.0;
        if (x >= 1 && x < len && s_pil[x] >= 0.0 && s_pil[x - 1] < 0.0) {
            for (int j = x - 1; j >= 0 && j >= x - 10; --j) {
                const double v = s_pil[j];
                if (!(v < 0.0)) break;
                if (v >= -300000 && v <= -100000) { hit = true; break; }
            }
            if (hit) {
                const double a = s_pil[x - 1], bb = s_pil[x];
                const double zc = (double)(x - 1) + (-a / (-a + bb));
                if (zc >= (double)x) redo = true;
                const double zcp = zc / (adjfreq / 3.75);
                off = zcp - floor(zcp);
            }
        }
        const unsigned m = __ballot_sync(FULL, hit);
        if (hit) {
            const int at = nfound + __popc(m & ((1u << lane) - 1u));
            if (at < PILOT_MAXOFF) so[at] = off;
        }
        nfound += __popc(m);
    }
    redo = __any_sync(FULL, redo);
    if (redo) {
        __syncwarp();
        if (lane == 0) {
            int n = 0;
            auto pil = [&](long long i) -> double { return s_pil[i]; };
            int i = 0;
            while (i < len) {
                double v = s_pil[i];
                if (v >= -300000 && v <= -100000) {
                    double zc;
                    if (calczc(pil, len, i, 0.0, 10, &zc)) {
                        double zcp = zc / (adjfreq / 3.75);
                        if (n < PILOT_MAXOFF) so[n] = zcp - floor(zcp);
                        ++n;
                        i = (int)(zc + 1);
                    }
                }
                ++i;
            }
            s_cnt[warp] = n;
        }
        __syncwarp();
        nfound = s_cnt[warp];
        __syncwarp();
    }
    if (lane == 0) {
        int n = nfound;
        if (n > PILOT_MAXOFF) { atomicOr(&p.status[f], 8); n = PILOT_MAXOFF; }
        // "if len(offsets) >= 3": the dict has l+1 entries at this point; offsets[1:-1] are kept
        s_cnt[warp] = (l + 1 >= 3 && n >= 2) ? n - 2 : 0;
    }
    __syncwarp();
    const int n = s_cnt[warp];
    // rank sort of so[1 .. n] into the workspace (ties keep their order)
    for (int i = lane; i < n; i += 32) {
        const double x = so[1 + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double y = so[1 + j];
            rank += (y < x) || (y == x && j < i);
        }
        my[rank] = x;
    }
    if (lane == 0) cnt[l] = n;
}

// Per field: np.median over all kept offsets of the field (exact order statistics by histogram selection over
// the values gathered in shared memory), the target phase, and the shifted line locations.
__global__ void __launch_bounds__(PILOT_MED_THREADS) pilot_median_kernel(const PilotParams p, const double* ws_offsets, const int* ws_count) {
    LDD_DYN_SMEM(psm);
    double* s_val = (double*)psm;                         // [npow2]
    __shared__ int s_start[320 + 8];
    __shared__ int s_total;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int nll = p.linecount[f] + 4;
    const double* lin = p.linelocs_in + (size_t)f * p.ll_stride;
    double* lout = p.linelocs_out + (size_t)f * p.ll_stride;
    const double* offs = ws_offsets + (size_t)f * p.ll_stride * PILOT_MAXOFF;
    const int* cnt = ws_count + (size_t)f * p.ll_stride;
    const double fq = p.freq;
    __shared__ int s_cnt[320 + 8];
    __shared__ int s_hist[4096];
    __shared__ double s_list[1024];
    __shared__ int s_wsum[32];
    __shared__ int s_n, s_bin, s_before, s_inrange;
    __shared__ double s_result;
    for (int l = tid; l < nll; l += PILOT_MED_THREADS) s_cnt[l] = cnt[l];
    __syncthreads();
    if (tid < 32) {
        // exclusive prefix of the line counts (nll <= 320 + 4: ten lines per lane)
        int run = 0;
        for (int l0 = 0; l0 < nll; l0 += 32) {
            const int l = l0 + tid;
            const int c = l < nll ? s_cnt[l] : 0;
            int incl = c;
            for (int d = 1; d < 32; d <<= 1) {
                int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (tid >= d) incl += up;
            }
            if (l < nll) s_start[l] = run + incl - c;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) s_total = run;
    }
    __syncthreads();
    const int total = s_total;
    double tgt = 0;
    if (total > 0) {
        // flat, coalesced sweep over the [nll][MAXOFF] workspace; the loads do not depend on the counts
        const int nflat = nll * PILOT_MAXOFF;
        LDD_UNROLL
        for (int it = 0; it < (320 + 8) * PILOT_MAXOFF / PILOT_MED_THREADS + 1; ++it) {
            const int e = tid + it * PILOT_MED_THREADS;
            const int ec = e < nflat ? e : 0;
            const double v = offs[ec];
            const int l = ec / PILOT_MAXOFF, q = ec - l * PILOT_MAXOFF;
            if (e < nflat && q < s_cnt[l]) s_val[s_start[l] + q] = v;
        }
        __syncthreads();
        // k-th smallest of s_val[0 .. total): multi-level histogram selection (the offsets are fractional parts in
        // [0, 1)): 4096 bins over the current range, descend into the bin that holds rank k until it holds few
        // enough values to rank them directly.  Binning decides membership at every level, so edges are consistent.
        auto kth = [&](int k) -> double {
            double lo = 0.0, width = 1.0;
            int kk = k;
            for (int level = 0; level < 6; ++level) {
                for (int i = tid; i < 4096; i += PILOT_MED_THREADS) s_hist[i] = 0;
                if (tid == 0) s_n = 0;
                __syncthreads();
                const double scale = 4096.0 / width;
                for (int i = tid; i < total; i += PILOT_MED_THREADS) {
                    const double v = s_val[i];
                    if (v >= lo && v < lo + width) {
                        int bn = (int)((v - lo) * scale);
                        bn = bn > 4095 ? 4095 : bn;
                        atomicAdd(&s_hist[bn], 1);
                    }
                }
                __syncthreads();
                // block-wide exclusive prefix over the bins, four consecutive bins per thread
                const int h0 = s_hist[4 * tid], h1 = s_hist[4 * tid + 1], h2 = s_hist[4 * tid + 2], h3 = s_hist[4 * tid + 3];
                const int part = h0 + h1 + h2 + h3;
                int incl = part;
                for (int d = 1; d < 32; d <<= 1) {
                    int up = __shfl_up_sync(0xffffffffu, incl, d);
                    if ((tid & 31) >= d) incl += up;
                }
                if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
                if (tid == 0) { s_bin = 4095; s_before = -1; }
                __syncthreads();
                if (tid < 32) {
                    int w = s_wsum[tid], wi = w;
                    for (int d = 1; d < 32; d <<= 1) {
                        int up = __shfl_up_sync(0xffffffffu, wi, d);
                        if (tid >= d) wi += up;
                    }
                    s_wsum[tid] = wi - w;
                    if (tid == 31) s_inrange = wi;
                }
                __syncthreads();
                const int excl = s_wsum[tid >> 5] + incl - part;
                if (kk >= excl && kk < excl + part) {
                    int cum = excl, bn = 4 * tid;
                    if (cum + h0 <= kk) { cum += h0; ++bn; if (cum + h1 <= kk) { cum += h1; ++bn; if (cum + h2 <= kk) { cum += h2; ++bn; } } }
                    s_bin = bn;
                    s_before = cum;
                }
                __syncthreads();
                if (s_before < 0) {                       // rank beyond the range's population (cannot happen for 0 <= k < total)
                    if (tid == 0) s_before = s_inrange - s_hist[4095];
                    __syncthreads();
                }
                const int bin = s_bin, inbin = s_hist[bin], before = s_before;
                const double blo = lo + (double)bin / scale, bhi = lo + (double)(bin + 1) / scale;
                if (inbin <= 1024 || level == 5) {
                    // gather the bin's members, rank them by counting (ties broken by position)
                    for (int i = tid; i < total; i += PILOT_MED_THREADS) {
                        const double v = s_val[i];
                        if (v >= lo && v < lo + width) {
                            int bn = (int)((v - lo) * scale);
                            bn = bn > 4095 ? 4095 : bn;
                            if (bn == bin) { int at = atomicAdd(&s_n, 1); if (at < 1024) s_list[at] = v; }
                        }
                    }
                    __syncthreads();
                    const int n = s_n < 1024 ? s_n : 1024;
                    int r = kk - before;
                    r = r < 0 ? 0 : (r >= n ? n - 1 : r);
                    if (tid < n) {
                        const double x = s_list[tid];
                        int rank = 0;
                        for (int j = 0; j < n; ++j) {
                            const double y = s_list[j];
                            rank += (y < x) || (y == x && j < tid);
                        }
                        if (rank == r) s_result = x;
                    }
                    __syncthreads();
                    const double result = s_result;
                    __syncthreads();
                    return result;
                }
                kk -= before;
                lo = blo;
                width = bhi - blo;
                __syncthreads();
            }
            return 0.0;
        };
        const double med = (total & 1) ? kth(total / 2) : (kth(total / 2 - 1) + kth(total / 2)) / 2.0;
        if (med >= 0.25 && med <= 0.75) tgt = .5;
    }
    for (int l = tid; l < nll; l += PILOT_MED_THREADS) {
        double v = lin[l];
        const int n = s_cnt[l];
        if (n > 0) {
            const double* my = offs + (size_t)l * PILOT_MAXOFF;    // sorted by pilot_lines_kernel
            const double med = (n & 1) ? my[n / 2] : (my[n / 2 - 1] + my[n / 2]) / 2.0;
            v += (tgt - med) * (fq / 3.75) * .25;
        }
        lout[l] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Host: A7 + the early-outs of Field.__init__.
namespace {

int fail_msg(ldd_handle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}

double median_of(std::vector<double> v) {
    size_t n = v.size();
    if (n == 0) return NAN;
    std::nth_element(v.begin(), v.begin() + n / 2, v.end());
    double hi = v[n / 2];
    if (n & 1) return hi;
    double lo = *std::max_element(v.begin(), v.begin() + n / 2);
    return (lo + hi) / 2.0;
}

struct Locator {
    const long long* pk;     // window-relative peak positions
    const double* val;
    int np;
    long long wlen;
    int linelen;
    bool pal;
    double med = 0, tol = 0;

    bool regular(long long k) const {          // is_regular_hsync, lddecode_core.py:534-542
        if (k < 0) k += np;                    // python negative index (never reached in practice)
        if (k < 0 || k >= np) return false;
        if (pk[k] > wlen) return false;
        return val[k] >= med - tol && val[k] <= med + tol;
    }

    // determine_field (lddecode_core.py:544-588) for peaknum >= 11: the last regular hsync before the vertical interval and
    // the parity vote from the line gaps either side of it.  False when no regular hsync precedes it (line0 is None: the
    // reference's caller skips such a candidate).
    bool field_vote(int i, long long* line0_out, int* vote_out) const {
        int vote = 0;
        long long line0 = -1;
        bool have0 = false;
        for (int k = i - 1; k > i - 20; --k) {
            if (regular(k)) {
                line0 = k; have0 = true;
                long long kk = k < 0 ? k + np : k;
                if (kk + 1 < np && (pk[kk + 1] - pk[kk]) > linelen * .75) vote -= 1;
                break;
            }
        }
        for (int k = i; k < i + 20; ++k) {
            if (regular(k)) {
                if (k >= 1 && (pk[k] - pk[k - 1]) > linelen * .75) vote += pal ? -1 : 1;
                break;
            }
        }
        if (pal) vote += 1;
        *line0_out = line0;
        *vote_out = vote;
        return have0;
    }
};

}  // namespace
}  // namespace ldd

using namespace ldd;

extern "C" int ldd_field_locate(ldd_handle* h, const long long* peaks, const double* vals, int npeaks,
                                long long window_len, long long start, ldd_field* out,
                                double* linelocs1, unsigned char* linebad, int ll_cap) {
    // a window without a single peak (silence, lead-in, an unreadable stretch) is a legitimate input: its list is empty
    if (!h || !out || npeaks < 0 || (npeaks > 0 && (!peaks || !vals))) return LDD_EINVAL;
    const ldd_config& c = h->cfg;
    memset(out, 0, sizeof *out);
    out->npeaks = npeaks;
    const int L = c.linelen;
    Locator lc{peaks, vals, npeaks, window_len, L, c.system == LDD_SYSTEM_PAL};
    // --- determine_vsyncs (lddecode_core.py:590-636)
    std::vector<std::array<long long, 3>> vs;
    if (npeaks >= 200) {
        std::vector<double> lv;
        for (int i = 0; i < npeaks; ++i)
            if (vals[i] >= 0.6 && vals[i] <= 0.8) lv.push_back(vals[i]);
        lc.med = median_of(lv);
        double sd = NAN;
        if (!lv.empty()) {
            double mean = np_sum_host(lv.data(), (int)lv.size()) / (double)lv.size();
            std::vector<double> dev(lv.size());
            for (size_t i = 0; i < lv.size(); ++i) { double q = lv[i] - mean; dev[i] = q * q; }
            sd = std::sqrt(np_sum_host(dev.data(), (int)dev.size()) / (double)dev.size());
        }
        lc.tol = std::max(sd * 2, .01);          // python max(nan, .01) == nan; std::max(nan, .01) == nan too (first arg kept)
        out->med_hsync = lc.med;
        out->hsync_tolerance = lc.tol;
        double prev = 1.0;
        for (int i = 0; i < npeaks; ++i) {
            double v = vals[i];
            if (v > .9 && prev < lc.med - lc.tol * 2) {
                // determine_field (lddecode_core.py:544-588)
                if (i < 11) { out->stage = LDD_FIELD_CRASH; return LDD_OK; }      // the reference raises TypeError here
                int vote = 0;
                long long line0 = -1;
                const bool have0 = lc.field_vote(i, &line0, &vote);
                if (have0) vs.push_back({(long long)i, line0, (long long)vote});
            }
            prev = v;
        }
        if (vs.size() >= 2) {
            std::vector<std::array<long long, 3>> raw = vs;
            for (size_t i = 0; i < vs.size(); ++i) {
                if (vs[i][2] == 0) {
                    vs[i][1] = -1;
                    if (i + 1 < vs.size() && raw[i + 1][2] != 0) vs[i][2] = -vs[i + 1][2];
                    else if (i >= 1 && raw[i - 1][2] != 0) vs[i][2] = -vs[i - 1][2];
                }
                if (vs[i][1] <= 0) vs[i][1] = vs[i][0] - (lc.pal ? 6 : 7);
                vs[i][2] = vs[i][2] < 0 ? 1 : 0;
            }
        }
    }
    out->nvsyncs = (int)vs.size();
    for (size_t i = 0; i < vs.size() && i < 4; ++i)
        for (int q = 0; q < 3; ++q) out->vsyncs[i][q] = (int)vs[i][q];
    // --- Field.__init__ early-outs (lddecode_core.py:909-926)
    if (vs.empty()) {
        out->stage = LDD_FIELD_NOVSYNC;
        out->nextfieldoffset = start + (long long)L * 200;
        return LDD_OK;
    }
    auto pyidx = [&](long long k) -> long long { return k < 0 ? k + npeaks : k; };
    if (vs.size() == 1 || npeaks < vs[1][1] + 4) {
        long long k = pyidx(vs[0][1] - 10);
        long long jump = (k >= 0 && k < npeaks) ? peaks[k] : 0;
        out->stage = LDD_FIELD_SHORT;
        out->nextfieldoffset = jump != 0 ? start + jump : start + (long long)L * 240;
        return LDD_OK;
    }
    {
        long long k = pyidx(vs[1][1] - 10);
        if (k < 0 || k >= npeaks) { out->stage = LDD_FIELD_CRASH; return LDD_OK; }
        out->nextfieldoffset = peaks[k];
        out->tbcstart = peaks[k];
    }
    // the raw vote sign of vsync 0 decides the parity (after the loop above it is 0/1 when there were >= 2)
    out->istop = (int)vs[0][2];
    const int frame_lines = lc.pal ? 625 : 525;
    out->linecount = frame_lines / 2 + (out->istop ? 1 : 0);
    const int nll = out->linecount + 4;
    if (!linelocs1 || !linebad || ll_cap < nll) return LDD_ECAP;
    // --- compute_linelocs (lddecode_core.py:638-713)
    std::vector<double> found(nll + 64, 0.0);
    std::vector<char> has(nll + 64, 0);
    // line numbers can be negative (lines before line0) or beyond the table: keep a map for the ones we need
    std::vector<std::pair<long long, double>> extra;          // (linenum, loc) outside [0, nll+63]
    auto setloc = [&](long long n, double v) {
        if (n >= 0 && n < (long long)found.size()) { found[n] = v; has[n] = 1; }
        else {
            for (auto& e : extra) if (e.first == n) { e.second = v; return; }
            extra.push_back({n, v});
        }
    };
    auto getloc = [&](long long n, double* v) -> bool {
        if (n >= 0 && n < (long long)found.size()) { if (has[n]) { *v = found[n]; return true; } return false; }
        for (auto& e : extra) if (e.first == n) { *v = e.second; return true; }
        return false;
    };
    std::vector<double> lens{(double)L};
    long long prev_i = -1, prev_n = 0;
    const long long v0line0 = pyidx(vs[0][1]);
    if (v0line0 < 0 || v0line0 >= npeaks) { out->stage = LDD_FIELD_BADLINES; return LDD_OK; }
    // np.median(linelens[-25:]) is recomputed by the reference on every iteration but only used for
    // the first line and for irregular gaps: evaluate it lazily (same value, the list is unchanged
    // between the top of the iteration and the point of use).
    auto medlen_now = [&]() -> double {
        size_t m = lens.size() < 25 ? lens.size() : 25;
        double tmp[25];
        std::copy(lens.end() - m, lens.end(), tmp);
        std::sort(tmp, tmp + m);
        return (m & 1) ? tmp[m / 2] : (tmp[m / 2 - 1] + tmp[m / 2]) / 2.0;
    };
    for (long long i = 0; i < vs[1][1]; ++i) {
        if (!lc.regular(i)) continue;
        long long n;
        if (prev_i >= 0) {
            long long gap = peaks[i] - peaks[prev_i];
            double ratio = (double)gap / (double)L;
            if (ratio >= .98 && ratio <= 1.02) { lens.push_back((double)gap); n = prev_n + 1; }
            else n = prev_n + (long long)std::nearbyint((double)gap / medlen_now());
        } else {
            n = (long long)std::nearbyint((double)(peaks[i] - peaks[v0line0]) / medlen_now());
        }
        setloc(n, (double)peaks[i]);
        prev_i = i; prev_n = n;
    }
    std::vector<double> filled(nll + 1, 0.0);
    std::vector<char> wasfound(nll + 1, 0);
    for (int l = 1; l < out->linecount + 5; ++l) {
        double v;
        if (getloc(l, &v)) { filled[l] = v; wasfound[l] = 1; continue; }
        bool hb = false, ha = false;
        long long before = 0, after = 0;
        double vb = 0, va = 0;
        for (long long i = l; i > -10; --i) if (getloc(i, &vb)) { before = i; hb = true; break; }
        for (long long i = l; i < out->linecount + 1; ++i) if (getloc(i, &va)) { after = i; ha = true; break; }
        if (!hb) {
            if (!ha) { out->stage = LDD_FIELD_BADLINES; return LDD_OK; }      // reference: KeyError -> "unable to decode frame"
            filled[l] = va - (double)L * (double)(after - l);
        } else if (ha) {
            double avg = (va - vb) / (double)(after - before);
            filled[l] = vb + avg * (double)(l - before);
        } else {
            // avglen = linelocs[prev_valid] - linelocs2[prev_valid - 1]
            double pm1;
            long long q = before - 1;
            if (q >= 1 && q <= nll) pm1 = filled[q];
            else if (!getloc(q, &pm1)) { out->stage = LDD_FIELD_BADLINES; return LDD_OK; }
            double avg = vb - pm1;
            filled[l] = vb + avg * (double)(l - before);
        }
    }
    for (int l = 1; l < out->linecount + 5; ++l) {
        linelocs1[l - 1] = filled[l];
        linebad[l - 1] = wasfound[l] ? 0 : 1;
    }
    for (int i = 0; i < 10 && i < nll; ++i) linebad[i] = 0;
    out->stage = LDD_FIELD_LOCATED;
    return LDD_OK;
}

extern "C" int ldd_field_vote(ldd_handle* h, const long long* peaks, const double* vals, int npeaks, long long window_len,
                              double med_hsync, double hsync_tolerance, int peaknum, int* line0, int* vote) {
    if (!h || !peaks || !vals || !line0 || !vote || npeaks < 0) return LDD_EINVAL;
    *line0 = -1; *vote = 0;
    if (peaknum < 11) return LDD_OK;                  // the reference returns None
    Locator lc{peaks, vals, npeaks, window_len, h->cfg.linelen, h->cfg.system == LDD_SYSTEM_PAL};
    lc.med = med_hsync; lc.tol = hsync_tolerance;
    long long l0 = -1;
    lc.field_vote(peaknum, &l0, vote);
    *line0 = (int)l0;
    return LDD_OK;
}

extern "C" int ldd_refine_hsync(ldd_handle* h, const float* d05_dev, long long n, const long long* base_dev,
                                const long long* winlen_dev, const int* linecount_dev, int nfields, int ll_stride,
                                const double* linelocs1_dev, const unsigned char* linebad_dev, double* linelocs2_dev,
                                unsigned char* linebad_out_dev, int* status_dev, void* stream) {
    if (!h || !d05_dev || !base_dev || !winlen_dev || !linecount_dev || !linelocs1_dev || !linebad_dev || !linelocs2_dev ||
        !linebad_out_dev || !status_dev) return LDD_EINVAL;
    if (nfields <= 0) return LDD_OK;
    if (ll_stride > 320 + 8) return LDD_EINVAL;      // shared-memory tables of the fix-up kernel
    HsyncParams p;
    p.d05 = d05_dev; p.n = n; p.ire0 = h->cfg.ire0; p.hz_ire = h->cfg.hz_ire; p.freq = h->cfg.freq_hz / 1e6;
    p.linelen = h->cfg.linelen; p.base = base_dev; p.winlen = winlen_dev; p.linecount = linecount_dev; p.ll_stride = ll_stride;
    p.linelocs1 = linelocs1_dev; p.linebad_in = linebad_dev; p.linelocs2 = linelocs2_dev; p.linebad_out = linebad_out_dev;
    p.status = status_dev;
    p.nfields = nfields;
    const int maxll = ll_stride;        // linecount + 4 <= ll_stride
    LDD_LAUNCH(refine_hsync_kernel, dim3((maxll + HS_WARPS - 1) / HS_WARPS, nfields), dim3(32 * HS_WARPS), 0, (cudaStream_t)stream, p);
    int rc = launch_status(h, "refine_hsync_kernel");
    if (rc) return rc;
    LDD_LAUNCH(refine_hsync_fixup_kernel, dim3(nfields), dim3(32), 0, (cudaStream_t)stream, p);
    return launch_status(h, "refine_hsync_fixup_kernel");
}

extern "C" int ldd_refine_burst(ldd_handle* h, const float* burst_dev, long long n, const long long* base_dev,
                                const int* linecount_dev, int nfields, int ll_stride, const double* linelocs_in_dev,
                                double* linelocs_out_dev, float* burstlevel_dev, int* status_dev, void* stream) {
    if (!h || !burst_dev || !base_dev || !linecount_dev || !linelocs_in_dev || !linelocs_out_dev || !burstlevel_dev || !status_dev)
        return LDD_EINVAL;
    if (nfields <= 0) return LDD_OK;
    if (ll_stride > 512) return LDD_EINVAL;
    BurstParams p;
    p.burst = burst_dev; p.n = n; p.freq = h->cfg.freq_hz / 1e6; p.linelen = h->cfg.linelen; p.outwidth = h->cfg.outlinelen;
    p.base = base_dev; p.linecount = linecount_dev; p.ll_stride = ll_stride; p.linelocs_in = linelocs_in_dev;
    p.linelocs_out = linelocs_out_dev; p.burstlevel = burstlevel_dev; p.status = status_dev;
    cudaStream_t st = (cudaStream_t)stream;
    size_t need = (size_t)nfields * 2 * ll_stride * sizeof(double) + 64;
    if (need > h->pilot_ws_bytes) {          // the per-line workspace is shared with the PAL pilot kernels (one system per handle)
        if (h->pilot_ws) { cudaStreamSynchronize(st); cudaFree(h->pilot_ws); h->pilot_ws = nullptr; }
        if (cudaMalloc(&h->pilot_ws, need) != cudaSuccess) { h->pilot_ws_bytes = 0; h->err = "cudaMalloc burst workspace"; return LDD_ENOMEM; }
        h->pilot_ws_bytes = need;
    }
    if (!h->burst_taps_set) {
        const double r = -0.26794919243112270647, c = 0.28867513459481288225;
        double taps[2 * BURST_K + 3];
        auto g = [&](int q) -> double { int a = q < 0 ? -q : q; return a > BURST_K ? 0.0 : c * pow(r, (double)a); };
        for (int m = 0; m <= 2 * (BURST_K + 1); ++m) { int k = m - (BURST_K + 1); taps[m] = 6.0 * (g(k - 1) - 2.0 * g(k) + g(k + 1)); }
        cudaMemcpyToSymbol(c_burst_taps, taps, sizeof taps);
        cudaStreamSynchronize((cudaStream_t)0);      // staged from pageable memory; the kernels run on non-blocking streams
        h->burst_taps_set = true;
    }
    double* ws = (double*)h->pilot_ws;
    LDD_LAUNCH(burst_lines_kernel, dim3((ll_stride + BURST_LPC - 1) / BURST_LPC, nfields), dim3(32 * BURST_WARPS), 0, st, p, ws);
    LDD_LAUNCH(burst_vote_kernel, dim3(nfields), dim3(BURST_VOTE_THREADS), 0, st, p, (const double*)ws);
    return launch_status(h, "burst kernels");
}

extern "C" int ldd_refine_pilot(ldd_handle* h, const float* demod_dev, const float* d05_dev, long long n,
                                const long long* base_dev, const int* linecount_dev, int nfields, int ll_stride,
                                const double* linelocs_in_dev, double* linelocs_out_dev, int* status_dev, void* stream) {
    if (!h || !demod_dev || !d05_dev || !base_dev || !linecount_dev || !linelocs_in_dev || !linelocs_out_dev || !status_dev)
        return LDD_EINVAL;
    if (nfields <= 0) return LDD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    size_t need = (size_t)nfields * ll_stride * (PILOT_MAXOFF * sizeof(double) + sizeof(int)) + 64;
    if (need > h->pilot_ws_bytes) {
        if (h->pilot_ws) { cudaStreamSynchronize(st); cudaFree(h->pilot_ws); h->pilot_ws = nullptr; }
        if (cudaMalloc(&h->pilot_ws, need) != cudaSuccess) { h->pilot_ws_bytes = 0; h->err = "cudaMalloc pilot workspace"; return LDD_ENOMEM; }
        h->pilot_ws_bytes = need;
    }
    double* offs = (double*)h->pilot_ws;
    int* cnt = (int*)((char*)h->pilot_ws + (size_t)nfields * ll_stride * PILOT_MAXOFF * sizeof(double));
    PilotParams p;
    p.demod = demod_dev; p.d05 = d05_dev; p.n = n; p.freq = h->cfg.freq_hz / 1e6; p.linelen = h->cfg.linelen;
    p.base = base_dev; p.linecount = linecount_dev; p.ll_stride = ll_stride; p.linelocs_in = linelocs_in_dev;
    p.linelocs_out = linelocs_out_dev; p.status = status_dev;
    if (ll_stride > 320 + 8) return LDD_EINVAL;
    LDD_LAUNCH(pilot_lines_kernel, dim3((ll_stride + PILOT_WARPS - 1) / PILOT_WARPS, nfields), dim3(32 * PILOT_WARPS), 0, st, p, offs, cnt);
    // all kept offsets of a field
    const size_t psmem = (size_t)ll_stride * PILOT_MAXOFF * sizeof(double);
    cudaFuncSetAttribute(pilot_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    LDD_LAUNCH(pilot_median_kernel, dim3(nfields), dim3(PILOT_MED_THREADS), psmem, st, p, (const double*)offs, (const int*)cnt);
    return launch_status(h, "pilot kernels");
}

// ---------------------------------------------------------------------------------------------
// Host: the field-to-field walk of Framer.readfield (lddecode_core.py:1194-1223) over a capture
// whose planes were demodulated on ONE global block grid (plane index k <-> capture sample
// k + blockcut).  Each window is what rf.demod(infile, readsample, readlen) would have covered;
// its peak list is cut out of the global chase when the window starts on a peak of that chase
// (the normal case: nextfieldoffset is a peak), otherwise the callback runs the chase for that
// window on the device.
namespace {

// Peaks of the chase started at window_start, taken from the global chase.  Returns false when the
// window does not start on a global peak.
bool window_from_global(const long long* gp, int ngp, long long b, long long wl, int L, int* k0_out, int* k1_out) {
    const long long half = L / 2, skip = (long long)(L * .4);
    const long long limit = b + wl - 2LL * L;
    int k0 = (int)(std::lower_bound(gp, gp + ngp, b) - gp);
    if (k0 >= ngp || gp[k0] != b) return false;
    // a peak belongs to the window's list iff the step that found it started below `limit`:
    // steps after peak p start at p + skip, then advance by `half` until the next peak is inside
    int k = k0 + 1;
    for (; k < ngp; ++k) {
        long long i0 = gp[k - 1] + skip;
        long long m = (gp[k] - i0) / half;
        long long istep = i0 + m * half;
        if (istep >= limit) break;
    }
    *k0_out = k0;
    *k1_out = k;
    return true;
}

// A window that starts OFF the global chain (the second read of a capture starts `blockcut` samples before a peak,
// lddecode_core.py:374-379).  Its own chase merges into the global one at its first peak; until then every step looks
// at a half-line window, and what the global chase saw there is known from the global peak list alone: between two
// global peaks p[k-1], p[k] the global chase stepped from s = p[k-1] + skip in half-line windows, all of them empty
// (maximum <= 0.2) except the last one, G, whose first arg-max is p[k].  A local window that lies inside the empty
// windows is empty; one that lies inside the empty windows plus G and contains p[k] finds p[k] (everything else in it
// is <= 0.2 < val(p[k]), or inside G and therefore no larger, strictly smaller if earlier).  Anything else (samples the
// global chase skipped, or part of G without its peak) cannot be decided here: return false and let the caller chase
// on the samples.  gstart / gend: first sample and loop bound of the global chase (plane coordinates).
bool window_offpeak_from_global(const long long* gp, int ngp, long long gstart, long long gend, long long b, long long wl, int L,
                                int* k0_out, int* k1_out) {
    const long long half = L / 2, skip = (long long)(L * .4);
    const long long limit = b + wl - 2LL * L;
    long long i = b;
    if (i < gstart) return false;
    int kmerge = -1;
    for (int guard = 0; guard < 64 && i < limit; ++guard) {
        // segment k: the global steps that follow peak k-1 (k = 0: the start of the global chase)
        int k = (int)(std::upper_bound(gp, gp + ngp, i - skip) - gp);      // number of peaks with p + skip <= i
        const long long s = k == 0 ? gstart : gp[k - 1] + skip;
        if (i < s) return false;
        const long long wend = i + half;
        if (k == ngp) {
            // after the last global peak: empty windows up to the end of the global chase's last step
            if (gend <= s) return false;
            const long long nsteps = (gend - s + half - 1) / half;         // steps at s + t*half < gend
            if (wend > s + nsteps * half) return false;
            i += half;
            continue;
        }
        const long long T = (gp[k] - s) / half;
        const long long g0 = s + T * half, g1 = g0 + half;                  // G = [g0, g1)
        if (i >= g1) return false;                                           // inside the stretch the global chase skipped
        if (wend <= g0) { i += half; continue; }                            // empty
        if (wend <= g1 && gp[k] >= i && gp[k] < wend) { kmerge = k; break; }
        return false;
    }
    if (kmerge < 0) {
        if (i >= limit) { *k0_out = 0; *k1_out = 0; return true; }          // no peak at all before the loop bound
        return false;
    }
    int k = kmerge + 1;
    for (; k < ngp; ++k) {
        long long i0 = gp[k - 1] + skip;
        long long m = (gp[k] - i0) / half;
        if (i0 + m * half >= limit) break;
    }
    *k0_out = kmerge;
    *k1_out = k;
    return true;
}

}  // namespace

extern "C" int ldd_window_peaks_from_global(const long long* gpeaks, int ngpeaks, long long gstart, long long gend,
                                            long long b, long long wl, int linelen, int* k0, int* k1) {
    if (!gpeaks || !k0 || !k1 || ngpeaks < 0 || linelen < 4) return 0;
    if (b == gstart) return 0;                       // the chase itself started here: the caller has the list already
    if (window_from_global(gpeaks, ngpeaks, b, wl, linelen, k0, k1)) return 1;
    return window_offpeak_from_global(gpeaks, ngpeaks, gstart, gend, b, wl, linelen, k0, k1) ? 1 : 0;
}

extern "C" int ldd_field_chain(ldd_handle* h, const long long* gpeaks, const double* gvals, int ngpeaks,
                               long long plane_len, long long plane_origin, long long ncap, long long readlen,
                               long long first_readsample, long long stop_readsample, int tolerant,
                               int max_fields, ldd_window_peaks_fn cb, void* ctx,
                               ldd_field* fields, long long* base, long long* winlen, long long* readsample_out,
                               double* linelocs1, unsigned char* linebad, int ll_stride, int* nfields_out) {
    if (!h || !gpeaks || !gvals || !fields || !base || !winlen || !readsample_out || !linelocs1 || !linebad || !nfields_out)
        return LDD_EINVAL;
    const ldd_config& c = h->cfg;
    const int L = c.linelen;
    long long readsample = first_readsample;
    int nf = 0;
    std::vector<long long> rel;
    while (nf < max_fields && readsample < stop_readsample) {
        ldd_range r;
        int rc = ldd_demod_range_query(h, readsample, readlen, &r);
        if (rc) return rc;
        if (r.last_needed > ncap) break;                               // rf.demod returns None -> readfield returns None
        // window output[j] <-> capture r.first_sample + blockcut + j; plane k <-> capture plane_origin + blockcut + k
        long long b = r.first_sample - plane_origin;
        long long wl = r.total_out;
        if (b < 0 || b + wl > plane_len) break;
        const long long* pk = nullptr;
        const double* vl = nullptr;
        int np = 0;
        int k0 = 0, k1 = 0;
        bool fast = (b == 0 && plane_origin == 0 && first_readsample == 0 && nf == 0) ? true : false;
        if (fast) {
            // the global chase itself started here
            long long limit = wl - 2LL * L, half = L / 2, skip = (long long)(L * .4);
            k0 = 0;
            int k = 0;
            long long iprev = 0;
            for (; k < ngpeaks; ++k) {
                long long i0 = k == 0 ? 0 : gpeaks[k - 1] + skip;
                long long m = (gpeaks[k] - i0) / half;
                if (i0 + m * half >= limit) break;
                iprev = i0;
            }
            (void)iprev;
            k1 = k;
        } else {
            fast = window_from_global(gpeaks, ngpeaks, b, wl, L, &k0, &k1);
            // (the global chase of the pipeline starts at plane sample 0 and runs while i < plane_len - 2 L)
            if (!fast && !getenv("LDD_NO_OFFPEAK_FAST"))
                fast = window_offpeak_from_global(gpeaks, ngpeaks, 0, plane_len - 2LL * L, b, wl, L, &k0, &k1);
        }
        if (fast) {
            rel.resize(k1 - k0);
            for (int k = k0; k < k1; ++k) rel[k - k0] = gpeaks[k] - b;
            pk = rel.data();
            vl = gvals + k0;
            np = k1 - k0;
        } else {
            if (!cb) return fail_msg(h, LDD_EINVAL, "window does not start on a peak and no callback given");
            rc = cb(ctx, b, wl, &pk, &vl, &np);
            if (rc) return rc;
        }
        ldd_field* f = &fields[nf];
        rc = ldd_field_locate(h, pk, vl, np, wl, 0, f, linelocs1 + (size_t)nf * ll_stride, linebad + (size_t)nf * ll_stride, ll_stride);
        if (rc) return rc;
        base[nf] = b;
        winlen[nf] = wl;
        readsample_out[nf] = readsample;
        ++nf;
        if (f->stage == LDD_FIELD_CRASH) {
            // the reference raises here (vsync inside the first 11 peaks).  A range that starts at an
            // arbitrary place (shard / chunk start) steps forward instead and tries again.
            if (!tolerant) break;
            readsample += 20LL * L;
            continue;
        }
        // Framer.readfield: where the next read starts (lddecode_core.py:1204-1212)
        long long next = readsample + f->nextfieldoffset;
        bool valid_so_far = f->stage == LDD_FIELD_LOCATED;
        if (!valid_so_far) {
            if (np < 100) next = readsample + (long long)(c.freq_hz * 10);
            else if (f->nvsyncs == 0) next = readsample + (long long)(c.freq_hz * 1);
        }
        if (next <= readsample) break;                                 // would loop forever (the reference does)
        readsample = next;
    }
    *nfields_out = nf;
    if (nf >= max_fields && readsample < stop_readsample) {
        // the table is full: an error only if the walk would have gone on (the next window is readable and on the planes)
        ldd_range r;
        if (ldd_demod_range_query(h, readsample, readlen, &r) == LDD_OK && r.last_needed <= ncap && r.first_sample - plane_origin >= 0 &&
            r.first_sample - plane_origin + r.total_out <= plane_len)
            return fail_msg(h, LDD_ECAP, "ldd_field_chain: max_fields reached before the end of the range");
    }
    return LDD_OK;
}

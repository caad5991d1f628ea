// Kernel (5): per-line time-base correction.  Restates Field.downscale + lddutils.scale +
// the final quantisation of FieldNTSC/FieldPAL.downscale (lddecode_core.py:789-812, 1023-1035,
// 1135-1159; lddutils.py:83-97): every output line is a cubic interpolating spline with
// not-a-knot ends (FITPACK splrep s=0,k=3) through the input samples int(b)..int(e) of the line,
// evaluated at outwidth equidistant points, times the line's wow factor, then scaled to uint16.
//
// The spline is solved in closed form instead of by a sequential tridiagonal sweep.  On unit
// spacing its second derivatives satisfy M[i-1] + 4 M[i] + M[i+1] = d[i], d[i] = 6 (y[i-1] - 2 y[i] + y[i+1]).
// The inverse of the infinite (1,4,1) operator is g[k] = c r^|k|, c = 1 / (2 sqrt 3), r = sqrt 3 - 2, so a
// particular solution is P = g * d = c (F + B - d) with the two one-pole recursions
// F[i] = d[i] + r F[i-1] (causal) and B[i] = d[i] + r B[i+1] (anti-causal), run over the line's samples and a
// halo of 32 real neighbours on each side (|r|^32 = 5e-19).  The recursions are evaluated chunk-parallel:
// every thread owns 16 consecutive samples, reduces them to the two chunk sums, and takes its entry states
// from the sums of its two neighbours on each side (|r|^16 = 7e-10 per chunk, so two neighbours are exact
// in float64).  That is ~8 float64 operations per input sample instead of a 51-tap FIR.  The two
// not-a-knot rows then fix the homogeneous part alpha r^i + beta r^(n-i) by a 2x2 solve.  One CTA per
// output line.
#include "ldd_internal.h"

namespace ldd {

constexpr int TBC_H = 32;                 // halo samples on each side of the line
constexpr int TBC_C = 16;                 // samples per thread chunk of the recursions
constexpr int TBC_MAXD = 4032;            // longest input line span supported
constexpr int TBC_THREADS = 256;
static_assert((TBC_MAXD + 2 * TBC_H - 1 + TBC_C - 1) / TBC_C <= TBC_THREADS, "one chunk per thread");
constexpr int TBC_NPOW = 64;              // r^64 = 4e-37: reach of the homogeneous correction

// r^i, i < TBC_NPOW: set once per process.
__constant__ double c_tbc_rpow[TBC_NPOW];

struct TbcParams {
    const float* plane;       // input plane
    long long n;              // its length
    double plane_add;         // plane value + plane_add = Hz (ire0 for demod/demod_05, 0 otherwise)
    const long long* base;    // [nfields] plane index of the field window's sample 0 (or NULL)
    const double* linelocs;   // [nfields][ll_stride] line positions relative to the window
    const int* linecount;     // [nfields]
    int ll_stride;
    int lineoffset;           // first line = linelocs[lineoffset]
    double lineloc_add;       // added to every line position (FieldNTSC.apply_offsets, lddecode_core.py:1161-1162)
    int outwidth;
    int wow;                  // multiply by (e-b)/linelen
    int linelen;
    int mode;                 // 0: float64 Hz, 1: uint16 TBC sample
    double ire0, hz_ire, vsync_ire, out_scale, out_off;
    void* out;                // [nfields][out_stride]
    long long out_stride;
    const long long* field_off; // [nfields] element offset of each field's line 0 in out (NULL: field * out_stride)
    long long line_stride;    // elements between consecutive lines of a field (outwidth, or 2 * outwidth for a frame)
    const float* burstlevel;  // NTSC final: [nfields][ll_stride] or NULL
    float clevel_k;           // float32(327.67 * clevel)
    int* status;              // [nfields]: OR of per-line error bits: 1 = a line was not resampled, plus 32 when the only
                              // reason was a span longer than this launch's maxd (<= TBC_MAXD: ldd_tbc_long_lines can still
                              // do it) or 64 for any other reason (window outside the plane, degenerate, > TBC_MAXD)
    int maxd;                 // longest input span this launch has shared memory for (<= TBC_MAXD)
    int min_dist;             // long-lines pass: only lines with a span above this are done (0: all)
};

// status bits of a line that cannot be resampled by this launch (see TbcParams::status)
__device__ inline int tbc_skip_bits(bool geom_ok, int dist) { return (geom_ok && dist <= TBC_MAXD) ? (1 | 32) : (1 | 64); }

// int -> double and double -> int without the conversion pipe (0 <= v < 2^31): 2^52 + v has v in its low mantissa bits
__device__ inline double tbc_i2d(int v) {
#ifdef LDD_EMU
    return (double)v;
#else
    return __hiloint2double(0x43300000, v) - 4503599627370496.0;
#endif
}
__device__ inline int tbc_floor_nonneg(double x) {        // (int)x for 0 <= x < 2^31
#ifdef LDD_EMU
    return (int)x;
#else
    int i = __double2loint(x + 4503599627370496.0);       // round to nearest
    return tbc_i2d(i) > x ? i - 1 : i;
#endif
}

__global__ void __launch_bounds__(TBC_THREADS, 4) tbc_kernel(const TbcParams p) {
    LDD_DYN_SMEM(smem_raw);
    // Both arrays are indexed by u = i + H through PX(u) = u + u/16: a thread owns 16 consecutive samples, and
    // the padding slot per 16 puts the chunks of neighbouring threads 17 doubles apart (no bank conflicts).
    const int maxq = (p.maxd + 1 + 2 * TBC_H + TBC_C - 1) / TBC_C;
    double* Ms = (double*)smem_raw;                       // M[-H .. dist+H]
    double* ys = Ms + ((TBC_C + 1) * maxq + 4);           // y[-H .. dist+H], zero padded to a whole chunk (float64: converted once)
    double* Lf = ys + ((TBC_C + 1) * maxq + 4);           // chunk sums of the causal recursion
    double* Lb = Lf + (maxq + 2);                         // ... of the anti-causal one
#define PX(u) ((u) + ((u) >> 4))

    const int tid = threadIdx.x;
    const int field = blockIdx.y, line = blockIdx.x;
    // all four table reads are issued together (line + lineoffset + 1 < ll_stride for every line of the grid)
    const int linecount = p.linecount[field];
    const double* ll = p.linelocs + (size_t)field * p.ll_stride;
    const double b = ll[p.lineoffset + line] + p.lineloc_add, e = ll[p.lineoffset + line + 1] + p.lineloc_add;
    const long long base = p.base ? p.base[field] : 0;
    if (line >= linecount) return;
    const long long ib = (long long)b, ie = (long long)e;
    const int dist = (int)(ie - ib);
    const int W = p.outwidth;
    char* outbase = (char*)p.out;
    if (p.min_dist > 0 && dist <= p.min_dist) return;    // long-lines pass: the first pass has dealt with this line
    const bool geom_ok = b >= 0.0 && dist >= 3 && base + ib + dist + 1 <= p.n && base + ib >= 0;
    if (!geom_ok || dist > p.maxd) {
        if (tid == 0) atomicOr(&p.status[field], tbc_skip_bits(geom_ok, dist));
        return;
    }
    const double r = -0.26794919243112270647;      // sqrt(3) - 2
    const double c = 0.28867513459481288225;       // 1 / (2 sqrt 3)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4, r16 = r8 * r8;   // r^16: entry-state weight of the second neighbour chunk
    // stage the samples (relative values; the spline is linear so plane_add is added at the end).
    // d[u] exists for u = 1 .. U-2; chunk q owns u = 1 + 16 q .. 16 + 16 q.
    const int U = dist + 1 + 2 * TBC_H;
    const int nq = (U - 2 + TBC_C - 1) / TBC_C;
    {
        const long long s0 = base + ib - TBC_H;
        const float* src = p.plane + s0;
        if (s0 >= 0 && s0 + U <= p.n) {
            // all loads of a batch are issued before the first conversion: one memory round trip per 8 samples of a thread
            constexpr int SB = 8;
            for (int i0 = tid; i0 < U; i0 += SB * TBC_THREADS) {
                float v[SB];
                LDD_UNROLL
                for (int k = 0; k < SB; ++k) {
                    const int i = i0 + k * TBC_THREADS;
                    v[k] = i < U ? src[i] : 0.f;
                }
                LDD_UNROLL
                for (int k = 0; k < SB; ++k) {
                    const int i = i0 + k * TBC_THREADS;
                    if (i < U) ys[PX(i)] = (double)v[k];
                }
            }
        } else {
            for (int i = tid; i < U; i += TBC_THREADS) {
                long long s = s0 + i;
                s = s < 0 ? 0 : (s >= p.n ? p.n - 1 : s);
                ys[PX(i)] = (double)p.plane[s];
            }
        }
        for (int i = U + tid; i < nq * TBC_C + 2; i += TBC_THREADS) ys[PX(i)] = 0.0;
    }
    __syncthreads();
    // Every thread owns at most one chunk (nq <= 256).  d beyond the staged samples is computed from the zero padding:
    // whatever the right-hand side is outside the line, g * d satisfies the interior equations exactly and differs
    // only by a homogeneous solution, which the not-a-knot solve below absorbs (and which has decayed by r^31 anyway).
    double d[TBC_C];
    const bool act = tid < nq;
    if (act) {
        const double* y = ys + tid * (TBC_C + 1);            // PX(16 q + k) = 17 q + k for k < 16; k = 16, 17 sit behind the padding slot
        double ym = y[0], y0 = y[1];
        LDD_UNROLL
        for (int k = 0; k < TBC_C; ++k) {
            const double yp = y[k + 2 < TBC_C ? k + 2 : k + 3];
            d[k] = 6.0 * (fma(-2.0, y0, ym) + yp);
            ym = y0;
            y0 = yp;
        }
        // chunk sums: Lf = sum r^(15-k) d[k], Lb = sum r^k d[k]
        double f = 0.0, bk = 0.0;
        LDD_UNROLL
        for (int k = 0; k < TBC_C; ++k) {
            f = fma(f, r, d[k]);
            bk = fma(bk, r, d[TBC_C - 1 - k]);
        }
        Lf[tid] = f;
        Lb[tid] = bk;
    }
    __syncthreads();
    if (act) {
        const int q = tid;
        double f = (q >= 1 ? Lf[q - 1] : 0.0) + (q >= 2 ? r16 * Lf[q - 2] : 0.0);
        double bk = (q + 1 < nq ? Lb[q + 1] : 0.0) + (q + 2 < nq ? r16 * Lb[q + 2] : 0.0);
        double* M = Ms + q * (TBC_C + 1);                   // M[k] <-> u = 1 + 16 q + k: PX = 17 q + 1 + k, behind the padding slot for k = 15
        // forward sweep parks F - d in the thread's own slots of Ms, the backward sweep completes them (registers: only d)
        LDD_UNROLL
        for (int k = 0; k < TBC_C; ++k) {
            f = fma(f, r, d[k]);
            M[k + 1 < TBC_C ? k + 1 : k + 2] = f - d[k];
        }
        LDD_UNROLL
        for (int k = TBC_C - 1; k >= 0; --k) {
            bk = fma(bk, r, d[k]);
            double* m = &M[k + 1 < TBC_C ? k + 1 : k + 2];
            *m = c * (*m + bk);
        }
    }
    __syncthreads();
#define M0(i) Ms[PX((i) + TBC_H)]
    {
        // not-a-knot rows -> homogeneous part (every thread solves the 2x2 system for itself)
        const double L = M0(0) - 2.0 * M0(1) + M0(2);
        const double R = M0(dist) - 2.0 * M0(dist - 1) + M0(dist - 2);
        const double A = (1.0 - r) * (1.0 - r);
        const double q = (dist - 2 < TBC_NPOW) ? c_tbc_rpow[dist - 2] : 0.0;
        const double den = A * (1.0 - q * q);
        const double alpha = (-L + q * R) / den, beta = (-R + q * L) / den;
        __syncthreads();                                  // all threads have read the uncorrected ends
        if (tid < TBC_NPOW) {
            const int i = tid;                            // front: both terms where they overlap
            if (i <= dist) {
                double corr = alpha * c_tbc_rpow[i];
                if (dist - i < TBC_NPOW) corr += beta * c_tbc_rpow[dist - i];
                M0(i) += corr;
            }
        } else if (tid < 2 * TBC_NPOW) {
            const int k = tid - TBC_NPOW, i = dist - k;   // back: indices the front threads do not own
            if (i >= TBC_NPOW) M0(i) += beta * c_tbc_rpow[k];
        }
    }
    __syncthreads();
    // evaluate (np.linspace(fb, fb + (e - b), W + 1)[:-1]: x_j = j * step + fb)
    const double fb = b - (double)ib;
    const double stop = (e - b) + fb;
    const double step = (stop - fb) / (double)W;
    const double wowf = p.wow ? (e - b) / (double)p.linelen : 1.0;
    const double sixth = 1.0 / 6.0;
    const double inv_hz_ire = 1.0 / p.hz_ire;
    for (int j = tid; j < W; j += TBC_THREADS) {
        double x = tbc_i2d(j) * step + fb;
        int i = tbc_floor_nonneg(x);
        if (i > dist - 1) i = dist - 1;
        double t = x - tbc_i2d(i), u = 1.0 - t;
        double Mi = M0(i), Mj = M0(i + 1);
        // S = Mi u^3/6 + Mj t^3/6 + (y_i - Mi/6) u + (y_j - Mj/6) t
        const double Mi6 = Mi * sixth, Mj6 = Mj * sixth;
        double S = u * fma(Mi6, fma(u, u, -1.0), ys[PX(i + TBC_H)]) + t * fma(Mj6, fma(t, t, -1.0), ys[PX(i + 1 + TBC_H)]);
        double hz = (S + p.plane_add) * wowf;
        size_t o = (p.field_off ? (size_t)p.field_off[field] : (size_t)field * (size_t)p.out_stride) + (size_t)line * (size_t)p.line_stride + j;
        if (p.mode == 0) {
            ((double*)outbase)[o] = hz;
        } else {
            double red = (hz - p.ire0) * inv_hz_ire;         // the reference divides; the product differs by at most one ulp
            red -= p.vsync_ire;
            double v = red * p.out_scale + p.out_off;
            v = v < 0.0 ? 0.0 : (v > 65535.0 ? 65535.0 : v);
            unsigned short q16 = (unsigned short)tbc_floor_nonneg(v + 0.5);
            if (p.burstlevel && line >= 1 && line < linecount - 1 && j < 2) {
                // burst polarity / level markers (lddecode_core.py:1144-1154)
                float bl = p.burstlevel[(size_t)field * p.ll_stride + line];
                if (j == 0) q16 = bl > 0.f ? 16384 : 32768;
                else q16 = (unsigned short)(p.clevel_k * fabsf(bl));
            }
            ((unsigned short*)outbase)[o] = q16;
        }
    }
}

#undef PX
#undef M0

// ---------------------------------------------------------------------------------------------
// float32 variant for uint16 output (the float32 / mixed lanes): the same closed-form spline, with
//   * the line's samples brought in by ONE bulk asynchronous copy (cp.async.bulk + mbarrier, the 1-D form of TMA)
//     straight into an unpadded shared array -- persistent CTAs walk over the lines and the copy of the next line
//     runs under the arithmetic of the current one (double-buffered staging);
//   * 17 samples per thread chunk instead of 16: with an odd chunk length consecutive threads' chunks start 17 banks
//     apart, so the unpadded array the bulk copy needs is conflict-free without a padding slot;
//   * float32 recursions and evaluation (the plane is float32 already; a uint16 step is 21-34 Hz, the float32 spline
//     is good to a fraction of a Hz), float64 only for the sample positions (a float32 position at 2300 samples would
//     be 2e-4 samples off: 10 Hz on a steep edge).
// Output within +-1 LSB of the float64 kernel (a sample moves only when it sits within ~1e-2 LSB of a rounding
// boundary).  The float64 kernel above stays the one for the exact lane and for float64 (Hz) output.
constexpr int TBF_C = 17;
constexpr int TBF_THREADS = 256;
constexpr int TBF_NPOW = 24;                            // reach of the homogeneous (not-a-knot) correction in float32
constexpr int TBF_MAXU = TBF_C * TBF_THREADS;            // 4352 staged samples at most

struct TbfGeom {            // one work item (field, line), filled by thread 0 when it issues the item's copy
    double b, e;
    long long src0;         // plane index of staged sample 0 (16-byte aligned element index when bulk)
    int field, line, dist, lead;   // lead = TBC_H + alignment shift: staged index of line sample 0
    int lc;                 // the field's line count (burst markers)
    int U;                  // staged samples
    int state;              // 0: nothing to do (line >= linecount), 1: bulk copy in flight, 2: load by hand, 3: not resampled (status bits in U)
};


__global__ void __launch_bounds__(TBF_THREADS, 5) tbc_f32_kernel(const TbcParams p, int nitems, int lines_per_field, int maxu) {
    LDD_DYN_SMEM(smem_raw);
    float* ysb[2];
    ysb[0] = (float*)smem_raw;
    ysb[1] = ysb[0] + maxu;
    float* Ms = ysb[1] + maxu;
    float* Lf = Ms + maxu + 32;                      // the last chunk's slots may reach past the staged samples
    float* Lb = Lf + TBF_THREADS;
    __shared__ TbfGeom geom[2];
    __shared__ unsigned long long mbar[2];           // 8-byte aligned by type
    __shared__ float s_rpow[TBF_NPOW + 2];           // r^i: |r|^24 = 2e-14, far below float32 resolution
    if (threadIdx.x < TBF_NPOW + 2) s_rpow[threadIdx.x] = (float)c_tbc_rpow[threadIdx.x];
    const int tid = threadIdx.x;
    const float r = -0.26794919243112270647f, c = 0.28867513459481288225f;

    // Thread 0 runs the copies.  The table values of an item (two line positions, window base, line count) are
    // fetched one iteration before they are needed, so issuing a copy never waits for global memory.
    struct Pre { double b, e; long long base; int lc, field, line; };
    auto prefetch = [&](int w) {
        Pre q;
        q.lc = -1; q.b = q.e = 0.0; q.base = 0; q.field = q.line = 0;
        if (w < nitems) {
            q.field = w / lines_per_field;
            q.line = w - q.field * lines_per_field;
            const double* ll = p.linelocs + (size_t)q.field * p.ll_stride;
            q.lc = p.linecount[q.field];
            q.b = ll[p.lineoffset + q.line];
            q.e = ll[p.lineoffset + q.line + 1];
            q.base = p.base ? p.base[q.field] : 0;
        }
        return q;
    };
    auto issue = [&](const Pre& q, int buf) {
        TbfGeom g;
        g.state = 0; g.field = q.field; g.line = q.line; g.lc = q.lc;
        if (q.lc >= 0 && q.line < q.lc) {
            g.b = q.b + p.lineloc_add;
            g.e = q.e + p.lineloc_add;
            const long long ib = (long long)g.b, ie = (long long)g.e;
            g.dist = (int)(ie - ib);
            const long long s0 = q.base + ib - TBC_H;
            const bool geom_ok = g.b >= 0.0 && g.dist >= 3 && q.base + ib + g.dist + 1 <= p.n && q.base + ib >= 0;
            if (!geom_ok || g.dist > p.maxd) {
                g.state = 3;
                g.U = tbc_skip_bits(geom_ok, g.dist);        // the status bits ride in U (no samples are staged)
            } else {
                const int shift = (int)(s0 & 3);
                const long long a0 = s0 - shift;
                const int U = (g.dist + 1 + 2 * TBC_H + shift + 3) & ~3;
                g.src0 = s0; g.lead = TBC_H; g.U = g.dist + 1 + 2 * TBC_H; g.state = 2;       // by hand: window at a plane edge
#ifndef LDD_EMU
                if (a0 >= 0 && a0 + U <= p.n && U <= maxu && (((uintptr_t)p.plane) & 15) == 0) {
                    g.src0 = a0; g.lead = TBC_H + shift; g.U = U; g.state = 1;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of this buffer before the async write
                    mbar_expect_tx(&mbar[buf], (unsigned)U * 4u);
                    bulk_g2s(ysb[buf], p.plane + a0, (unsigned)U * 4u, &mbar[buf]);
                }
#endif
            }
        }
        geom[buf] = g;
    };

    if (tid == 0) {
#ifndef LDD_EMU
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();
    Pre pre;
    if (tid == 0) {
        issue(prefetch(blockIdx.x), 0);
        pre = prefetch(blockIdx.x + gridDim.x);
    }
    __syncthreads();
    unsigned phase[2] = {0u, 0u};
    int cur = 0;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x, cur ^= 1) {
        // the next item's copy runs under this item's arithmetic (its buffer was released by the barrier that ended
        // the previous iteration)
        if (tid == 0) {
            issue(pre, cur ^ 1);
            pre = prefetch(w + 2 * gridDim.x);
        }
        const TbfGeom g = geom[cur];
        float* ys = ysb[cur];
        if (g.state == 0 || g.state == 3) {
            if (g.state == 3 && tid == 0) atomicOr(&p.status[g.field], g.U);
            __syncthreads();
            continue;
        }
        if (g.state == 1) {
#ifndef LDD_EMU
            mbar_wait(&mbar[cur], phase[cur]);
            phase[cur] ^= 1u;
#endif
        } else {
            for (int i = tid; i < g.U; i += TBF_THREADS) {
                long long s = g.src0 + i;
                s = s < 0 ? 0 : (s >= p.n ? p.n - 1 : s);
                ys[i] = p.plane[s];
            }
            __syncthreads();
        }
        const int U = g.U, dist = g.dist;
        // d[u] for u = 1 .. U-2; chunk q owns u = 1 + 17 q .. 17 + 17 q.  Beyond the staged samples the right-hand side
        // is taken as zero: that only changes the homogeneous part, which the not-a-knot solve absorbs.
        const int nq = (U - 2 + TBF_C - 1) / TBF_C;
        float d[TBF_C];
        const bool act = tid < nq;
        if (act) {
            const int u0 = 1 + TBF_C * tid;
            float ym = ys[u0 - 1], y0 = ys[u0];
            LDD_UNROLL
            for (int k = 0; k < TBF_C; ++k) {
                const int u = u0 + k;
                const float yp = (u + 1 < U) ? ys[u + 1] : 0.f;
                d[k] = (u < U - 1) ? 6.f * (fmaf(-2.f, y0, ym) + yp) : 0.f;
                ym = y0;
                y0 = yp;
            }
            float f = 0.f, bk = 0.f;
            LDD_UNROLL
            for (int k = 0; k < TBF_C; ++k) {
                f = fmaf(f, r, d[k]);
                bk = fmaf(bk, r, d[TBF_C - 1 - k]);
            }
            Lf[tid] = f;
            Lb[tid] = bk;
        }
        __syncthreads();
        if (act) {
            // r^17 = 2e-10: one neighbour chunk is exact in float32
            float f = tid >= 1 ? Lf[tid - 1] : 0.f;
            float bk = tid + 1 < nq ? Lb[tid + 1] : 0.f;
            float* M = Ms + 1 + TBF_C * tid;
            LDD_UNROLL
            for (int k = 0; k < TBF_C; ++k) {
                f = fmaf(f, r, d[k]);
                M[k] = f - d[k];
            }
            LDD_UNROLL
            for (int k = TBF_C - 1; k >= 0; --k) {
                bk = fmaf(bk, r, d[k]);
                M[k] = c * (M[k] + bk);
            }
        }
        __syncthreads();
        const int lead = g.lead;
        // not-a-knot rows -> homogeneous part alpha r^i + beta r^(dist-i): every thread solves the 2x2 system for itself
        // and adds the correction to the (few) values it reads near the line's ends -- nothing is written back, so the
        // recursions' result needs no further barrier
        float alpha, beta;
        {
            const float* Mz = Ms + lead;
            const float L = Mz[0] - 2.f * Mz[1] + Mz[2];
            const float R = Mz[dist] - 2.f * Mz[dist - 1] + Mz[dist - 2];
            const float A = (1.f - r) * (1.f - r);
            const float q = (dist - 2 < TBF_NPOW) ? s_rpow[dist - 2] : 0.f;
            const float den = A * (1.f - q * q);
            alpha = (-L + q * R) / den;
            beta = (-R + q * L) / den;
        }
        // evaluate: polynomial in float32, the affine map to the uint16 scale folded into one FMA.  Positions
        // x_j = fb + j step (np.linspace) need ~1e-5 samples at x ~ 2300, which a float32 x does not have; split instead:
        // step = si + dd with si its nearest integer (2 for a nominal line), x_j = j si + (fb + j dd), and the bracket -- a
        // few samples at most for any line the kernel accepts -- is a float32 FMA of two-part constants: no float64 and
        // no 64-bit conversion per output sample.  Where that rounds across an integer the neighbouring interval is
        // evaluated at t = 1 - eps instead of t = eps, which the spline's continuity makes the same value.
        const double ibd = (double)(long long)g.b;
        const double fb = g.b - ibd;
        const double stop = (g.e - g.b) + fb;
        const int W = p.outwidth;
        const double step = (stop - fb) / (double)W;
        const int si = (int)(step + 0.5);
        const double dd = step - (double)si;
        const float d_hi = (float)dd, d_lo = (float)(dd - (double)d_hi);
        const float fb_hi = (float)fb, fb_lo = (float)(fb - (double)fb_hi);
        const double wowf = p.wow ? (g.e - g.b) / (double)p.linelen : 1.0;
        // v = ((S + add) wow - ire0) / hz_ire - vsync_ire) * out_scale + out_off  =  S * ka + kb
        const double k1 = p.out_scale / p.hz_ire;
        const float ka = (float)(wowf * k1);
        const float kb = (float)((p.plane_add * wowf - p.ire0) * k1 - p.vsync_ire * p.out_scale + p.out_off);
        const float sixth = 1.f / 6.f;
        const int field = g.field, line = g.line;
        const int linecount = g.lc;
        unsigned short* outl = (unsigned short*)p.out + (p.field_off ? (size_t)p.field_off[field] : (size_t)field * (size_t)p.out_stride) +
                               (size_t)line * (size_t)p.line_stride;
        const float* Mz = Ms + lead;
        const float* yz = ys + lead;
        for (int j = tid; j < W; j += TBF_THREADS) {
            const float jf = (float)j;
            const float gx = fmaf(jf, d_hi, fb_hi) + fmaf(jf, d_lo, fb_lo);
            const float fl = floorf(gx);
            float t = gx - fl;
            int i = j * si + (int)fl;
            if (i > dist - 1) { t += (float)(i - (dist - 1)); i = dist - 1; }      // np.linspace's last points: beyond the last interval's start
            if (i < 0) { t += (float)i; i = 0; }
            const float u = 1.f - t;
            float Mi = Mz[i], Mj = Mz[i + 1];
            if (i < TBF_NPOW) { Mi = fmaf(alpha, s_rpow[i], Mi); Mj = fmaf(alpha, s_rpow[i + 1], Mj); }
            if (dist - i <= TBF_NPOW) { Mi = fmaf(beta, s_rpow[dist - i], Mi); Mj = fmaf(beta, s_rpow[dist - i - 1], Mj); }
            const float Mi6 = Mi * sixth, Mj6 = Mj * sixth;
            const float S = u * fmaf(Mi6, fmaf(u, u, -1.f), yz[i]) + t * fmaf(Mj6, fmaf(t, t, -1.f), yz[i + 1]);
            float v = fmaf(S, ka, kb);
            v = v < 0.f ? 0.f : (v > 65535.f ? 65535.f : v);
            unsigned short q16 = (unsigned short)(int)(v + 0.5f);
            if (p.burstlevel && line >= 1 && line < linecount - 1 && j < 2) {
                float bl = p.burstlevel[(size_t)field * p.ll_stride + line];
                if (j == 0) q16 = bl > 0.f ? 16384 : 32768;
                else q16 = (unsigned short)(p.clevel_k * fabsf(bl));
            }
            outl[j] = q16;
        }
        __syncthreads();        // everybody is done with ys[cur] / Ms before the next iteration's copy and recursions
    }
}

}  // namespace ldd

using namespace ldd;

static int tbc_launch(ldd_handle* h, const float* plane_dev, long long n, double plane_add,
                      const long long* base_dev, const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                      int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                      void* out_dev, long long out_stride, const long long* out_off_dev, long long line_stride,
                      const float* burstlevel_dev, double colorlevel, int* status_dev, void* stream, bool long_only) {
    if (!h || !plane_dev || !linelocs_dev || !linecount_dev || !out_dev || !status_dev) return LDD_EINVAL;
    if (nfields <= 0 || max_linecount <= 0) return LDD_OK;
    if (outwidth < 1 || (mode != 0 && mode != 1)) return LDD_EINVAL;
    if (lineoffset < 0 || lineoffset + max_linecount + 1 > ll_stride) return LDD_EINVAL;      // every CTA of the grid reads its two line positions
    const ldd_config& c = h->cfg;
    TbcParams p;
    p.plane = plane_dev; p.n = n; p.plane_add = plane_add;
    p.base = base_dev;
    p.linelocs = linelocs_dev; p.linecount = linecount_dev; p.ll_stride = ll_stride;
    p.lineoffset = lineoffset; p.lineloc_add = lineloc_add; p.outwidth = outwidth; p.wow = wow; p.linelen = c.linelen; p.mode = mode;
    p.ire0 = c.ire0; p.hz_ire = c.hz_ire; p.vsync_ire = c.vsync_ire;
    if (c.system == LDD_SYSTEM_NTSC) {            // lddecode_core.py:1141-1142
        p.out_scale = (double)(0xc800 - 0x0400) / (100.0 - c.vsync_ire);
        p.out_off = 1024.0;
    } else {                                      // lddecode_core.py:1029-1030
        p.out_scale = (double)(0xd300 - 0x0100) / (100.0 - c.vsync_ire);
        p.out_off = 256.0;
    }
    p.out = out_dev; p.out_stride = out_stride;
    p.field_off = out_off_dev; p.line_stride = line_stride > 0 ? line_stride : outwidth;
    p.burstlevel = burstlevel_dev;
    // np.uint16(327.67 * clevel * np.abs(float32 burstlevel)): python-float product, then float32 arithmetic
    double clevel = (1.0 / colorlevel) / (1700000.0 / 140.0);
    p.clevel_k = (float)(327.67 * clevel);
    p.status = status_dev;
    // shared memory for lines up to 25 % longer than nominal (longer ones are flagged in status like lines beyond
    // TBC_MAXD): 50 KB per CTA for PAL at 8fsc, so four CTAs share an SM
    p.maxd = c.linelen + c.linelen / 4 + 64;
    if (p.maxd > TBC_MAXD) p.maxd = TBC_MAXD;
    p.min_dist = 0;
    const bool f32_lane = mode == 1 && c.precision != LDD_PREC_F64 && !getenv("LDD_TBC_F64");
    int maxu = (p.maxd + 1 + 2 * TBC_H + 3 + 3) & ~3;
    if (f32_lane && maxu > TBF_MAXU) { maxu = TBF_MAXU; p.maxd = TBF_MAXU - 2 * TBC_H - 8; }
    if (long_only) {
        // second pass (ldd_tbc_long_lines): the exact kernel with shared memory for the longest span it supports, over the
        // lines the first pass left out because they were longer than ITS shared memory
        p.min_dist = p.maxd;
        p.maxd = TBC_MAXD;
    }
    const int maxq = (p.maxd + 1 + 2 * TBC_H + TBC_C - 1) / TBC_C;
    size_t smem = (size_t)(2 * ((TBC_C + 1) * maxq + 4) + 2 * (maxq + 2)) * sizeof(double);
    if (!h->tbc_taps_set) {
        const double r = -0.26794919243112270647;
        double rp[TBC_NPOW];
        for (int i = 0; i < TBC_NPOW; ++i) rp[i] = pow(r, (double)i);
        cudaMemcpyToSymbol(c_tbc_rpow, rp, sizeof rp);
        cudaStreamSynchronize((cudaStream_t)0);      // staged from pageable memory; the kernels run on non-blocking streams
        h->tbc_taps_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (f32_lane && !long_only) {
        // float32 lanes: bulk-copy staged, persistent CTAs
        const size_t smem32 = (size_t)(3 * maxu + 32 + 2 * TBF_THREADS) * sizeof(float);
        cudaFuncSetAttribute(tbc_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
        cudaFuncSetAttribute(tbc_f32_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        const int nitems = nfields * max_linecount;
        int per_sm = (int)((h->smem_optin + 1024) / (smem32 + 1024));
        if (per_sm > 5) per_sm = 5;                  // 48 registers x 256 threads: five CTAs per SM
        if (per_sm < 1) per_sm = 1;
        int grid = h->sm_count * per_sm;
        if (grid > nitems) grid = nitems;
        LDD_LAUNCH(tbc_f32_kernel, dim3(grid), dim3(TBF_THREADS), smem32, st, p, nitems, max_linecount, maxu);
        return launch_status(h, "tbc_f32_kernel");
    }
    cudaFuncSetAttribute(tbc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tbc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);      // several CTAs of ~50 KB per SM
    LDD_LAUNCH(tbc_kernel, dim3(max_linecount, nfields), dim3(TBC_THREADS), smem, st, p);
    return launch_status(h, "tbc_kernel");
}

extern "C" int ldd_tbc_fields_ex(ldd_handle* h, const float* plane_dev, long long n, double plane_add,
                                 const long long* base_dev, const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                                 int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                                 void* out_dev, long long out_stride, const long long* out_off_dev, long long line_stride,
                                 const float* burstlevel_dev, double colorlevel, int* status_dev, void* stream) {
    return tbc_launch(h, plane_dev, n, plane_add, base_dev, linelocs_dev, ll_stride, linecount_dev, nfields, max_linecount, lineoffset,
                      lineloc_add, outwidth, wow, mode, out_dev, out_stride, out_off_dev, line_stride, burstlevel_dev, colorlevel,
                      status_dev, stream, false);
}

extern "C" int ldd_tbc_long_lines(ldd_handle* h, const float* plane_dev, long long n, double plane_add,
                                  const long long* base_dev, const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                                  int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                                  void* out_dev, long long out_stride, const long long* out_off_dev, long long line_stride,
                                  const float* burstlevel_dev, double colorlevel, int* status_dev, void* stream) {
    return tbc_launch(h, plane_dev, n, plane_add, base_dev, linelocs_dev, ll_stride, linecount_dev, nfields, max_linecount, lineoffset,
                      lineloc_add, outwidth, wow, mode, out_dev, out_stride, out_off_dev, line_stride, burstlevel_dev, colorlevel,
                      status_dev, stream, true);
}

extern "C" int ldd_tbc_fields(ldd_handle* h, const float* plane_dev, long long n, double plane_add,
                              const long long* base_dev, const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                              int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                              void* out_dev, long long out_stride, const float* burstlevel_dev, double colorlevel,
                              int* status_dev, void* stream) {
    return ldd_tbc_fields_ex(h, plane_dev, n, plane_add, base_dev, linelocs_dev, ll_stride, linecount_dev, nfields, max_linecount,
                             lineoffset, lineloc_add, outwidth, wow, mode, out_dev, out_stride, nullptr, 0, burstlevel_dev, colorlevel,
                             status_dev, stream);
}

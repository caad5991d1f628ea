// Fused block demodulation: unpack -> FFT -> RF filter -> inverse FFT -> FM discriminator ->
// FFT -> post filters -> inverse FFTs -> sync threshold + circular 1-pole scan -> stitched stores.
// One CTA carries one block of N samples through the whole chain; nothing but the raw samples
// is read from HBM and nothing but the kept region of the output planes is written.
//
// Restates RFDecode.demodblock (lddecode_core.py:288-330) and the overlap-save stitching of
// RFDecode.demod (:385-422) with these algebraic changes (all exact up to rounding, checked
// against the oracle in tests/):
//   * real-input transforms run as length-M = N/2 complex transforms + an untangle step;
//   * ifft(X * RFVideo) is split into its even and odd output samples (decimation in frequency)
//     so that it is two length-M transforms as well;
//   * the demodulated signal has ire0 subtracted before its forward transform and the filters'
//     DC gain times ire0 added back at the store (keeps float32 planes at sub-Hz resolution);
//   * np.roll(., -F05_offset) is a phase ramp folded into the FVideo05 table;
//   * ifft(fft(sync) * FPsync) is evaluated as the circular first-order recursion it equals,
//     by a three-phase scan (thread-local, warp-shuffle, cross-warp) in float64.
#include "ldd_internal.h"

namespace ldd {

// The planes are written once and never re-read by this kernel: store them with the streaming
// (evict-first) hint so that ~0.9 GB of output per second of video does not push the L2-resident
// scratch slices out to DRAM.
#ifdef LDD_EMU
template <class U> __device__ inline void st_stream(U* p, U v) { *p = v; }
#else
template <class U> __device__ inline void st_stream(U* p, U v) { __stcs(p, v); }
#endif

// Phase timing (profiling builds only: -DLDD_PHASE_TIMING; tools/gpu_demod_phases.py): thread 0 of every CTA adds the
// clock cycles between consecutive marks to g_phase[lane offset + phase].
#ifdef LDD_PHASE_TIMING
__device__ unsigned long long g_phase[64];
#define PHASE_BEGIN() long long ph_last = clock64()
#define PHASE(i) do { if (threadIdx.x == 0) { long long ph_t = clock64(); atomicAdd(&g_phase[(sizeof(T) == 8 ? 32 : 0) + (i)], (unsigned long long)(ph_t - ph_last)); ph_last = ph_t; } } while (0)
#else
#define PHASE_BEGIN()
#define PHASE(i)
#endif

template <class T> struct Math;
template <> struct Math<double> {
    static __device__ inline double atan2(double y, double x) { return ::atan2(y, x); }
};
// float32 atan2 for the FM discriminator: octant reduction to t = min/max in [0, 1], atan(t) = t + t^3 q(t^2)
// with a degree-7 minimax q, quadrant fix-ups with two-part constants.  Maximum error 3.3e-7 rad (the same as a
// 2-ulp library atan2f near pi) in about half the instructions; atan2(0, 0) = 0 like np.angle.
__device__ inline float atan2_f32(float y, float x) {
#ifdef LDD_EMU
    return ::atan2f(y, x);
#else
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mx));
    float t = mn * r;
    t = fmaf(fmaf(-mx, t, mn), r, t);              // one Newton step on the quotient
    t = mx > 0.f ? t : 0.f;
    const float s = t * t;
    float q = 3.866738873e-03f;
    q = fmaf(q, s, -2.002674714e-02f);
    q = fmaf(q, s, 4.891432077e-02f);
    q = fmaf(q, s, -8.009681851e-02f);
    q = fmaf(q, s, 1.086575910e-01f);
    q = fmaf(q, s, -1.425704509e-01f);
    q = fmaf(q, s, 1.999868155e-01f);
    q = fmaf(q, s, -3.333332241e-01f);
    float a = fmaf(t * s, q, t);
    if (ay > ax) a = (1.57079637e+00f - a) + -4.37113883e-08f;
    if (x < 0.f) a = (3.14159274e+00f - a) + -8.74227766e-08f;
    return copysignf(a, y);
#endif
}
template <> struct Math<float> {
    static __device__ inline float atan2(float y, float x) { return atan2_f32(y, x); }
};

// ---- raw sample fetch with the unpackers fused in (ddunpack.c / lddutils.py:131-229) -----------
__device__ inline int fetch_sample(const void* rf, int fmt, long long s) {
    switch (fmt) {
        case LDD_FMT_U8: return (int)((const unsigned char*)rf)[s];
        case LDD_FMT_S16: return (int)((const short*)rf)[s];
        case LDD_FMT_U16: return (int)((const unsigned short*)rf)[s];
        case LDD_FMT_R30: {
            long long w = s / 3;
            int f = (int)(s - w * 3);
            unsigned v = ((const unsigned*)rf)[w];
            return (int)((v >> (10 * f)) & 0x3ffu);
        }
        default: {   // LDD_FMT_LDS40
            long long g = s >> 2;
            int f = (int)(s & 3);
            const unsigned char* b = (const unsigned char*)rf + g * 5;
            unsigned hi = b[f], lo = b[f + 1];
            // s0 = b0<<2 | b1>>6 ; s1 = (b1&0x3f)<<4 | b2>>4 ; s2 = (b2&0xf)<<6 | b3>>2 ; s3 = (b3&3)<<8 | b4
            return (int)(((hi << (2 + 2 * f)) | (lo >> (6 - 2 * f))) & 0x3ffu);
        }
    }
}

// ---- untangle: length-M FFT of z[n] = x[2n] + j x[2n+1]  ->  X[k], k = 0..M (X[M] packed in X[0].y)
template <class T, bool PAD>
__device__ inline void untangle(Cx<T>* Z, int M, const Cx<T>* __restrict__ WN, int tid, int nthr) {
    const T half = (T)0.5;
    for (int k = tid; k <= M / 2; k += nthr) {
        if (k == 0) {
            Cx<T> z = Z[0];
            Z[0] = mk<T>(z.x + z.y, z.x - z.y);
        } else {
            const int ik = pidx<PAD>(k), im = pidx<PAD>(M - k);
            Cx<T> a = Z[ik], b = conj(Z[im]);
            Cx<T> E = scale(a + b, half);
            Cx<T> Od = scale(mul_mj(a - b), half);
            Cx<T> Tw = WN[k] * Od;
            Z[ik] = E + Tw;
            Z[im] = conj(E - Tw);
        }
    }
}

// ---- tangle: conj-symmetric spectrum Y = D * F (k = 0..M)  ->  conj(Q), whose forward length-M
// FFT r gives the real signal: y[2n] = r.x, y[2n+1] = -r.y (F carries the 1/M).
template <class T, bool PAD>
__device__ inline void tangle(const Cx<T>* D, Cx<T>* Q, const Cx<T>* __restrict__ F, int M,
                              const Cx<T>* __restrict__ WN, int tid, int nthr) {
    const T half = (T)0.5;
    for (int k = tid; k <= M / 2; k += nthr) {
        if (k == 0) {
            Cx<T> d = D[0];
            T y0 = d.x * F[0].x, ym = d.y * F[M].x;
            Q[0] = mk<T>((y0 + ym) * half, -(y0 - ym) * half);
        } else {
            const int ik = pidx<PAD>(k), im = pidx<PAD>(M - k);
            Cx<T> a = D[ik] * F[k], b = conj(D[im] * F[M - k]);
            Cx<T> E = scale(a + b, half);
            Cx<T> Od = mulc(scale(a - b, half), WN[k]);      // * W_N^{-k}
            Cx<T> q = E + mul_pj(Od);
            Cx<T> qm = conj(E) + mul_pj(conj(Od));
            Q[ik] = conj(q);
            Q[im] = conj(qm);
        }
    }
}

// Compile-time variants (M, NT constants): the table values of a batch of iterations are fetched before
// the batch's arithmetic, so that a thread has several L2 round trips in flight instead of one per iteration.
// wb != nullptr: *wb = W_N^tid, the twiddles W_N^(tid + j NT) are products with constants instead of table loads
// (M / 2 / NT == 8 only: the float32 lane of the default block length).
template <class T, bool PAD, int M, int NT>
__device__ inline void untangle_static(Cx<T>* Z, const Cx<T>* __restrict__ WN, int tid, const Cx<T>* wb = nullptr) {
    constexpr int IT = (M / 2) / NT, B = sizeof(T) == 4 ? IT : (IT % 4 == 0 ? 4 : 1);
    const T half = (T)0.5;
    LDD_UNROLL
    for (int it0 = 0; it0 < IT; it0 += B) {
        // table values and the shared-memory operands of the whole batch first (independent loads in flight), then
        // the arithmetic and the stores (the compiler may not move a shared load above an earlier shared store)
        Cx<T> w[B], za[B], zb[B];
        if (IT == 8 && wb) {
            const Cx<T> base = *wb;
            LDD_UNROLL
            for (int i = 0; i < B; ++i) w[i] = wn_of<T>(base, it0 + i);
        } else {
            LDD_UNROLL
            for (int i = 0; i < B; ++i) w[i] = WN[tid + (it0 + i) * NT];
        }
        LDD_UNROLL
        for (int i = 0; i < B; ++i) {
            za[i] = Z[pidx<PAD>(tid) + (it0 + i) * pstride<PAD>(NT)];
            zb[i] = Z[(it0 + i == 0 && tid == 0) ? 0 : pidx<PAD>(M - tid) - (it0 + i) * pstride<PAD>(NT)];    // k = 0 has no partner
        }
        LDD_UNROLL
        for (int i = 0; i < B; ++i) {
            const int k = tid + (it0 + i) * NT;
            if (it0 + i == 0 && k == 0) {
                Cx<T> z = za[i];
                Z[0] = mk<T>(z.x + z.y, z.x - z.y);
            } else {
                const int ik = pidx<PAD>(tid) + (it0 + i) * pstride<PAD>(NT), im = pidx<PAD>(M - tid) - (it0 + i) * pstride<PAD>(NT);
                Cx<T> a = za[i], b = conj(zb[i]);
                Cx<T> E = scale(a + b, half);
                Cx<T> Od = scale(mul_mj(a - b), half);
                Cx<T> Tw = w[i] * Od;
                Z[ik] = E + Tw;
                Z[im] = conj(E - Tw);
            }
        }
    }
    if (tid == 0) Z[pidx<PAD>(M / 2)] = conj(Z[pidx<PAD>(M / 2)]);      // k = M/2 pairs with itself
}

template <class T, bool PAD, int M, int NT>
__device__ inline void tangle_static(const Cx<T>* D, Cx<T>* Q, const Cx<T>* __restrict__ F, const Cx<T>* __restrict__ WN, int tid,
                                     const Cx<T>* wb = nullptr) {
    constexpr int IT = (M / 2) / NT, B = (IT % 4 == 0) ? (sizeof(T) == 4 ? 4 : 2) : 1;
    const T half = (T)0.5;
    const bool cw = IT == 8 && wb;
    const Cx<T> wbase = cw ? *wb : mk<T>((T)1, (T)0);
    LDD_UNROLL
    for (int it0 = 0; it0 < IT; it0 += B) {
        Cx<T> fa[B], fb[B], w[B], da[B], db[B];
        LDD_UNROLL
        for (int i = 0; i < B; ++i) {
            const int k = tid + (it0 + i) * NT;
            fa[i] = F[k];
            fb[i] = F[M - k];
            w[i] = cw ? wn_of<T>(wbase, it0 + i) : WN[k];
        }
        LDD_UNROLL
        for (int i = 0; i < B; ++i) {
            da[i] = D[pidx<PAD>(tid) + (it0 + i) * pstride<PAD>(NT)];
            db[i] = D[(it0 + i == 0 && tid == 0) ? 0 : pidx<PAD>(M - tid) - (it0 + i) * pstride<PAD>(NT)];
        }
        LDD_UNROLL
        for (int i = 0; i < B; ++i) {
            const int k = tid + (it0 + i) * NT;
            if (it0 + i == 0 && k == 0) {
                Cx<T> d = da[i];
                T y0 = d.x * fa[i].x, ym = d.y * fb[i].x;
                Q[0] = mk<T>((y0 + ym) * half, -(y0 - ym) * half);
            } else {
                const int ik = pidx<PAD>(tid) + (it0 + i) * pstride<PAD>(NT), im = pidx<PAD>(M - tid) - (it0 + i) * pstride<PAD>(NT);
                Cx<T> a = da[i] * fa[i], b = conj(db[i] * fb[i]);
                Cx<T> E = scale(a + b, half);
                Cx<T> Od = mulc(scale(a - b, half), w[i]);
                Cx<T> q = E + mul_pj(Od);
                Cx<T> qm = conj(E) + mul_pj(conj(Od));
                Q[ik] = conj(q);
                Q[im] = conj(qm);
            }
        }
    }
    if (tid == 0) Q[pidx<PAD>(M / 2)] = D[pidx<PAD>(M / 2)] * F[M / 2];   // k = M/2: q = conj(a), stored conj(q)
}

// SP: the block arrays live in the global scratch and a padded shared-memory buffer is the ping-pong
// partner of every length-M transform (every second Stockham pass stays on chip: the float64 lane is
// bound by L2 traffic, ~7 TB/s chip-wide at 9 Gsamples/s).
// CM != 0: the transform length M is the compile-time constant CM (and blockDim.x == NT, plan = radix 16
// while possible): the default block length gets fully constant-folded indexing.

// Per-kernel constants of the sync scan (step J): a thread owns CH consecutive samples.
struct ScanConsts {
    int CH;
    double c, Ach, A32, AchLane, AchLane1, cN, cn0;
};

__device__ inline ScanConsts scan_consts(const DemodParams& p, int N, int nthr, int tid) {
    ScanConsts k;
    k.CH = N / nthr;           // N and nthr are powers of two, CH >= 1
    const int lane = tid & 31;
    k.c = p.fp_c;
    k.Ach = pow(k.c, (double)k.CH);
    k.A32 = pow(k.Ach, 32.0);
    k.AchLane = pow(k.Ach, (double)lane);
    k.AchLane1 = pow(k.Ach, (double)(lane + 1));
    k.cN = pow(k.c, (double)N);
    k.cn0 = pow(k.c, (double)(tid * k.CH));
    return k;
}

struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

// Tables of the sync recursion's input term (shared memory, filled once per CTA): the term b0 s[n] + b1 s[n-1] takes four
// values, and four steps of the recursion from a zero state take 32 (five consecutive decisions).
struct ScanTab {
    double u[4];       // index s[n] << 1 | s[n-1]
    double q[32];      // index bits 0..4 = s[n-1], s[n], .., s[n+3]: ((u_n c + u_n+1) c + u_n+2) c + u_n+3
    double c4;         // c^4
};
__device__ inline void scan_tab_fill(ScanTab& tb, const DemodParams& p, int tid) {
    const double c = p.fp_c;
    if (tid < 4) tb.u[tid] = ((tid & 2) ? p.fp_b0 : 0.0) + ((tid & 1) ? p.fp_b1 : 0.0);
    if (tid < 32) {
        double a = 0.0;
        for (int j = 0; j < 4; ++j) {
            const int two = (tid >> j) & 3;
            a = fma(c, a, ((two & 2) ? p.fp_b0 : 0.0) + ((two & 1) ? p.fp_b1 : 0.0));
        }
        tb.q[tid] = a;
    }
    if (tid == 0) tb.c4 = (c * c) * (c * c);
}

// Step J for 32 samples per thread (the default block length on 512 threads), shared by every lane so that equal sync
// decisions give bit-equal demod_sync planes: `mask` holds this thread's 32 decisions (bit i = sample 32 tid + i); the
// recursion y[n] = c y[n-1] + (b0 s[n] + b1 s[n-1]) runs in float64.  Its input term is looked up by the decision bits
// (tb.u); the chunk sums advance four samples per step (tb.q).  ys: float64 staging of N + N/32 doubles for the
// coalesced store.  Contains barriers.
__device__ inline void sync_scan32(const DemodParams& p, const ScanConsts& sc, const ScanTab& tb, unsigned mask, double* ys, int keep0,
                                   int keep1, long long o, double* s_warp, double* s_total, unsigned* s_last) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const double c = sc.c;
    typedef float T;
    PHASE_BEGIN();
    // decision of the sample before this thread's chunk (circular)
    unsigned up = __shfl_up_sync(0xffffffffu, mask, 1);
    if (lane == 31) s_last[warp] = mask;
    __syncthreads();
    PHASE(16);
    if (lane == 0) up = s_last[warp == 0 ? (nthr >> 5) - 1 : warp - 1];
    const unsigned long long m2 = ((unsigned long long)mask << 1) | (unsigned long long)(up >> 31);     // bit i: s[i-1], bit i+1: s[i]
    double acc = 0.0;
    {
        const double c4 = tb.c4;
        double q[8];
        LDD_UNROLL
        for (int g = 0; g < 8; ++g) q[g] = tb.q[(unsigned)(m2 >> (4 * g)) & 31u];
        LDD_UNROLL
        for (int g = 0; g < 8; ++g) acc = fma(c4, acc, q[g]);
    }
    PHASE(17);
    // inclusive scan of the affine maps y -> Ach*y + acc over threads
    double incl = acc, mult = sc.Ach;
    for (int d = 1; d < 32; d <<= 1) {
        double upv = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += mult * upv;
        mult *= mult;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    PHASE(18);
    // state entering this warp (zero initial state): carry_w = sum_{v<w} A32^(w-1-v) s_warp[v], by a scan over the warp
    // sums that every warp repeats for itself (the sequential loop cost the last warp 15 dependent steps)
    double carry;
    {
        const int nw = nthr >> 5;
        double ci = lane < nw ? s_warp[lane] : 0.0, cm = sc.A32;
        for (int d = 1; d < nw; d <<= 1) {
            double upv = __shfl_up_sync(0xffffffffu, ci, d);
            if (lane >= d) ci += cm * upv;
            cm *= cm;
        }
        carry = __shfl_sync(0xffffffffu, ci, warp > 0 ? warp - 1 : 0);
        if (warp == 0) carry = 0.0;
    }
    if (tid == nthr - 1) {
        double tot = sc.AchLane1 * carry + incl;
        *s_total = tot / (1.0 - sc.cN);     // periodic steady state y[-1]
    }
    __syncthreads();
    PHASE(19);
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    double st = (lane == 0 ? 0.0 : excl) + sc.AchLane * carry + sc.cn0 * *s_total;
    double* yt = ys + 33 * tid;          // n + (n >> 5) for n = 32 tid + i
    LDD_UNROLL
    for (int i0 = 0; i0 < 32; i0 += 8) {
        double u[8];
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) u[i] = tb.u[(unsigned)(m2 >> (i0 + i)) & 3u];
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            st = fma(c, st, u[i]);
            yt[i0 + i] = st;
        }
    }
    __syncthreads();
    PHASE(20);
    // coalesced copy of the kept samples (ys index n + (n >> 5); 512 consecutive samples per round)
    double* outk = (double*)p.plane[LDD_P_SYNC] + o - keep0;
    constexpr int SB = 6;
    for (int n0 = keep0 + tid; n0 < keep1; n0 += SB * nthr) {
        double v[SB];
        LDD_UNROLL
        for (int i = 0; i < SB; ++i) {
            const int n = n0 + i * nthr;
            v[i] = n < keep1 ? ys[n + (n >> 5)] : 0.0;
        }
        LDD_UNROLL
        for (int i = 0; i < SB; ++i) {
            const int n = n0 + i * nthr;
            if (n < keep1) st_stream(&outk[n], v[i]);
        }
    }
    PHASE(21);
}

#include "ldd_demod8k.cuh"

// One block of N samples through the whole chain.  smem: the dynamic shared memory of the CTA (block arrays when PAD,
// the ping-pong partner when SP); scratch_slot: this CTA's slice of the global scratch (when !PAD); stw: per-thread
// twiddles (compile-time plan, float32).  Returns (block-uniform) whether a demod_05 sample of the block lies within
// p.flag_margin of a sync threshold (p.flag_margin > 0 only).  Ends with a barrier.
template <class T, int NT, bool PAD, bool SP, int CM>
__device__ inline int demod_block(const DemodParams& p, const int blk, char* smem, void* scratch_slot, Cx<T>* stw, const ScanConsts& sc,
                                  const ScanTab* stab) {
    const int tid = threadIdx.x, nthr = CM ? NT : (int)blockDim.x;
    const int M = CM ? CM : p.M, N = CM ? 2 * CM : p.N;
    const Cx<T>* WM = (const Cx<T>*)p.WM;
    const Cx<T>* WN = (const Cx<T>*)p.WN;
    const Cx<T>* Hv = (const Cx<T>*)p.Hv;

    // PAD <=> the block arrays live in shared memory (decided at compile time so that the compiler can
    // keep every derived pointer in the shared address space: LDS/STS with 32-bit addresses instead of
    // generic loads with 64-bit address arithmetic)
    Cx<T>* b0;
    Cx<T>* sp = nullptr;
    if (PAD) {
        b0 = (Cx<T>*)smem;
    } else {
        b0 = (Cx<T>*)scratch_slot;
        if (SP) sp = (Cx<T>*)smem;
    }
    const bool sp_ok = SP && ((CM ? static_npass(CM ? CM : 2) : p.plan_m.npass) & 1) == 0;
    constexpr bool TW8K = (CM == 8192 && NT == 512 && sizeof(T) == 4);
    // length-M transform of `a`; `other` is a free array usable as the partner when shared memory is not
    auto FFTM = [&](Cx<T>* a, Cx<T>* other) -> Cx<T>* {
        if constexpr (TW8K) {
            if (SP) return fft8k_run<T, PAD, true>(a, sp, stw, tid);
            return fft8k_run<T, PAD, PAD>(a, other, stw, tid);
        } else if constexpr (CM != 0) {
            if (SP && sp_ok) return fft_run_static<T, CM, NT, PAD, true>(a, sp, WM, tid);
            return fft_run_static<T, CM, NT, PAD, PAD>(a, other, WM, tid);
        } else {
            if (SP && sp_ok) return fft_run<T, PAD, true>(a, sp, p.plan_m, WM, 1, tid, nthr);
            return fft_run<T, PAD>(a, other, p.plan_m, WM, 1, tid, nthr);
        }
    };
    Cx<T>* b1 = b0 + pspan<PAD>(M);
    Cx<T>* b2 = b1 + pspan<PAD>(M);
#define IX(i) pidx<PAD>(i)
    __shared__ double s_warp[32];
    __shared__ double s_total;
    __shared__ unsigned s_last[32];
    const int CH = sc.CH;
    const int lane = tid & 31, warp = tid >> 5;
    const double c = sc.c, Ach = sc.Ach, A32 = sc.A32, AchLane = sc.AchLane, AchLane1 = sc.AchLane1, cN = sc.cN, cn0 = sc.cn0;
    int flagged = 0;
    PHASE_BEGIN();
    {
        const long long in0 = p.first_sample + (long long)blk * p.stride;
        const long long o = (long long)blk * p.stride;
        long long copylen = p.stride;
        if (o + (N - p.blockcut) > p.total_out) copylen = p.total_out - o;
        if (copylen > N - p.blockcut) copylen = N - p.blockcut;
        if (copylen < 0) copylen = 0;
        const int keep0 = p.blockcut, keep1 = p.blockcut + (int)copylen;

        // A. samples -> z[n] = x[2n] + j x[2n+1]
        bool loaded = false;
        if (in0 + N > p.rf_limit) {
            // the last block of a capture whose end does not fall on the block grid: zeros beyond the end (the range
            // planner only lets such a block in where the reference's own last window still has real samples to read)
            for (int n = tid; n < M; n += nthr) {
                const long long s = in0 + 2 * n;
                const int s0 = s < p.rf_limit ? fetch_sample(p.rf, p.fmt, s) : 0;
                const int s1 = s + 1 < p.rf_limit ? fetch_sample(p.rf, p.fmt, s + 1) : 0;
                b0[IX(n)] = mk<T>((T)s0, (T)s1);
            }
            loaded = true;
        }
        if constexpr (CM != 0) {
            if (loaded) {
            } else
            if (p.fmt == LDD_FMT_U8 && ((((uintptr_t)p.rf + (uintptr_t)in0) & 3) == 0)) {
                // four samples per 32-bit load, all loads of a thread in flight together
                constexpr int IT = CM / 2 / NT;
                const unsigned* r32 = (const unsigned*)((const unsigned char*)p.rf + in0);
                unsigned v[IT];
                LDD_UNROLL
                for (int i = 0; i < IT; ++i) v[i] = r32[tid + i * NT];
                LDD_UNROLL
                for (int i = 0; i < IT; ++i) {
                    const int n = 2 * (tid + i * NT);
                    b0[IX(n)] = mk<T>((T)(int)(v[i] & 0xffu), (T)(int)((v[i] >> 8) & 0xffu));
                    b0[IX(n + 1)] = mk<T>((T)(int)((v[i] >> 16) & 0xffu), (T)(int)(v[i] >> 24));
                }
                loaded = true;
            }
        }
        if (loaded) {
        } else if (p.fmt == LDD_FMT_U8 && ((in0 & 1) == 0) && ((((uintptr_t)p.rf) & 1) == 0)) {
            const unsigned short* r16 = (const unsigned short*)((const unsigned char*)p.rf + in0);
            for (int n = tid; n < M; n += nthr) {
                unsigned v = r16[n];
                b0[IX(n)] = mk<T>((T)(int)(v & 0xffu), (T)(int)(v >> 8));
            }
        } else {
            for (int n = tid; n < M; n += nthr) {
                int s0 = fetch_sample(p.rf, p.fmt, in0 + 2 * n);
                int s1 = fetch_sample(p.rf, p.fmt, in0 + 2 * n + 1);
                b0[IX(n)] = mk<T>((T)s0, (T)s1);
            }
        }
        __syncthreads();
        PHASE(0);

        // B/C. X = rfft(x)
        Cx<T>* X = FFTM(b0, b1);
        PHASE(1);
        Cx<T>* f1 = (X == b0) ? b1 : b0;      // free
        Cx<T>* f2 = b2;                        // free
        const Cx<T>* wnb = TW8K ? stw + 3 * NT + tid : nullptr;         // this thread's W_N^tid (float32 lane, default block length)
        if constexpr (CM != 0) untangle_static<T, PAD, CM, NT>(X, WN, tid, wnb);
        else untangle<T, PAD>(X, M, WN, tid, nthr);
        __syncthreads();
        PHASE(2);

        // D. analog audio, phase 1 (lddecode_core.py:322-326): two length-A inverse transforms of a
        //    slice of X, FM discriminator at freq_arf, + audio_lowfreq.
        if (p.A > 0) {
            const int A = p.A, hA = A / 2;
            const Cx<T>* AL = (const Cx<T>*)p.AL;
            const Cx<T>* AR = (const Cx<T>*)p.AR;
            Cx<T>* gl = f1;
            Cx<T>* gr = f2;
            for (int j = tid; j < A; j += nthr) {
                Cx<T> xa = (j < hA) ? X[IX(p.a_lo + j)] : conj(X[IX(p.a_hi - (j - hA))]);
                gl[IX(j)] = conj(xa * AL[j]);
                gr[IX(j)] = conj(xa * AR[j]);
            }
            __syncthreads();
            Cx<T>* rl = fft_run<T, PAD>(gl, gl + pspan<PAD>(A), p.plan_a, WM, p.wstride_a, tid, nthr);
            Cx<T>* rr = fft_run<T, PAD>(gr, gr + pspan<PAD>(A), p.plan_a, WM, p.wstride_a, tid, nthr);
            // angles in place (x: left, y: right of the same sample) then neighbour difference
            Cx<T>* ang = (rl == gl) ? gl + pspan<PAD>(A) : gl;   // the other half of f1
            for (int j = tid; j < A; j += nthr) {
                Cx<T> l = rl[IX(j)], r = rr[IX(j)];
                ang[IX(j)] = mk<T>(Math<T>::atan2(-l.y, l.x), Math<T>::atan2(-r.y, r.x));
            }
            __syncthreads();
            const int a0 = keep0 / p.audio_ds, a1 = keep1 / p.audio_ds;
            const long long ao = o / p.audio_ds;
            const double twopi = 6.283185307179586476925286766559;
            for (int j = a0 + tid; j < a1; j += nthr) {
                double dl = 0.0, dr = 0.0;
                if (j > 0) {
                    Cx<T> c1 = ang[IX(j)], c0 = ang[IX(j - 1)];
                    dl = (double)c1.x - (double)c0.x;
                    dr = (double)c1.y - (double)c0.y;
                    if (dl < 0) dl += twopi;
                    if (dr < 0) dr += twopi;
                }
                long long oi = ao + (j - a0);
                if (oi < p.audio_total) {
                    p.audio_l[oi] = dl * p.audio_scale + p.audio_lowfreq;
                    p.audio_r[oi] = dr * p.audio_scale + p.audio_lowfreq;
                }
            }
            __syncthreads();
            PHASE(3);
        }

        // Per-block MTF level (CAV discs: the reference lowers mtf_level by 1e-4 per frame, lddecode_core.py:1300-1306):
        // the table holds RFVideo * MTF^L0; a block whose frame is n frames past the ramp's origin needs MTF^(L0 + n step),
        // i.e. Hv * exp(delta ln MTF) with |delta ln MTF| < 0.05 -- a third-order series (error < 3e-7 relative at the
        // largest delta the host lets through before it re-bases L0).
        T dl = (T)0;
        if (p.mtf_period > 0.0) {
            const double centre = (double)o + 0.5 * (double)p.stride;
            double dlt = p.mtf_step * floor((centre - p.mtf_pos0) / p.mtf_period);
            if (dlt < -p.mtf_level0) dlt = -p.mtf_level0;                    // max(level, 0)
            if (centre < p.mtf_hold_until) dlt = p.mtf_hold_level - p.mtf_level0;   // the decode's first frame: start-up level
            dl = (T)dlt;
        }
        const Cx<T>* LnM = (const Cx<T>*)p.lnM;
        // (two instances of step E under one block-uniform branch, so that the common case keeps its code)
        auto HVr = [&](int k, Cx<T> hv) -> Cx<T> {
            const Cx<T> z = scale(LnM[k], dl);
            const Cx<T> z2 = z * z;
            const Cx<T> e = mk<T>((T)1 + z.x, z.y) + scale(z2, (T)0.5) + scale(z2 * z, (T)(1.0 / 6.0));
            return hv * e;
        };
        // E. Y = X_full * Hv, split into even/odd output samples: U[k] = Y[k] + Y[k+M],
        //    V[k] = (Y[k] - Y[k+M]) W_N^{-k}; stored conjugated for inverse-by-forward.
        Cx<T>* U = f1;
        Cx<T>* V = f2;
        auto estep = [&](int k, int ik, int im, Cx<T> xa, Cx<T> xb, Cx<T> h0, Cx<T> h1, Cx<T> h2, Cx<T> h3, Cx<T> w) {   // Hv[k], Hv[k+M], Hv[M-k], Hv[2M-k], WN[k]
            if (k == 0) {
                Cx<T> x = xa;
                Cx<T> y0 = scale(h0, x.x), y1 = scale(h1, x.y);
                U[0] = conj(y0 + y1);
                V[0] = conj(y0 - y1);
            } else {
                Cx<T> y0 = xa * h0, y1 = conj(xb) * h1;
                U[ik] = conj(y0 + y1);
                V[ik] = conj(mulc(y0 - y1, w));
                if (k != M - k) {
                    Cx<T> z0 = xb * h2, z1 = conj(xa) * h3;
                    U[im] = conj(z0 + z1);
                    // W_N^{-(M-k)} = -conj(W_N^{-k}) = -W_N^{k}
                    Cx<T> d = z0 - z1;
                    V[im] = conj(mk<T>(-d.x, -d.y) * w);
                }
            }
        };
        auto stepE = [&](auto ramp) {
            auto HV = [&](int k) -> Cx<T> {
                Cx<T> hv = Hv[k];
                if constexpr (decltype(ramp)::value) hv = HVr(k, hv);
                return hv;
            };
            if constexpr (CM != 0) {
                constexpr int IT = CM / 2 / NT, B = (IT % 4 == 0) ? (sizeof(T) == 4 ? 4 : 2) : 1;
                LDD_UNROLL
                for (int it0 = 0; it0 < IT; it0 += B) {
                    Cx<T> h0[B], h1[B], h2[B], h3[B], w[B], xa[B], xb[B];
                    LDD_UNROLL
                    for (int i = 0; i < B; ++i) {
                        const int k = tid + (it0 + i) * NT;
                        h0[i] = HV(k); h1[i] = HV(k + M); h2[i] = HV(M - k); h3[i] = HV((2 * M - k) & (2 * M - 1));
                        if constexpr (TW8K) w[i] = wn_of<T>(*wnb, it0 + i);
                        else w[i] = WN[k];
                    }
                    LDD_UNROLL
                    for (int i = 0; i < B; ++i) {
                        xa[i] = X[IX(tid) + (it0 + i) * pstride<PAD>(NT)];
                        xb[i] = X[(it0 + i == 0 && tid == 0) ? 0 : IX(M - tid) - (it0 + i) * pstride<PAD>(NT)];
                    }
                    LDD_UNROLL
                    for (int i = 0; i < B; ++i)
                        estep(tid + (it0 + i) * NT, IX(tid) + (it0 + i) * pstride<PAD>(NT), IX(M - tid) - (it0 + i) * pstride<PAD>(NT), xa[i], xb[i],
                              h0[i], h1[i], h2[i], h3[i], w[i]);
                }
                if (tid == 0) estep(M / 2, IX(M / 2), IX(M / 2), X[IX(M / 2)], X[IX(M / 2)], HV(M / 2), HV(M / 2 + M), HV(M / 2), HV(M / 2 + M), WN[M / 2]);
            } else {
                for (int k = tid; k <= M / 2; k += nthr)
                    estep(k, IX(k), IX(M - k), X[IX(k)], X[IX(k == 0 ? 0 : M - k)], HV(k), HV(k + M), HV(M - k), HV((2 * M - k) & (2 * M - 1)), WN[k]);
            }
        };
        if (dl != (T)0) stepE(TrueTag{});
        else stepE(FalseTag{});
        __syncthreads();
        PHASE(4);

        // F. h[2n] = conj(ru[n]), h[2n+1] = conj(rv[n]) (up to a positive scale)
        Cx<T>* ru;
        Cx<T>* fu;
        Cx<T>* rv;
        Cx<T>* fv;
        if (SP && sp_ok) {
            // the two transforms are independent: run them pass by pass with one barrier per pass pair
            if constexpr (TW8K) fft8k_run_pair<T, PAD, true, PAD>(U, sp, V, X, stw, tid);
            else if constexpr (CM != 0) fft_run_pair_static<T, CM, NT, PAD, true, PAD>(U, sp, V, X, WM, tid);
            else fft_run_pair<T, PAD, true, PAD>(U, sp, V, X, p.plan_m, WM, 1, tid, nthr);
            ru = U; rv = V; fu = X; fv = X;
        } else {
            ru = FFTM(U, X);
            fu = (ru == U) ? X : U;
            rv = FFTM(V, fu);
            fv = (rv == V) ? fu : V;
        }
        PHASE(5);

        // G. FM discriminator (lddutils.py:320-334): angle, neighbour difference, fold to [0, 2pi),
        //    scale to Hz; minus ire0; packed for the next real transform.
        if constexpr (CM != 0) {
            constexpr int IT = CM / NT;
            LDD_UNROLL
            for (int i = 0; i < IT; ++i) {
                const int ix = IX(tid) + i * pstride<PAD>(NT);
                Cx<T> a = ru[ix], b = rv[ix];
                ru[ix] = mk<T>(Math<T>::atan2(-a.y, a.x), Math<T>::atan2(-b.y, b.x));
            }
        } else {
            for (int n = tid; n < M; n += nthr) {
                Cx<T> a = ru[IX(n)], b = rv[IX(n)];
                ru[IX(n)] = mk<T>(Math<T>::atan2(-a.y, a.x), Math<T>::atan2(-b.y, b.x));
            }
        }
        __syncthreads();
        PHASE(6);
        {
            const T twopi = (T)6.283185307179586476925286766559;
            const T hz = (T)p.hz_per_rad, ire0 = (T)p.ire0;
            auto diff = [&](int n, int ix, int ixm) {
                Cx<T> a = ru[ix];
                T d0 = (T)0;
                if (n > 0) {
                    d0 = a.x - ru[ixm].y;
                    if (d0 < 0) d0 += twopi;
                }
                T d1 = a.y - a.x;
                if (d1 < 0) d1 += twopi;
                rv[ix] = mk<T>(d0 * hz - ire0, d1 * hz - ire0);
            };
            if constexpr (CM != 0) {
                constexpr int IT = CM / NT;
                LDD_UNROLL
                for (int i = 0; i < IT; ++i)      // pidx(n - 1) = pidx(tid - 1) + i * pitch also for tid = 0 (arithmetic shift)
                    diff(tid + i * NT, IX(tid) + i * pstride<PAD>(NT), IX(tid - 1) + i * pstride<PAD>(NT));
            } else {
                for (int n = tid; n < M; n += nthr) diff(n, IX(n), IX(n - 1));
            }
        }
        __syncthreads();
        PHASE(7);

        // H. D = rfft(demod - ire0)
        Cx<T>* D = FFTM(rv, ru);
        PHASE(8);
        // the two free arrays; the filtered block lands in g1 for an even number of passes and in g2 for an
        // odd one: make that an end array (b0 or b2) so that step J finds two adjacent free arrays
        Cx<T>* g1 = (D == rv) ? ru : rv;
        Cx<T>* g2 = fv;
        if (PAD) {
            const bool even = ((CM ? static_npass(CM ? CM : 2) : p.plan_m.npass) & 1) == 0;
            Cx<T>* land = (g1 == b1) ? g2 : g1;
            Cx<T>* oth = (land == g1) ? g2 : g1;
            g1 = even ? land : oth;
            g2 = even ? oth : land;
        }
        if constexpr (CM != 0) untangle_static<T, PAD, CM, NT>(D, WN, tid, wnb);
        else untangle<T, PAD>(D, M, WN, tid, nthr);
        __syncthreads();
        PHASE(9);

        // I. post filters.  Order: video, burst, (pilot), video05 last because its whole block feeds the sync scan.
        Cx<T>* r05 = nullptr;
        for (int oi = 0; oi < 4; ++oi) {
            const int m = (0x1320 >> (4 * oi)) & 15;           // 0, 2, 3, 1
            if ((m >= p.nfilt || p.only05) && m != 1) continue;
            if constexpr (CM != 0) tangle_static<T, PAD, CM, NT>(D, g1, (const Cx<T>*)p.F[m], WN, tid, wnb);
            else tangle<T, PAD>(D, g1, (const Cx<T>*)p.F[m], M, WN, tid, nthr);
            __syncthreads();
            PHASE(10);
            Cx<T>* r = FFTM(g1, g2);
            PHASE(11);
            static_assert(LDD_P_DEMOD == 0 && LDD_P_DEMOD05 == 1 && LDD_P_BURST == 3 && LDD_P_PILOT == 4, "plane order");
            float* out = (float*)p.plane[m < 2 ? m : m + 1];
            const T addc = (T)p.addc[m];
            // kept samples 2n, 2n+1 -> out[o + 2n - keep0]; keep0, o and copylen parity: handle singly
            if (((keep0 | keep1) & 1) == 0 && ((o & 1) == 0) && ((((uintptr_t)out) & 7) == 0)) {
                // whole (even, odd) sample pairs inside the kept region: one 8-byte streaming store each
                float2* out2 = (float2*)(out + o) - keep0 / 2;
                if (CM != 0) {
                    // constant strides (pidx(n + NT) = pidx(n) + pitch for any n), unrolled in batches of five
                    const int nlo = keep0 / 2 + tid, nhi = keep1 / 2;
                    const Cx<T>* r0 = r + IX(nlo);
                    float2* o0 = out2 + nlo;
                    constexpr int SB = 5;
                    for (int i0 = 0; nlo + i0 * NT < nhi; i0 += SB) {
                        Cx<T> v[SB];
                        LDD_UNROLL
                        for (int i = 0; i < SB; ++i) v[i] = (nlo + (i0 + i) * NT < nhi) ? r0[(i0 + i) * pstride<PAD>(NT)] : mk<T>((T)0, (T)0);
                        LDD_UNROLL
                        for (int i = 0; i < SB; ++i)
                            if (nlo + (i0 + i) * NT < nhi)
                                st_stream(&o0[(i0 + i) * NT], make_float2((float)(v[i].x + addc), (float)(-v[i].y + addc)));
                    }
                } else {
                    for (int n = keep0 / 2 + tid; n < keep1 / 2; n += nthr) {
                        Cx<T> v = r[IX(n)];
                        st_stream(&out2[n], make_float2((float)(v.x + addc), (float)(-v.y + addc)));
                    }
                }
            } else {
                for (int n = tid; n < M; n += nthr) {
                    Cx<T> v = r[IX(n)];
                    int i0 = 2 * n, i1 = 2 * n + 1;
                    float v0 = (float)(v.x + addc), v1 = (float)(-v.y + addc);
                    if (i0 >= keep0 && i0 < keep1) st_stream(&out[o + (i0 - keep0)], v0);
                    if (i1 >= keep0 && i1 < keep1) st_stream(&out[o + (i1 - keep0)], v1);
                }
            }
            if (m == 1) r05 = r;
            __syncthreads();
            PHASE(12);
        }

        // J. sync: s[n] = lo <= demod_05[n] <= hi (lddecode_core.py:308); demod_sync = circular
        //    y[n] = b0 s[n] + b1 s[n-1] + c y[n-1]  (== ifft(fft(s) * FPsync), :310).
        {
            const T* x05 = (const T*)r05;     // interleaved: sample 2n = r.x, 2n+1 = -r.y
            const double add = p.addc[1] + p.sync_ref;
            const int n0 = tid * CH;
            auto val05 = [&](int n) -> double {
                n = (n + N) & (N - 1);
                double v = (double)x05[2 * IX(n >> 1) + (n & 1)];
                if (n & 1) v = -v;
                return v + add;
            };
            auto insync = [&](double v) -> double { return (v >= p.sync_lo && v <= p.sync_hi) ? 1.0 : 0.0; };
            double* ys_fast = nullptr;
            if (PAD) ys_fast = (r05 == b0) ? (double*)b1 : (r05 == b2) ? (double*)b0 : nullptr;
            else if (SP && (const void*)r05 != (const void*)sp) ys_fast = (double*)sp;
            if (CH == 32 && ys_fast && stab) {
                // default geometry: decisions into a bit mask, recursion and store by the code every lane shares
                unsigned mask = 0u;
                int near = 0;
                for (int i = 0; i < 32; ++i) {
                    const double v = val05(n0 + i);
                    if (v >= p.sync_lo && v <= p.sync_hi) mask |= (1u << i);
                    if (p.flag_margin > 0.0) near |= (fabs(v - p.sync_lo) < p.flag_margin) | (fabs(v - p.sync_hi) < p.flag_margin);
                }
                if (p.flag_margin > 0.0) flagged = __syncthreads_or(near | (in0 + N > p.rf_limit));
                sync_scan32(p, sc, *stab, mask, ys_fast, keep0, keep1, o, s_warp, &s_total, s_last);
                PHASE(13);
            } else {
            double sprev = insync(val05(n0 - 1));
            const double sprev0 = sprev;
            // the binary decisions of this thread's chunk, evaluated once (CH <= 64), else re-evaluated
            const bool use_mask = CH <= 64;
            unsigned long long mask = 0ull;
            double acc = 0.0;
            int near = 0;
            for (int i = 0; i < CH; ++i) {
                const double v = val05(n0 + i);
                const double s = insync(v);
                if (use_mask && s != 0.0) mask |= (1ull << i);
                acc = c * acc + (p.fp_b0 * s + p.fp_b1 * sprev);
                sprev = s;
                // mixed lane: is this float32 sample close enough to a threshold that float64 could decide otherwise?
                if (p.flag_margin > 0.0) near |= (fabs(v - p.sync_lo) < p.flag_margin) | (fabs(v - p.sync_hi) < p.flag_margin);
            }
            // (a block that reaches past the end of the capture is always re-run: next to the zero padding the analytic
            // signal fades out and its angle is rounding noise, so float32 decisions there mean nothing)
            if (p.flag_margin > 0.0) flagged = __syncthreads_or(near | (in0 + N > p.rf_limit));
            // inclusive scan of the affine maps y -> Ach*y + acc over threads
            double incl = acc, mult = Ach;
            for (int d = 1; d < 32; d <<= 1) {
                double up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += mult * up;
                mult *= mult;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            PHASE(13);
            double carry = 0.0;                 // state entering this warp (zero initial state)
            for (int w = 0; w < warp; ++w) carry = A32 * carry + s_warp[w];
            if (tid == nthr - 1) {
                double tot = AchLane1 * carry + incl;
                s_total = tot / (1.0 - cN);     // periodic steady state y[-1]
            }
            __syncthreads();
            // state entering this thread's chunk
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            double st = (lane == 0 ? 0.0 : excl) + AchLane * carry + cn0 * s_total;
            double* out = (double*)p.plane[LDD_P_SYNC];     // float64: peak search must see the reference's ordering
            // A thread's chunk is contiguous, so storing it directly puts the lanes of a warp CH*8 bytes apart
            // (one sector per lane and store).  When two adjacent block arrays (or the partner buffer) are
            // free, the chunk goes to shared memory first (one padding slot per 32 keeps that conflict-free)
            // and the plane is written with fully coalesced stores.
            double* ys = nullptr;
            if (PAD) ys = (r05 == b0) ? (double*)b1 : (r05 == b2) ? (double*)b0 : nullptr;
            else if (SP && (const void*)r05 != (const void*)sp) ys = (double*)sp;
            sprev = sprev0;
            for (int i = 0; i < CH; ++i) {
                int n = n0 + i;
                double s = use_mask ? (double)((mask >> i) & 1ull) : insync(val05(n));
                st = c * st + (p.fp_b0 * s + p.fp_b1 * sprev);
                sprev = s;
                if (ys) ys[n + (n >> 5)] = st;
                else if (n >= keep0 && n < keep1) st_stream(&out[o + (n - keep0)], st);
            }
            if (ys) {
                __syncthreads();
                PHASE(14);
                double* outk = out + o - keep0;
                for (int n = keep0 + tid; n < keep1; n += nthr) st_stream(&outk[n], ys[n + (n >> 5)]);
            }
            }
        }
        __syncthreads();
        PHASE(15);
    }
    return flagged;
#undef IX
}

// shared-memory twiddle constants of the float32 lane at the default block length: of the in-place block or of the
// compile-time Stockham plan (one or the other per launch)
union CstBuf {
    d8::Consts c;
    d8::Consts64 c64;
    Cx<float> tw[4 * 512];
};

// Persistent kernel of the single-precision lanes and of the exact lane: CTA b takes blocks b, b + grid, ...; with a
// block list (second pass of the two-launch mixed lane) it takes the list's entries instead.
template <class T, int NT, bool PAD, int MINB = 1, bool SP = false, int CM = 0>
__global__ void __launch_bounds__(NT, MINB) demod_kernel(const DemodParams p) {
    const int tid = threadIdx.x, nthr = CM ? NT : (int)blockDim.x;
    LDD_DYN_SMEM(smem);
    constexpr bool TW8K = (CM == 8192 && NT == 512 && sizeof(T) == 4);
    Cx<T>* stw = nullptr;
    bool inplace = false;           // the in-place block of ldd_demod8k.cuh (needs the permuted tables)
    d8::Consts* cst = nullptr;
    if constexpr (TW8K) {
        // per-thread twiddles in shared memory: of the in-place block, or of the compile-time Stockham plan
        __shared__ CstBuf s_cst;
        inplace = PAD && p.HvP != nullptr;
        if (inplace) {
            cst = &s_cst.c;
            d8::consts_fill(*cst, (const Cx<float>*)p.WM, (const Cx<float>*)p.WN, tid);
        } else {
            stw = (Cx<T>*)s_cst.tw;
            fft_tw_fill<T, CM, NT>(stw, (const Cx<T>*)p.WM, tid, (const Cx<T>*)p.WN);
        }
        __syncthreads();
    }
    const ScanConsts sc = scan_consts(p, CM ? 2 * CM : p.N, nthr, tid);
    void* slot = PAD ? nullptr : (void*)((char*)p.scratch + (size_t)blockIdx.x * p.scratch_per_cta);
    const int nwork = p.block_list ? *p.block_count : p.nblocks;
    __shared__ double s8_warp[32];
    __shared__ double s8_total;
    __shared__ unsigned s8_last[32];
    __shared__ ScanTab s_stab;
    __shared__ unsigned long long s_bars[2];
    d8::Stage stg;
    d8::stage_init(stg, s_bars, tid);
    scan_tab_fill(s_stab, p, tid);
    __syncthreads();
    for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const int blk = p.block_list ? p.block_list[wi] : wi;
        int fl;
        if constexpr (TW8K) {
            if (inplace) fl = d8::demod_block8k(p, blk, smem, *cst, sc, s_stab, stg, s8_warp, &s8_total, s8_last);
            else fl = demod_block<T, NT, PAD, SP, CM>(p, blk, smem, slot, stw, sc, &s_stab);
        } else
        fl = demod_block<T, NT, PAD, SP, CM>(p, blk, smem, slot, stw, sc, &s_stab);
        if (p.flag_list && fl && tid == 0) { int at = atomicAdd(p.flag_count, 1); p.flag_list[at] = blk; }
    }
}

// Mixed lane, default block length, ONE launch: blocks are handed out by an atomic counter; a CTA runs a block in float32
// in shared memory and, when the block holds a demod_05 sample inside the guard band of a sync threshold, straight away
// again in float64 (only demod_05 -> demod_sync; the float64 lane's arrays go to this CTA's slice of the L2-resident
// scratch and the ping-pong partner into the shared memory the float32 arrays just vacated).  No second launch with its
// single under-filled wave, and CTAs that drew re-runs simply draw fewer blocks.  pq: the float64 parameter set.
template <int NT, int CM>
__global__ void __launch_bounds__(NT, 1) demod_mixed_kernel(const DemodParams pf, const DemodParams pq, int* queue) {
    const int tid = threadIdx.x;
    LDD_DYN_SMEM(smem);
    __shared__ CstBuf s_cst;
    __shared__ int s_blk;
    __shared__ double s8_warp[32];
    __shared__ double s8_total;
    __shared__ unsigned s8_last[32];
    __shared__ ScanTab s_stab;
    __shared__ unsigned long long s_bars[2];
    d8::Stage stg;
    d8::stage_init(stg, s_bars, tid);
    scan_tab_fill(s_stab, pf, tid);
    const bool inplace = pf.HvP != nullptr;      // the in-place block of ldd_demod8k.cuh (needs the permuted tables)
    Cx<float>* s_tw = s_cst.tw;
    d8::Consts* cst = &s_cst.c;
    if (inplace) d8::consts_fill(*cst, (const Cx<float>*)pf.WM, (const Cx<float>*)pf.WN, tid);
    else fft_tw_fill<float, CM, NT>(s_tw, (const Cx<float>*)pf.WM, tid, (const Cx<float>*)pf.WN);
    __syncthreads();
    const ScanConsts sc = scan_consts(pf, 2 * CM, NT, tid);
    void* slot64 = (void*)((char*)pq.scratch + (size_t)blockIdx.x * pq.scratch_per_cta);
    for (;;) {
        if (tid == 0) s_blk = atomicAdd(queue, 1);
        d8::stage_fence();          // (a float64 re-run wrote the arrays the next block's table staging lands in)
        __syncthreads();
        if (s_blk >= pf.nblocks) break;
        // highest block first: the block that reaches past the end of the capture is always re-run, and by the slow
        // generic float64 block (below) -- it must not be the last one drawn
        const int blk = pf.nblocks - 1 - s_blk;
        const int fl = inplace ? d8::demod_block8k(pf, blk, smem, *cst, sc, s_stab, stg, s8_warp, &s8_total, s8_last)
                               : demod_block<float, NT, true, false, CM>(pf, blk, smem, nullptr, s_tw, sc, &s_stab);
        if (fl) {
            if (tid == 0) atomicAdd(queue + 1, 1);                    // statistics: blocks re-run
            // (a block that reads zeros past the end of the capture decides on rounding noise there: it keeps the generic
            // block, whose noise is the exact lane's, so that the two lanes' sync planes stay bit-identical)
            if (inplace && pq.HvP && pf.first_sample + (long long)blk * pf.stride + 2 * CM <= pf.rf_limit) {
                // float64 on the in-place transforms (one array in shared memory); its constants borrow the float32 ones' room
                d8::rerun8k(pq, blk, smem, (Cx<double>*)slot64, s_cst.c64, sc, s_stab, s8_warp, &s8_total, s8_last);
                d8::consts_fill(*cst, (const Cx<float>*)pf.WM, (const Cx<float>*)pf.WN, tid);
                __syncthreads();
            } else {
                demod_block<double, NT, false, true, CM>(pq, blk, smem, slot64, nullptr, sc, &s_stab);
            }
        }
    }
}

#ifdef LDD_PHASE_TIMING
}  // namespace ldd
// profiling builds only: copies out (and clears) the 64 phase counters
extern "C" int ldd_debug_phases(unsigned long long* out64) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out64, ldd::g_phase, sizeof(unsigned long long) * 64) != cudaSuccess) return LDD_ECUDA;
    unsigned long long z[64] = {0};
    return cudaMemcpyToSymbol(ldd::g_phase, z, sizeof z) == cudaSuccess ? LDD_OK : LDD_ECUDA;
}
namespace ldd {
#endif

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        (void)what;
        return LDD_ECUDA;
    }
    return LDD_OK;
}

template <class T, int NT, bool PAD, int MINB = 1, bool SP = false, int CM = 0>
static int launch_variant(const DemodParams& p, int grid, cudaStream_t st, size_t smem_bytes) {
    void (*kern)(const DemodParams) = demod_kernel<T, NT, PAD, MINB, SP, CM>;
    if (smem_bytes) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    LDD_LAUNCH(kern, dim3(grid), dim3(NT), smem_bytes, st, p);
    return check_launch("demod_kernel");
}

// the compile-time variants assume the default plan (radix 16 while possible) for M = m
static bool static_plan_ok(const DemodParams& p, int m) {
    if (p.M != m || p.N != 2 * m || getenv("LDD_NO_STATIC_PLAN")) return false;
    FftPlan d = make_plan(m, 16);
    if (d.npass != p.plan_m.npass) return false;
    for (int i = 0; i < d.npass; ++i)
        if (d.radix[i] != p.plan_m.radix[i]) return false;
    return true;
}

int launch_demod_f64(const DemodParams& p, int grid, int threads, cudaStream_t st, size_t sp_bytes) {
    if (sp_bytes && threads == 512) {
        if (static_plan_ok(p, 8192)) return launch_variant<double, 512, false, 1, true, 8192>(p, grid, st, sp_bytes);
        return launch_variant<double, 512, false, 1, true>(p, grid, st, sp_bytes);
    }
    switch (threads) {
        case 1024: return launch_variant<double, 1024, false>(p, grid, st, 0);
        case 512: return launch_variant<double, 512, false>(p, grid, st, 0);
        case 2256: return launch_variant<double, 256, false, 2>(p, grid, st, 0);      // 2 CTAs of 256 threads per SM
        default: return launch_variant<double, 256, false>(p, grid, st, 0);
    }
}

bool demod_mixed_fused_ok(const DemodParams& p, int threads, size_t smem_bytes, size_t sp_bytes) {
    return threads == 512 && smem_bytes && sp_bytes && sp_bytes <= smem_bytes && static_plan_ok(p, 8192) && !getenv("LDD_MIXED_TWO_LAUNCH");
}

int launch_demod_mixed(const DemodParams& pf, const DemodParams& pq, int* queue, int grid, cudaStream_t st, size_t smem_bytes) {
    void (*kern)(const DemodParams, const DemodParams, int*) = demod_mixed_kernel<512, 8192>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    LDD_LAUNCH(kern, dim3(grid), dim3(512), smem_bytes, st, pf, pq, queue);
    return check_launch("demod_mixed_kernel");
}

int launch_demod_f32(const DemodParams& p, int grid, int threads, cudaStream_t st, size_t smem_bytes) {
    if (smem_bytes) {
        if (threads == 1024) return launch_variant<float, 1024, true>(p, grid, st, smem_bytes);
        if (static_plan_ok(p, 8192)) return launch_variant<float, 512, true, 1, false, 8192>(p, grid, st, smem_bytes);
        return launch_variant<float, 512, true>(p, grid, st, smem_bytes);
    }
    if (threads == 1024) return launch_variant<float, 1024, false>(p, grid, st, 0);
    return launch_variant<float, 512, false>(p, grid, st, 0);
}

}  // namespace ldd

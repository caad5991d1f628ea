// Block-cooperative Stockham autosort FFT (radix 2/4/8/16), templated on the real type.
//
// One CTA transforms one length-M complex sequence that lives either in shared memory (fp32
// fast lane) or in an L2-resident global scratch slice (fp64 exact lane); the code only sees
// generic pointers.  Every pass is out of place (src -> dst) followed by a CTA barrier, so a
// transform ping-pongs between two buffers.  Only the forward transform exists; inverses are
// taken as conj(FFT(conj(x))) with the conjugations folded into the fused stages around it
// (ldd_demod.cu).  Replaces numpy.fft.fft/ifft as called at lddecode_core.py:289-313, 322-326.
#pragma once
#include "ldd_platform.h"

namespace ldd {

template <class T>
struct alignas(2 * sizeof(T)) Cx {
    T x, y;
};

template <class T> LDD_HD inline Cx<T> mk(T a, T b) { Cx<T> r; r.x = a; r.y = b; return r; }
template <class T> LDD_HD inline Cx<T> operator+(Cx<T> a, Cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <class T> LDD_HD inline Cx<T> operator-(Cx<T> a, Cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <class T> LDD_HD inline Cx<T> operator*(Cx<T> a, Cx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <class T> LDD_HD inline Cx<T> conj(Cx<T> a) { return mk<T>(a.x, -a.y); }
template <class T> LDD_HD inline Cx<T> scale(Cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
// a * conj(b)
template <class T> LDD_HD inline Cx<T> mulc(Cx<T> a, Cx<T> b) { return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// multiply by -j / +j
template <class T> LDD_HD inline Cx<T> mul_mj(Cx<T> a) { return mk<T>(a.y, -a.x); }
template <class T> LDD_HD inline Cx<T> mul_pj(Cx<T> a) { return mk<T>(-a.y, a.x); }

// ---- in-register DFTs, natural order in and out (forward, e^{-2 pi i nk/R}) -------------------
template <class T, int R> struct Dft;

template <class T> struct Dft<T, 2> {
    static LDD_HD inline void run(Cx<T>* v) {
        Cx<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};
template <class T> struct Dft<T, 4> {
    static LDD_HD inline void run(Cx<T>* v) {
        Cx<T> t0 = v[0] + v[2], t1 = v[0] - v[2], t2 = v[1] + v[3], t3 = mul_mj(v[1] - v[3]);
        v[0] = t0 + t2;
        v[1] = t1 + t3;
        v[2] = t0 - t2;
        v[3] = t1 - t3;
    }
};
template <class T> struct Dft<T, 8> {
    static LDD_HD inline void run(Cx<T>* v) {
        Cx<T> e[4] = {v[0], v[2], v[4], v[6]};
        Cx<T> o[4] = {v[1], v[3], v[5], v[7]};
        Dft<T, 4>::run(e);
        Dft<T, 4>::run(o);
        const T h = (T)0.70710678118654752440;
        Cx<T> w1 = mk<T>((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);     // * (1-j)/sqrt2
        Cx<T> w2 = mul_mj(o[2]);
        Cx<T> w3 = mk<T>((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);    // * (-1-j)/sqrt2
        v[0] = e[0] + o[0]; v[4] = e[0] - o[0];
        v[1] = e[1] + w1;   v[5] = e[1] - w1;
        v[2] = e[2] + w2;   v[6] = e[2] - w2;
        v[3] = e[3] + w3;   v[7] = e[3] - w3;
    }
};
template <class T> struct Dft<T, 16> {
    static LDD_HD inline void run(Cx<T>* v) {
        Cx<T> e[8], o[8];
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<T, 8>::run(e);
        Dft<T, 8>::run(o);
        const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;   // cos, sin(pi/8)
        const T h = (T)0.70710678118654752440;
        Cx<T> w[8];
        w[0] = o[0];
        w[1] = o[1] * mk<T>(c1, -s1);
        w[2] = mk<T>((o[2].x + o[2].y) * h, (o[2].y - o[2].x) * h);
        w[3] = o[3] * mk<T>(s1, -c1);
        w[4] = mul_mj(o[4]);
        w[5] = o[5] * mk<T>(-s1, -c1);
        w[6] = mk<T>((o[6].y - o[6].x) * h, -(o[6].x + o[6].y) * h);
        w[7] = o[7] * mk<T>(-c1, -s1);
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) { v[i] = e[i] + w[i]; v[i + 8] = e[i] - w[i]; }
    }
};

// Shared-memory lane: one padding element after every 16 keeps the stride-R stores of the first
// Stockham passes off a single bank pair (a complex64 spans two banks; stride 16 elements = 128 B
// would put all 32 lanes on the same two banks).  Global-memory lane: identity.
template <bool PAD> LDD_HD inline int pidx(int i) { return PAD ? i + (i >> 4) : i; }
template <bool PAD> LDD_HD inline int pspan(int n) { return PAD ? n + (n >> 4) : n; }
// pidx(i + n) - pidx(i) for n a multiple of 16: a compile-time element stride, so that a thread's accesses of a
// pass are one base address plus immediate offsets
template <bool PAD> LDD_HD constexpr int pstride(int n) { return PAD ? n + n / 16 : n; }

// ---- one Stockham pass -------------------------------------------------------------------------
// src, dst: length-M sequences.  Ns: product of the radices of the passes already done.
// W: table of e^{-2 pi i k / Mtab}, k in [0, Mtab); wstride = Mtab / M.
template <class T, int R, bool PIN, bool POUT>
__device__ inline void fft_pass(const Cx<T>* src, Cx<T>* dst, int M, int Ns,
                                const Cx<T>* __restrict__ W, int wstride, int tid, int nthr) {
    const int nb = M / R;
    for (int j = tid; j < nb; j += nthr) {
        Cx<T> v[R];
        LDD_UNROLL
        for (int r = 0; r < R; ++r) v[r] = src[pidx<PIN>(j + r * nb)];
        const int k = j & (Ns - 1);
        if (Ns > 1) {
            // w^r by repeated squaring / short products from one table lookup (<= 4 products deep)
            Cx<T> p[R];
            p[1] = W[(size_t)k * (size_t)(nb / Ns) * (size_t)wstride];
            LDD_UNROLL
            for (int r = 2; r < R; ++r) p[r] = (r & 1) ? p[r - 1] * p[1] : p[r / 2] * p[r / 2];
            LDD_UNROLL
            for (int r = 1; r < R; ++r) v[r] = v[r] * p[r];
        }
        Dft<T, R>::run(v);
        const int j0 = (j - k) * R + k;
        LDD_UNROLL
        for (int r = 0; r < R; ++r) dst[pidx<POUT>(j0 + r * Ns)] = v[r];
    }
}

template <class T, bool PIN, bool POUT>
__device__ inline void fft_pass_any(int R, const Cx<T>* src, Cx<T>* dst, int M, int Ns, const Cx<T>* __restrict__ W,
                                    int wstride, int tid, int nthr) {
    switch (R) {
        case 16: fft_pass<T, 16, PIN, POUT>(src, dst, M, Ns, W, wstride, tid, nthr); break;
        case 8: fft_pass<T, 8, PIN, POUT>(src, dst, M, Ns, W, wstride, tid, nthr); break;
        case 4: fft_pass<T, 4, PIN, POUT>(src, dst, M, Ns, W, wstride, tid, nthr); break;
        default: fft_pass<T, 2, PIN, POUT>(src, dst, M, Ns, W, wstride, tid, nthr); break;
    }
}

// Compile-time plan (radix 16 while possible, then the remaining power of two): every index, stride
// and trip count of the passes is a constant, which removes the integer work that otherwise outweighs
// the butterflies.  Same pass sequence as make_plan(M, 16); result pointer as fft_run.
template <class T, int M, int NT, bool PA, bool PB, int Ns = 1>
__device__ inline Cx<T>* fft_run_static(Cx<T>* a, Cx<T>* b, const Cx<T>* __restrict__ W, int tid) {
    constexpr int rem = M / Ns;
    if constexpr (rem <= 1) {
        return a;
    } else {
        constexpr int R = rem >= 16 ? 16 : rem;
        fft_pass<T, R, PA, PB>(a, b, M, Ns, W, 1, tid, NT);
        __syncthreads();
        return fft_run_static<T, M, NT, PB, PA, Ns * R>(b, a, W, tid);
    }
}

// fft_run_pair with the compile-time plan.
template <class T, int M, int NT, bool PA, bool PB1, bool PB2, int Ns = 1, bool FLIP = false>
__device__ inline void fft_run_pair_static(Cx<T>* a1, Cx<T>* b1, Cx<T>* a2, Cx<T>* b2, const Cx<T>* __restrict__ W, int tid) {
    constexpr int rem = M / Ns;
    if constexpr (rem > 1) {
        constexpr int R = rem >= 16 ? 16 : rem;
        if constexpr (!FLIP) {
            fft_pass<T, R, PA, PB1>(a1, b1, M, Ns, W, 1, tid, NT);
            fft_pass<T, R, PA, PB2>(a2, b2, M, Ns, W, 1, tid, NT);
        } else {
            fft_pass<T, R, PB1, PA>(b1, a1, M, Ns, W, 1, tid, NT);
            fft_pass<T, R, PB2, PA>(b2, a2, M, Ns, W, 1, tid, NT);
        }
        __syncthreads();
        fft_run_pair_static<T, M, NT, PA, PB1, PB2, Ns * R, !FLIP>(a1, b1, a2, b2, W, tid);
    }
}

// ---- compile-time plan with per-thread twiddles -------------------------------------------------
// For M = 16 NT (8192 points, 512 threads: passes 16, 16, 16, 2) every thread works on the same butterfly
// index in every transform, so the base twiddle of each pass is a per-thread constant.  `stw` holds them
// (3 NT entries, filled once per CTA by fft_tw_fill): no table lookup in global memory is left in the passes.
// (`stw` has 4 NT entries; the last NT hold W_2M^tid, the base of the untangle / tangle twiddles: wn_of.)
template <class T, int M, int NT>
__device__ inline void fft_tw_fill(Cx<T>* stw, const Cx<T>* __restrict__ W, int tid, const Cx<T>* __restrict__ WN = nullptr) {
    static_assert(M == 16 * NT && M == 8192, "per-thread twiddles: 8192 points on 512 threads");
    constexpr int nb = M / 16;
    stw[tid] = W[(tid & 15) * (nb / 16)];               // pass 2: Ns = 16
    stw[NT + tid] = W[(tid & 255) * (nb / 256)];        // pass 3: Ns = 256
    stw[2 * NT + tid] = W[tid];                         // pass 4 (radix 2, Ns = M/2): W^(tid + i NT) = W^tid * W16^i
    if (WN) stw[3 * NT + tid] = WN[tid];
}

// e^{-2 pi i n / 32}, n in [0, 8)
template <class T> LDD_HD inline Cx<T> w32(int n) {
    switch (n & 7) {
        case 0: return mk<T>((T)1, (T)0);
        case 1: return mk<T>((T)0.98078528040323043058, (T)-0.19509032201612824808);
        case 2: return mk<T>((T)0.92387953251128675613, (T)-0.38268343236508977173);
        case 3: return mk<T>((T)0.83146961230254523567, (T)-0.55557023301960217765);
        case 4: return mk<T>((T)0.70710678118654752440, (T)-0.70710678118654752440);
        case 5: return mk<T>((T)0.55557023301960217765, (T)-0.83146961230254523567);
        case 6: return mk<T>((T)0.38268343236508977173, (T)-0.92387953251128675613);
        default: return mk<T>((T)0.19509032201612824808, (T)-0.98078528040323043058);
    }
}
// W_N^(tid + j NT) for N = 32 NT from the thread's base W_N^tid (j < 8, a compile-time constant after unrolling):
// one product instead of a table load from L2
template <class T> LDD_HD inline Cx<T> wn_of(Cx<T> base, int j) { return j == 0 ? base : base * w32<T>(j); }

// one radix-R butterfly per thread (j = tid, M / R == number of threads)
template <class T, int R, int M, int Ns, bool PIN, bool POUT>
__device__ inline void fft_pass_one(const Cx<T>* src, Cx<T>* dst, Cx<T> w1, int tid) {
    constexpr int nb = M / R;
    static_assert(nb % 16 == 0 && (Ns == 1 || Ns % 16 == 0), "padded strides");
    Cx<T> v[R];
    const Cx<T>* s0 = src + pidx<PIN>(tid);
    LDD_UNROLL
    for (int r = 0; r < R; ++r) v[r] = s0[r * pstride<PIN>(nb)];
    if (Ns > 1) {
        Cx<T> p[R];
        p[1] = w1;
        LDD_UNROLL
        for (int r = 2; r < R; ++r) p[r] = (r & 1) ? p[r - 1] * p[1] : p[r / 2] * p[r / 2];
        LDD_UNROLL
        for (int r = 1; r < R; ++r) v[r] = v[r] * p[r];
    }
    Dft<T, R>::run(v);
    const int k = tid & (Ns - 1);
    const int j0 = (tid - k) * R + k;
    Cx<T>* d0 = dst + pidx<POUT>(j0);
    LDD_UNROLL
    for (int r = 0; r < R; ++r) d0[Ns == 1 ? r : r * pstride<POUT>(Ns)] = v[r];     // Ns == 1: j0 = R tid, r < 16 stays in its group of 16
}

// e^{-2 pi i n / 16}
template <class T> LDD_HD inline Cx<T> w16(int n) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
    switch (n & 7) {
        case 0: return mk<T>((T)1, (T)0);
        case 1: return mk<T>(c1, -s1);
        case 2: return mk<T>(h, -h);
        case 3: return mk<T>(s1, -c1);
        case 4: return mk<T>((T)0, (T)-1);
        case 5: return mk<T>(-s1, -c1);
        case 6: return mk<T>(-h, -h);
        default: return mk<T>(-c1, -s1);
    }
}

// the final radix-2 pass (Ns = M/2): M / (2 NT) butterflies per thread, j = tid + i NT
template <class T, int M, int NT, bool PIN, bool POUT>
__device__ inline void fft_pass_last2(const Cx<T>* src, Cx<T>* dst, Cx<T> wt, int tid) {
    constexpr int nb = M / 2, IT = nb / NT;
    static_assert(IT == 8, "W16 constants");
    constexpr int G = sizeof(T) == 4 ? IT : 2;            // butterflies loaded together (register budget)
    const Cx<T>* s0 = src + pidx<PIN>(tid);
    Cx<T>* d0 = dst + pidx<POUT>(tid);
    LDD_UNROLL
    for (int i0 = 0; i0 < IT; i0 += G) {
        Cx<T> a[G], b[G];
        LDD_UNROLL
        for (int g = 0; g < G; ++g) {
            a[g] = s0[(i0 + g) * pstride<PIN>(NT)];
            b[g] = s0[(i0 + g) * pstride<PIN>(NT) + pstride<PIN>(nb)];
        }
        LDD_UNROLL
        for (int g = 0; g < G; ++g) {
            const int i = i0 + g;
            Cx<T> t = b[g] * (i == 0 ? wt : wt * w16<T>(i));
            d0[i * pstride<POUT>(NT)] = a[g] + t;
            d0[i * pstride<POUT>(NT) + pstride<POUT>(nb)] = a[g] - t;
        }
    }
}

// 8192-point transform on 512 threads, twiddles from stw: a -> b -> a -> b -> a (result in a).  Ends with a barrier.
template <class T, bool PA, bool PB>
__device__ inline Cx<T>* fft8k_run(Cx<T>* a, Cx<T>* b, const Cx<T>* stw, int tid) {
    constexpr int M = 8192, NT = 512;
    fft_pass_one<T, 16, M, 1, PA, PB>(a, b, mk<T>((T)1, (T)0), tid);
    __syncthreads();
    fft_pass_one<T, 16, M, 16, PB, PA>(b, a, stw[tid], tid);
    __syncthreads();
    fft_pass_one<T, 16, M, 256, PA, PB>(a, b, stw[NT + tid], tid);
    __syncthreads();
    fft_pass_last2<T, M, NT, PB, PA>(b, a, stw[2 * NT + tid], tid);
    __syncthreads();
    return a;
}

// two independent transforms, one barrier per pass pair (results in a1, a2)
template <class T, bool PA, bool PB1, bool PB2>
__device__ inline void fft8k_run_pair(Cx<T>* a1, Cx<T>* b1, Cx<T>* a2, Cx<T>* b2, const Cx<T>* stw, int tid) {
    constexpr int M = 8192, NT = 512;
    const Cx<T> one = mk<T>((T)1, (T)0);
    fft_pass_one<T, 16, M, 1, PA, PB1>(a1, b1, one, tid);
    fft_pass_one<T, 16, M, 1, PA, PB2>(a2, b2, one, tid);
    __syncthreads();
    const Cx<T> w2 = stw[tid];
    fft_pass_one<T, 16, M, 16, PB1, PA>(b1, a1, w2, tid);
    fft_pass_one<T, 16, M, 16, PB2, PA>(b2, a2, w2, tid);
    __syncthreads();
    const Cx<T> w3 = stw[NT + tid];
    fft_pass_one<T, 16, M, 256, PA, PB1>(a1, b1, w3, tid);
    fft_pass_one<T, 16, M, 256, PA, PB2>(a2, b2, w3, tid);
    __syncthreads();
    const Cx<T> w4 = stw[2 * NT + tid];
    fft_pass_last2<T, M, NT, PB1, PA>(b1, a1, w4, tid);
    fft_pass_last2<T, M, NT, PB2, PA>(b2, a2, w4, tid);
    __syncthreads();
}

constexpr int static_npass(int m) { int n = 0; while (m >= 16) { m /= 16; ++n; } return n + (m > 1 ? 1 : 0); }

struct FftPlan {
    int n;            // transform length (power of two)
    int npass;
    int radix[16];
};

inline FftPlan make_plan(int n, int rmax = 16) {
    FftPlan p;
    p.n = n;
    p.npass = 0;
    int rem = n;
    while (rem >= rmax && p.npass < 15) { p.radix[p.npass++] = rmax; rem /= rmax; }
    if (rem > 1) p.radix[p.npass++] = rem;       // a smaller power of two
    return p;
}

// Runs all passes, ping-ponging between a (padding PA) and b (padding PB); the result is in the
// returned pointer (a after an even number of passes).  Ends with a barrier.  With a in the global
// scratch and b in shared memory every second pass stays on chip.
template <class T, bool PA, bool PB = PA>
__device__ inline Cx<T>* fft_run(Cx<T>* a, Cx<T>* b, const FftPlan& plan, const Cx<T>* __restrict__ W,
                                 int wstride, int tid, int nthr) {
    int Ns = 1;
    for (int p = 0; p < plan.npass; ++p) {
        if ((p & 1) == 0) fft_pass_any<T, PA, PB>(plan.radix[p], a, b, plan.n, Ns, W, wstride, tid, nthr);
        else fft_pass_any<T, PB, PA>(plan.radix[p], b, a, plan.n, Ns, W, wstride, tid, nthr);
        Ns *= plan.radix[p];
        __syncthreads();
    }
    return (plan.npass & 1) ? b : a;
}

// Two independent transforms with the same plan, pass by pass, ONE barrier per pass pair: a warp that
// finishes its butterfly of the first transform starts loading for the second while slower warps
// still compute, so the L2 latency of one overlaps the arithmetic of the other.  (a1 <-> b1 with
// paddings PA/PB1, a2 <-> b2 with PA/PB2.)  Results as fft_run: in a* after an even number of passes.
template <class T, bool PA, bool PB1, bool PB2>
__device__ inline void fft_run_pair(Cx<T>* a1, Cx<T>* b1, Cx<T>* a2, Cx<T>* b2, const FftPlan& plan,
                                    const Cx<T>* __restrict__ W, int wstride, int tid, int nthr) {
    int Ns = 1;
    for (int p = 0; p < plan.npass; ++p) {
        if ((p & 1) == 0) {
            fft_pass_any<T, PA, PB1>(plan.radix[p], a1, b1, plan.n, Ns, W, wstride, tid, nthr);
            fft_pass_any<T, PA, PB2>(plan.radix[p], a2, b2, plan.n, Ns, W, wstride, tid, nthr);
        } else {
            fft_pass_any<T, PB1, PA>(plan.radix[p], b1, a1, plan.n, Ns, W, wstride, tid, nthr);
            fft_pass_any<T, PB2, PA>(plan.radix[p], b2, a2, plan.n, Ns, W, wstride, tid, nthr);
        }
        Ns *= plan.radix[p];
        __syncthreads();
    }
}

}  // namespace ldd

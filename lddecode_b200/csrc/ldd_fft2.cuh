// In-place 8192-point transforms for the float32 lane: decimation in frequency (natural order in, digit-permuted
// order out) for the transforms that produce a spectrum, decimation in time (digit-permuted in, natural out) for the
// ones that produce samples, so that spectra are never reordered: every element-wise step between two transforms works
// on the permuted positions (tables are uploaded in that order).  Replaces numpy.fft.fft/ifft as called at
// lddecode_core.py:289-313 for the default block length.
//
// 8192 = 2 x 16 x 16 x 16 on 512 threads.  Stage 1 (radix 2, stride 4096) is the only one that spans the CTA and is
// meant to be fused into the step that produces / consumes the samples; after it the two halves of the array are
// independent 4096-point problems owned by the two halves of the CTA (warps 0-7, 8-15): stage 2 (radix 16, stride 256)
// needs a 256-thread barrier, stage 3 (stride 16) and stage 4 (stride 1, no twiddles) stay inside one warp's 512
// contiguous elements and need __syncwarp only.  Every stage reads and writes the same 16 addresses per thread, so no
// second array and no barrier between a stage's loads and stores.
//
// Position p = 4096 q1 + 256 q2 + 16 q3 + q4 of the permuted order holds index k = q1 + 2 q2 + 32 q3 + 512 q4.
#pragma once
#include "ldd_fft.cuh"

namespace ldd {
namespace f2 {

constexpr int M = 8192, NT = 512;

// PADK padding elements after every 16: PADK = 1 keeps every access pattern below conflict-free for 8-byte accesses,
// PADK = 2 also keeps 16-byte alignment (two elements per access in stage 4)
template <int PADK> LDD_HD constexpr int span() { return M + (M >> 4) * PADK; }
template <int PADK> LDD_HD inline int pix(int i) { return i + (i >> 4) * PADK; }
template <int PADK> LDD_HD constexpr int pst(int n) { return n + (n >> 4) * PADK; }      // n a multiple of 16

// position in the permuted order -> index, and back
LDD_HD inline int idx_of_pos(int p) { return (p >> 12) | (((p >> 8) & 15) << 1) | (((p >> 4) & 15) << 5) | ((p & 15) << 9); }
LDD_HD inline int pos_of_idx(int k) { return ((k & 1) << 12) | (((k >> 1) & 15) << 8) | (((k >> 5) & 15) << 4) | ((k >> 9) & 15); }

// barrier of one half of the CTA (256 threads: warps 0-7 use barrier 1, warps 8-15 barrier 2)
__device__ inline void half_sync(int half) {
#ifdef LDD_EMU
    (void)half;
    __syncthreads();
#else
    if (half) asm volatile("bar.sync 2, 256;" ::: "memory");
    else asm volatile("bar.sync 1, 256;" ::: "memory");
#endif
}

// per-thread twiddle bases: w1 = W_8192^tid, w2 = W_4096^(tid & 255), w3 = W_256^(tid & 15)
template <class T>
struct TwT {
    Cx<T> w1, w2, w3;
};
typedef TwT<float> Tw;
// W: table of e^{-2 pi i k / 8192}, k in [0, 8192)
__device__ inline Tw tw_make(const Cx<float>* __restrict__ W, int tid) {
    Tw t;
    t.w1 = W[tid];
    t.w2 = W[2 * (tid & 255)];
    t.w3 = W[32 * (tid & 15)];
    return t;
}

// v[q] *= w^q, q = 1..15 (products at most four deep)
template <class T>
__device__ inline void tw_apply(Cx<T>* v, Cx<T> w) {
    Cx<T> p[16];
    p[1] = w;
    LDD_UNROLL
    for (int r = 2; r < 16; ++r) p[r] = (r & 1) ? p[r - 1] * p[1] : p[r / 2] * p[r / 2];
    LDD_UNROLL
    for (int r = 1; r < 16; ++r) v[r] = v[r] * p[r];
}

// ---- the stages; DIT = twiddles on the inputs (decimation in time), else on the outputs --------------------------
template <int PADK, bool DIT, class T>
__device__ inline void stage1(Cx<T>* x, Cx<T> w1, int tid) {
    Cx<T>* x0 = x + pix<PADK>(tid);
    LDD_UNROLL
    for (int i = 0; i < 8; ++i) {
        Cx<T> a = x0[i * pst<PADK>(NT)], b = x0[i * pst<PADK>(NT) + pst<PADK>(M / 2)];
        const Cx<T> w = i == 0 ? w1 : w1 * w16<T>(i);
        if (DIT) {
            b = b * w;
            x0[i * pst<PADK>(NT)] = a + b;
            x0[i * pst<PADK>(NT) + pst<PADK>(M / 2)] = a - b;
        } else {
            x0[i * pst<PADK>(NT)] = a + b;
            x0[i * pst<PADK>(NT) + pst<PADK>(M / 2)] = (a - b) * w;
        }
    }
}

template <int PADK, bool DIT, class T>
__device__ inline void stage2(Cx<T>* x, Cx<T> w2, int tid) {
    Cx<T>* x0 = x + pix<PADK>((tid >> 8) * 4096 + (tid & 255));
    Cx<T> v[16];
    LDD_UNROLL
    for (int r = 0; r < 16; ++r) v[r] = x0[r * pst<PADK>(256)];
    if (DIT) tw_apply(v, w2);
    Dft<T, 16>::run(v);
    if (!DIT) tw_apply(v, w2);
    LDD_UNROLL
    for (int r = 0; r < 16; ++r) x0[r * pst<PADK>(256)] = v[r];
}

template <int PADK, bool DIT, class T>
__device__ inline void stage3(Cx<T>* x, Cx<T> w3, int tid) {
    // warp w owns elements [512 w, 512 w + 512): two chunks of 256, 16 butterflies each
    Cx<T>* x0 = x + pix<PADK>((tid >> 4) * 256 + (tid & 15));
    Cx<T> v[16];
    LDD_UNROLL
    for (int r = 0; r < 16; ++r) v[r] = x0[r * pst<PADK>(16)];
    if (DIT) tw_apply(v, w3);
    Dft<T, 16>::run(v);
    if (!DIT) tw_apply(v, w3);
    LDD_UNROLL
    for (int r = 0; r < 16; ++r) x0[r * pst<PADK>(16)] = v[r];
}

template <int PADK, class T>
__device__ inline void stage4(Cx<T>* x, int tid) {
    Cx<T>* x0 = x + tid * pst<PADK>(16);
    Cx<T> v[16];
    if constexpr (PADK == 2 && sizeof(T) == 4) {
        // two elements per 16-byte access
        const float4* x4 = (const float4*)x0;
        LDD_UNROLL
        for (int r = 0; r < 8; ++r) {
            float4 t = x4[r];
            v[2 * r] = mk<T>(t.x, t.y);
            v[2 * r + 1] = mk<T>(t.z, t.w);
        }
    } else {
        LDD_UNROLL
        for (int r = 0; r < 16; ++r) v[r] = x0[r];
    }
    Dft<T, 16>::run(v);
    if constexpr (PADK == 2 && sizeof(T) == 4) {
        float4* x4 = (float4*)x0;
        LDD_UNROLL
        for (int r = 0; r < 8; ++r) x4[r] = make_float4(v[2 * r].x, v[2 * r].y, v[2 * r + 1].x, v[2 * r + 1].y);
    } else {
        LDD_UNROLL
        for (int r = 0; r < 16; ++r) x0[r] = v[r];
    }
}

// Stages 2-4 of the forward (DIF) transform.  Before: stage 1 done and a CTA barrier passed.  After: this thread's warp
// has finished its own 512 elements (__syncwarp passed); other warps may still be running.
template <int PADK, class T>
__device__ inline void dif_234(Cx<T>* x, const TwT<T>& tw, int tid) {
    stage2<PADK, false>(x, tw.w2, tid);
    half_sync(tid >> 8);
    stage3<PADK, false>(x, tw.w3, tid);
    __syncwarp();
    stage4<PADK>(x, tid);
    __syncwarp();
}

// Stages 4-2 of the DIT transform.  Before: this warp's 512 elements are in place (written by this warp, or a barrier
// passed).  After: a barrier of this half has NOT been passed yet: stage 1 needs a CTA barrier first.
template <int PADK, class T>
__device__ inline void dit_432(Cx<T>* x, const TwT<T>& tw, int tid) {
    stage4<PADK>(x, tid);
    __syncwarp();
    stage3<PADK, true>(x, tw.w3, tid);
    half_sync(tid >> 8);
    stage2<PADK, true>(x, tw.w2, tid);
}

}  // namespace f2
}  // namespace ldd

// The float32 block of the default geometry (N = 16384 samples, 512 threads, block arrays in shared memory) on the
// in-place transforms of ldd_fft2.cuh.  Same chain and algebra as demod_block (ldd_demod.cu; RFDecode.demodblock,
// lddecode_core.py:288-330), restructured so that
//   * spectra stay in the digit-permuted order the decimation-in-frequency transform leaves them in; the element-wise
//     steps (untangle, RF filter + even/odd split, post filters + tangle) work on positions, with the filter tables
//     uploaded in that order (coalesced reads) -- no transform pass exists only to reorder;
//   * the radix-2 stage of every transform is fused into the step that produces or consumes the samples (sample load,
//     discriminator, plane store), one shared-memory round trip less per transform;
//   * after that stage the two halves of an array are independent: warps 0-7 and 8-15 run their 4096-point problems and
//     the element-wise steps between them on 256-thread barriers, the stride-16 / stride-1 stages on __syncwarp;
//   * transforms are in place, so three arrays hold the spectrum and TWO filtered copies: the post filters run in pairs
//     that share the spectrum loads and twiddles;
//   * the untangle of the RF spectrum is fused with the RF filter step (same thread owns the same (k, M-k) pair).
// Included by ldd_demod.cu inside namespace ldd.
#pragma once
// (ldd_fft2.cuh is included by ldd_internal.h)

namespace d8 {

typedef Cx<float> C;
constexpr int M = 8192, N = 16384, NT = 512, PK = 1;
constexpr int SPAN = f2::span<PK>();
constexpr int P512 = f2::pst<PK>(512), P4096 = f2::pst<PK>(4096);
__device__ inline int PXi(int i) { return f2::pix<PK>(i); }

// static shared memory: per-thread twiddle bases and the row constants of the pair enumeration
struct Consts {
    C w1[NT];        // W_8192^tid
    C w2[NT];        // W_4096^(tid & 255)
    C w3[NT];        // W_256^(tid & 15)
    C wt[NT];        // W_N^(32 q3 + 512 q4), tid & 255 = 16 q3 + q4
    C row[2][9];     // W_N^(h + 2 i)
};

__device__ inline void consts_fill(Consts& S, const C* __restrict__ WM, const C* __restrict__ WN, int tid) {
    S.w1[tid] = WM[tid];
    S.w2[tid] = WM[2 * (tid & 255)];
    S.w3[tid] = WM[32 * (tid & 15)];
    const int th = tid & 255;
    S.wt[tid] = WN[32 * (th >> 4) + 512 * (th & 15)];
    if (tid < 18) S.row[tid / 9][tid % 9] = WN[(tid / 9) + 2 * (tid % 9)];
}

// Bulk-asynchronous (TMA) staging of filter tables: a table part is SPAN contiguous entries in global memory, in the block
// arrays' own padded layout, and lands in an array that is dead at that point; the step that applies the table then
// reads it from the very slots it overwrites.  Two mbarriers (RF filter table, first post filter pair), their phase
// parities live across the blocks of a CTA.
struct Stage {
    unsigned long long* bars;      // [2] in shared memory
    unsigned ph_e, ph_f;
};
__device__ inline void stage_init(Stage& st, unsigned long long* bars, int tid) {
    st.bars = bars;
    st.ph_e = st.ph_f = 0u;
#ifndef LDD_EMU
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
#endif
}
// two table parts into two arrays (thread 0 issues; the caller has passed a CTA barrier behind the arrays' last use, with
// stage_fence() in front of it on every thread)
__device__ inline void stage_issue(unsigned long long* bar, C* dst0, const C* src0, C* dst1, const C* src1, int tid) {
#ifdef LDD_EMU
    for (int i = tid; i < SPAN; i += NT) {
        dst0[i] = src0[i];
        if (dst1) dst1[i] = src1[i];
    }
#else
    if (tid == 0) {
        const unsigned bytes = (unsigned)(SPAN * sizeof(C));
        mbar_expect_tx(bar, dst1 ? 2 * bytes : bytes);
        bulk_g2s(dst0, src0, bytes, bar);
        if (dst1) bulk_g2s(dst1, src1, bytes, bar);
    }
#endif
}
__device__ inline void stage_fence() {
#ifndef LDD_EMU
    fence_proxy_async();
#endif
}
__device__ inline void stage_wait(unsigned long long* bar, unsigned& parity) {
#ifndef LDD_EMU
    mbar_wait(bar, parity);
    parity ^= 1u;
#else
    (void)bar; (void)parity;
#endif
}

// One (k, M-k) pair of a spectrum in permuted order: positions p (index k) and pp (index M - k), w = W_N^k.
// kind 0: a proper pair; 1: k = 0 (p = 0, holds the packed X[0], X[M]); 2: k = M/2 (p = 8, pairs with itself).
struct Item {
    int p, pp, kind;
    C w;
};

// Iteration it (0..7) of thread tid.  Half h = tid >> 8 owns positions [4096 h, 4096 h + 4096); the pairs of a half stay
// inside it.  Rows (q2 = it) pair with rows 16 - h - it; in half 0 the rows 0 and 8 pair with themselves and share
// iteration 0 (threads 0-127 of the half: row 8, threads 128-255: row 0).
__device__ inline Item pair_item(const Consts& S, const C* __restrict__ WN, int tid, int it) {
    const int h = tid >> 8, th = tid & 255;
    Item I;
    I.kind = 0;
    if (h == 1 || it >= 1) {
        I.p = (h << 12) + 256 * it + th;
        I.pp = (h << 12) + 256 * (16 - h - it) + 255 - th;
        I.w = S.row[h][it] * S.wt[tid];
    } else if (th < 128) {
        I.p = 2048 + th;
        I.pp = 2048 + 255 - th;
        I.w = S.row[0][8] * S.wt[tid];
    } else {
        const int t2 = th - 128, q3 = t2 >> 3, q4 = t2 & 7;
        I.p = 16 * q3 + q4;
        if (t2 == 0) {
            I.kind = 1;
            I.pp = 0;
            I.w = mk<float>(1.f, 0.f);
        } else {
            I.pp = q3 ? 16 * (16 - q3) + (15 - q4) : (16 - q4);
            I.w = WN[32 * q3 + 512 * q4];
        }
    }
    return I;
}
// the k = M/2 item (one thread of the CTA)
__device__ inline Item self_item() {
    Item I;
    I.p = I.pp = 8;
    I.kind = 2;
    I.w = mk<float>(0.f, -1.f);          // W_N^(N/4)
    return I;
}

// untangle of one pair: za = Z[k], zb = Z[M-k] of the half-length transform -> X[k], X[M-k]
__device__ inline void untangle_pair(const Item& I, C za, C zb, C& xa, C& xb) {
    if (I.kind == 0) {
        const C a = za, b = conj(zb);
        const C E = scale(a + b, 0.5f);
        const C Od = scale(mul_mj(a - b), 0.5f);
        const C Tw = I.w * Od;
        xa = E + Tw;
        xb = conj(E - Tw);
    } else if (I.kind == 1) {
        xa = xb = mk<float>(za.x + za.y, za.x - za.y);
    } else {
        xa = xb = conj(za);
    }
}

// stages 4-2 of the DIT transform of one or two arrays (before: own warp's data complete; after: needs a CTA barrier)
__device__ inline void dit_pair(C* a, C* b, const f2::Tw& tw, int tid) {
    f2::stage4<PK>(a, tid);
    if (b) f2::stage4<PK>(b, tid);
    __syncwarp();
    f2::stage3<PK, true>(a, tw.w3, tid);
    if (b) f2::stage3<PK, true>(b, tw.w3, tid);
    f2::half_sync(tid >> 8);
    f2::stage2<PK, true>(a, tw.w2, tid);
    if (b) f2::stage2<PK, true>(b, tw.w2, tid);
}

// The whole block.  smem: 3 * SPAN complex.  Returns (block-uniform) the mixed lane's flag.  Ends with a barrier.
__device__ inline int demod_block8k(const DemodParams& p, const int blk, char* smem, const Consts& S, const ScanConsts& sc,
                                    const ScanTab& stab, Stage& stg, double* s_warp, double* s_total, unsigned* s_last) {
    const int tid = threadIdx.x, half = tid >> 8;
    C* const b0 = (C*)smem;
    C* const b1 = b0 + SPAN;
    C* const b2 = b1 + SPAN;
    const C* __restrict__ WN = (const C*)p.WN;
    f2::Tw tw;
    tw.w1 = S.w1[tid];
    tw.w2 = S.w2[tid];
    tw.w3 = S.w3[tid];
    const int ix0 = PXi(tid);
    int flagged = 0;
    typedef float T;                 // (the PHASE macro looks at sizeof(T))
    PHASE_BEGIN();

    const long long in0 = p.first_sample + (long long)blk * p.stride;
    const long long o = (long long)blk * p.stride;
    long long copylen = p.stride;
    if (o + (N - p.blockcut) > p.total_out) copylen = p.total_out - o;
    if (copylen > N - p.blockcut) copylen = N - p.blockcut;
    if (copylen < 0) copylen = 0;
    const int keep0 = p.blockcut, keep1 = p.blockcut + (int)copylen;

    const C* __restrict__ HvP = (const C*)p.HvP;
    const C* __restrict__ LnP = (const C*)p.lnMP;
    // the RF filter table of step E: entries k < M into b1, k >= M into b2, under the sample load and the first transform
    // (both arrays are free: the previous block ended with a barrier; with analog audio on they are the audio step's
    // workspace and step E reads the table from global memory)
    const bool pre_e = p.A == 0;
    if (pre_e) stage_issue(&stg.bars[0], b1, HvP, b2, HvP + SPAN, tid);
#ifdef LDD_EMU
    if (pre_e) __syncthreads();
#endif

    // A. samples -> z[n] = x[2n] + j x[2n+1], stage 1 of the forward transform on the way into b0.  A thread takes the
    //    sample groups g = tid + 512 i (four samples = z[2g], z[2g+1]) and their partners 2048 groups on; the arithmetic is
    //    the same whatever the format and alignment (only the fetch differs), so that packed and unpacked captures, and
    //    differently aligned shards of one capture, give bit-equal planes.
    {
        C za[4][2], zb[4][2];            // [i][e]: z[n], z[n + 4096] for n = 2 g + e
        const bool full = in0 + N <= p.rf_limit;
        const unsigned char* rfb = (const unsigned char*)p.rf;
        if (full && p.fmt == LDD_FMT_U8 && ((((uintptr_t)rfb + (uintptr_t)in0) & 3) == 0)) {
            // 8-bit samples: one aligned 32-bit word per group
            const unsigned* r32 = (const unsigned*)(rfb + in0);
            unsigned va[4], vb[4];
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) { va[i] = r32[tid + i * NT]; vb[i] = r32[tid + i * NT + 2048]; }
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                za[i][0] = mk<float>((float)(int)(va[i] & 0xffu), (float)(int)((va[i] >> 8) & 0xffu));
                za[i][1] = mk<float>((float)(int)((va[i] >> 16) & 0xffu), (float)(int)(va[i] >> 24));
                zb[i][0] = mk<float>((float)(int)(vb[i] & 0xffu), (float)(int)((vb[i] >> 8) & 0xffu));
                zb[i][1] = mk<float>((float)(int)((vb[i] >> 16) & 0xffu), (float)(int)(vb[i] >> 24));
            }
        } else if (full && p.fmt == LDD_FMT_LDS40 && (in0 & 3) == 0 && ((((uintptr_t)rfb) & 3) == 0) && in0 + N + 4 <= p.rf_limit) {
            // 4 x 10 bit in 5 bytes (.lds): group g = bytes [5 g, 5 g + 5) of the block, taken from the two aligned words
            // that hold them
            const unsigned char* gb = rfb + (in0 >> 2) * 5;
            const unsigned off0 = (unsigned)((uintptr_t)gb & 3);
            const unsigned* w32 = (const unsigned*)(gb - off0);
            unsigned la[4][2], lb[4][2];
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                const unsigned ba = off0 + 5u * (unsigned)(tid + i * NT), bb = ba + 5u * 2048u;
                la[i][0] = w32[ba >> 2]; la[i][1] = w32[(ba >> 2) + 1];
                lb[i][0] = w32[bb >> 2]; lb[i][1] = w32[(bb >> 2) + 1];
            }
            auto unpack = [](unsigned w0, unsigned w1, unsigned byteoff, C* out2) {
                const unsigned long long v = (((unsigned long long)w1 << 32) | w0) >> (8 * (byteoff & 3));
                const unsigned c0 = (unsigned)v & 0xffu, c1 = (unsigned)(v >> 8) & 0xffu, c2 = (unsigned)(v >> 16) & 0xffu,
                               c3 = (unsigned)(v >> 24) & 0xffu, c4 = (unsigned)(v >> 32) & 0xffu;
                out2[0] = mk<float>((float)(int)((c0 << 2) | (c1 >> 6)), (float)(int)(((c1 & 0x3fu) << 4) | (c2 >> 4)));
                out2[1] = mk<float>((float)(int)(((c2 & 0xfu) << 6) | (c3 >> 2)), (float)(int)(((c3 & 3u) << 8) | c4));
            };
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                const unsigned ba = off0 + 5u * (unsigned)(tid + i * NT);
                unpack(la[i][0], la[i][1], ba, za[i]);
                unpack(lb[i][0], lb[i][1], ba + 5u * 2048u, zb[i]);
            }
        } else if (full) {
            // any other format / alignment: one fetch per sample
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                const long long s1 = in0 + 4 * (tid + i * NT), s2 = s1 + M;
                LDD_UNROLL
                for (int e = 0; e < 2; ++e) {
                    za[i][e] = mk<float>((float)fetch_sample(p.rf, p.fmt, s1 + 2 * e), (float)fetch_sample(p.rf, p.fmt, s1 + 2 * e + 1));
                    zb[i][e] = mk<float>((float)fetch_sample(p.rf, p.fmt, s2 + 2 * e), (float)fetch_sample(p.rf, p.fmt, s2 + 2 * e + 1));
                }
            }
        } else {
            // a block that reaches past the end of the capture reads zeros there
            auto fz = [&](long long sidx) -> float { return sidx < p.rf_limit ? (float)fetch_sample(p.rf, p.fmt, sidx) : 0.f; };
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                const long long s1 = in0 + 4 * (tid + i * NT), s2 = s1 + M;
                LDD_UNROLL
                for (int e = 0; e < 2; ++e) {
                    za[i][e] = mk<float>(fz(s1 + 2 * e), fz(s1 + 2 * e + 1));
                    zb[i][e] = mk<float>(fz(s2 + 2 * e), fz(s2 + 2 * e + 1));
                }
            }
        }
        // stage 1 for n = 2 g + e = (2 tid + e) + 1024 i: W_8192^n = W_8192^(2 tid + e) W_8^i
        const C wa = tw.w1 * tw.w1;
        const C wb = wa * mk<float>(0.99999970586288221916f, -0.00076699031874270453f);       // W_8192^1
        const int ixe = PXi(2 * tid);
        LDD_UNROLL
        for (int i = 0; i < 4; ++i) {
            LDD_UNROLL
            for (int e = 0; e < 2; ++e) {
                C w = e ? wb : wa;
                if (i == 1) w = mk<float>((w.x + w.y) * 0.70710678118654752440f, (w.y - w.x) * 0.70710678118654752440f);      // * W_8
                if (i == 2) w = mul_mj(w);
                if (i == 3) w = mk<float>((w.y - w.x) * 0.70710678118654752440f, -(w.x + w.y) * 0.70710678118654752440f);     // * W_8^3
                const int ix = ixe + e + i * f2::pst<PK>(1024);
                b0[ix] = za[i][e] + zb[i][e];
                b0[ix + P4096] = (za[i][e] - zb[i][e]) * w;
            }
        }
    }
    __syncthreads();
    PHASE(0);

    // B. X = rfft(x): stages 2-4, spectrum in permuted order in b0
    f2::dif_234<PK>(b0, tw, tid);
    if (pre_e) stage_wait(&stg.bars[0], stg.ph_e);
    f2::half_sync(half);
    PHASE(1);

    // per-block MTF level (see demod_block)
    float dl = 0.f;
    if (p.mtf_period > 0.0) {
        const double centre = (double)o + 0.5 * (double)p.stride;
        double dlt = p.mtf_step * floor((centre - p.mtf_pos0) / p.mtf_period);
        if (dlt < -p.mtf_level0) dlt = -p.mtf_level0;
        if (centre < p.mtf_hold_until) dlt = p.mtf_hold_level - p.mtf_level0;
        dl = (float)dlt;
    }
    auto ramp = [&](C hv, C ln) -> C {
        const C z = scale(ln, dl);
        const C z2 = z * z;
        const C e = mk<float>(1.f + z.x, z.y) + scale(z2, 0.5f) + scale(z2 * z, 1.f / 6.f);
        return hv * e;
    };
    // E for one pair: X[k], X[M-k] -> U, V (even / odd output samples of ifft(X_full * Hv), stored conjugated)
    auto e_pair = [&](const Item& I, C xa, C xb, C h0, C h1, C h2, C h3) {
        if (I.kind == 1) {
            const C y0 = scale(h0, xa.x), y1 = scale(h1, xa.y);
            b0[0] = conj(y0 + y1);
            b1[0] = conj(y0 - y1);
        } else {
            const int ip = PXi(I.p);
            const C y0 = xa * h0, y1 = conj(xb) * h1;
            b0[ip] = conj(y0 + y1);
            b1[ip] = conj(mulc(y0 - y1, I.w));
            if (I.kind == 0) {
                const int iq = PXi(I.pp);
                const C z0 = xb * h2, z1 = conj(xa) * h3;
                b0[iq] = conj(z0 + z1);
                const C d = z0 - z1;
                b1[iq] = conj(mk<float>(-d.x, -d.y) * I.w);
            }
        }
    };
    auto e_step = [&](auto rtag, bool fused) {
        constexpr bool RAMP = decltype(rtag)::value;
        LDD_UNROLL
        for (int it0 = 0; it0 < 8; it0 += 4) {
            Item I[4];
            C h0[4], h1[4], h2[4], h3[4], za[4], zb[4];
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                I[i] = pair_item(S, WN, tid, it0 + i);
                const int ip = PXi(I[i].p), iq = PXi(I[i].pp);
                if (fused) {         // staged (pre_e): the table sits in the slots of b1 / b2 this thread is about to overwrite
                    h0[i] = b1[ip]; h1[i] = b2[ip]; h2[i] = b1[iq]; h3[i] = b2[iq];
                } else {
                    h0[i] = HvP[ip]; h1[i] = HvP[SPAN + ip]; h2[i] = HvP[iq]; h3[i] = HvP[SPAN + iq];
                }
                if constexpr (RAMP) {
                    h0[i] = ramp(h0[i], LnP[ip]); h1[i] = ramp(h1[i], LnP[SPAN + ip]);
                    h2[i] = ramp(h2[i], LnP[iq]); h3[i] = ramp(h3[i], LnP[SPAN + iq]);
                }
            }
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) { za[i] = b0[PXi(I[i].p)]; zb[i] = b0[PXi(I[i].pp)]; }
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                C xa = za[i], xb = zb[i];
                if (fused) untangle_pair(I[i], za[i], zb[i], xa, xb);
                e_pair(I[i], xa, xb, h0[i], h1[i], h2[i], h3[i]);
            }
        }
        if (tid == 255) {
            const Item I = self_item();
            C h0 = fused ? b1[PXi(8)] : HvP[PXi(8)], h1 = fused ? b2[PXi(8)] : HvP[SPAN + PXi(8)];
            if constexpr (RAMP) { h0 = ramp(h0, LnP[PXi(8)]); h1 = ramp(h1, LnP[SPAN + PXi(8)]); }
            C xa = b0[PXi(8)], xb = xa;
            if (fused) untangle_pair(I, xa, xa, xa, xb);
            e_pair(I, xa, xb, h0, h1, h0, h1);
        }
    };

    if (p.A > 0) {
        // C. untangle in place, D. analog audio phase 1 on the untangled spectrum (as demod_block; the spectrum is read
        //    through its permuted positions), then E from the stored spectrum
        LDD_UNROLL
        for (int it = 0; it < 8; ++it) {
            const Item I = pair_item(S, WN, tid, it);
            const int ip = PXi(I.p), iq = PXi(I.pp);
            C xa, xb;
            untangle_pair(I, b0[ip], b0[iq], xa, xb);
            b0[ip] = xa;
            if (I.kind == 0) b0[iq] = xb;
        }
        if (tid == 255) b0[PXi(8)] = conj(b0[PXi(8)]);
        __syncthreads();
        PHASE(2);
        {
            const int A = p.A, hA = A / 2, nthr = NT;
            const C* AL = (const C*)p.AL;
            const C* AR = (const C*)p.AR;
            const C* WM = (const C*)p.WM;
            C* gl = b1;
            C* gr = b2;
            for (int j = tid; j < A; j += nthr) {
                C xa = (j < hA) ? b0[PXi(f2::pos_of_idx(p.a_lo + j))] : conj(b0[PXi(f2::pos_of_idx(p.a_hi - (j - hA)))]);
                gl[pidx<true>(j)] = conj(xa * AL[j]);
                gr[pidx<true>(j)] = conj(xa * AR[j]);
            }
            __syncthreads();
            // the two channels' transforms side by side: warps 0-7 take the left one, warps 8-15 the right one, radix 8 so
            // that all 256 threads of a half have a butterfly in every pass, 256-thread barriers between the passes
            C* rl;
            C* rr;
            {
                C* src = half ? gr : gl;
                C* dst = src + pspan<true>(A);
                int Ns = 1;
                while (Ns < A) {
                    const int R = (A / Ns) >= 8 ? 8 : (A / Ns);
                    fft_pass_any<float, true, true>(R, src, dst, A, Ns, WM, p.wstride_a, tid & 255, 256);
                    Ns *= R;
                    f2::half_sync(half);
                    C* t = src; src = dst; dst = t;
                }
                // (both halves ran the same number of passes: the results sit at the same offset of their arrays)
                rl = src - (half ? gr - gl : 0);
                rr = rl + (gr - gl);
            }
            __syncthreads();
            C* ang = (rl == gl) ? gl + pspan<true>(A) : gl;
            for (int j = tid; j < A; j += nthr) {
                C l = rl[pidx<true>(j)], r = rr[pidx<true>(j)];
                ang[pidx<true>(j)] = mk<float>(Math<float>::atan2(-l.y, l.x), Math<float>::atan2(-r.y, r.x));
            }
            __syncthreads();
            const int a0 = keep0 / p.audio_ds, a1 = keep1 / p.audio_ds;
            const long long ao = o / p.audio_ds;
            const double twopi = 6.283185307179586476925286766559;
            for (int j = a0 + tid; j < a1; j += nthr) {
                double d_l = 0.0, d_r = 0.0;
                if (j > 0) {
                    C c1 = ang[pidx<true>(j)], c0 = ang[pidx<true>(j - 1)];
                    d_l = (double)c1.x - (double)c0.x;
                    d_r = (double)c1.y - (double)c0.y;
                    if (d_l < 0) d_l += twopi;
                    if (d_r < 0) d_r += twopi;
                }
                long long oi = ao + (j - a0);
                if (oi < p.audio_total) {
                    p.audio_l[oi] = d_l * p.audio_scale + p.audio_lowfreq;
                    p.audio_r[oi] = d_r * p.audio_scale + p.audio_lowfreq;
                }
            }
            __syncthreads();
        }
        PHASE(3);
        if (dl != 0.f) e_step(TrueTag{}, false);
        else e_step(FalseTag{}, false);
    } else {
        // C+E fused: the thread that untangles a pair filters it
        if (dl != 0.f) e_step(TrueTag{}, true);
        else e_step(FalseTag{}, true);
    }
    f2::half_sync(half);
    PHASE(4);

    // F. the two inverse transforms (U in b0, V in b1), stages 4-2
    dit_pair(b0, b1, tw, tid);
    __syncthreads();
    PHASE(5);

    // G. stage 1 of both + FM discriminator angles (lddutils.py:320-334): angles of (u, v) pairs in place in b0
    {
        LDD_UNROLL
        for (int i0 = 0; i0 < 8; i0 += 2) {
            C ua[2], ub[2], va[2], vb[2];
            LDD_UNROLL
            for (int i = 0; i < 2; ++i) {
                const int ix = ix0 + (i0 + i) * P512;
                ua[i] = b0[ix]; ub[i] = b0[ix + P4096]; va[i] = b1[ix]; vb[i] = b1[ix + P4096];
            }
            LDD_UNROLL
            for (int i = 0; i < 2; ++i) {
                const int ix = ix0 + (i0 + i) * P512;
                const C w = (i0 + i) == 0 ? tw.w1 : tw.w1 * w16<float>(i0 + i);
                const C tu = ub[i] * w, tv = vb[i] * w;
                const C r1 = ua[i] + tu, r2 = ua[i] - tu, q1 = va[i] + tv, q2 = va[i] - tv;
                b0[ix] = mk<float>(Math<float>::atan2(-r1.y, r1.x), Math<float>::atan2(-q1.y, q1.x));
                b0[ix + P4096] = mk<float>(Math<float>::atan2(-r2.y, r2.x), Math<float>::atan2(-q2.y, q2.x));
            }
        }
    }
    __syncthreads();
    PHASE(6);
    // neighbour difference, fold to [0, 2 pi), Hz, minus ire0; packed as the next real transform's input, whose stage 1
    // runs on the way into b1
    {
        const float twopi = 6.283185307179586476925286766559f;
        const float hz = (float)p.hz_per_rad, ire0 = (float)p.ire0;
        const int ixm0 = PXi(tid - 1);          // arithmetic shift: also right for tid = 0, i >= 1
        C x1[8], x2[8];
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int ix = ix0 + i * P512, ixm = ixm0 + i * P512;
            const C a = b0[ix], a2 = b0[ix + P4096];
            float d0 = 0.f;
            if (tid + i > 0) {
                d0 = a.x - b0[ixm].y;
                if (d0 < 0) d0 += twopi;
            }
            float d1 = a.y - a.x;
            if (d1 < 0) d1 += twopi;
            x1[i] = mk<float>(d0 * hz - ire0, d1 * hz - ire0);
            float e0 = a2.x - b0[ixm + P4096].y;
            if (e0 < 0) e0 += twopi;
            float e1 = a2.y - a2.x;
            if (e1 < 0) e1 += twopi;
            x2[i] = mk<float>(e0 * hz - ire0, e1 * hz - ire0);
        }
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int ix = ix0 + i * P512;
            const C w = i == 0 ? tw.w1 : tw.w1 * w16<float>(i);
            b1[ix] = x1[i] + x2[i];
            b1[ix + P4096] = (x1[i] - x2[i]) * w;
        }
    }
    stage_fence();
    __syncthreads();
    PHASE(7);
    // the first post filter pair's tables into the arrays they will be applied in (b0: video, b2: burst; both dead now),
    // under the next transform
    const bool pre_f = !p.only05;
    if (pre_f) stage_issue(&stg.bars[1], b0, (const C*)p.FP[0], b2, (const C*)p.FP[2], tid);
#ifdef LDD_EMU
    if (pre_f) __syncthreads();
#endif

    // H. D = rfft(demod - ire0): stages 2-4 in b1, untangle in place (the same thread owns a pair here and in the
    //    post-filter step, so no barrier is needed between them)
    f2::dif_234<PK>(b1, tw, tid);
    if (pre_f) stage_wait(&stg.bars[1], stg.ph_f);
    f2::half_sync(half);
    PHASE(8);
    LDD_UNROLL
    for (int it0 = 0; it0 < 8; it0 += 4) {
        Item I[4];
        C za[4], zb[4];
        LDD_UNROLL
        for (int i = 0; i < 4; ++i) {
            I[i] = pair_item(S, WN, tid, it0 + i);
            za[i] = b1[PXi(I[i].p)];
            zb[i] = b1[PXi(I[i].pp)];
        }
        LDD_UNROLL
        for (int i = 0; i < 4; ++i) {
            C xa, xb;
            untangle_pair(I[i], za[i], zb[i], xa, xb);
            b1[PXi(I[i].p)] = xa;
            if (I[i].kind == 0) b1[PXi(I[i].pp)] = xb;
        }
    }
    if (tid == 255) b1[PXi(8)] = conj(b1[PXi(8)]);
    PHASE(9);

    // I. post filters in pairs: (video, burst) then (pilot, video05) / (video05); video05 always lands in b2
    // (NTSC: video, video05, burst; PAL: + pilot.  Table and plane indices are compile-time constants per set.)
    LDD_UNROLL
    for (int set = 0; set < 2; ++set) {
        if (set == 0 && p.only05) continue;
        // filters of this set: ma -> b0 (may be absent), mb -> b2
        const int ma = set == 0 ? 0 : 3, mb = set == 0 ? 2 : 1;
        const bool have_a = set == 0 ? true : (p.nfilt > 3 && !p.only05);
        const C* __restrict__ Fa = have_a ? (const C*)(set == 0 ? p.FP[0] : p.FP[3]) : nullptr;
        const C* __restrict__ Fb = (const C*)(set == 0 ? p.FP[2] : p.FP[1]);
        // tangle: Y = D * F (conj-symmetric) -> conj(Q), whose forward transform gives the real signal
        auto tangle_pair = [&](const Item& I, C da, C db, C fa, C fb, C* Q) {
            if (I.kind == 1) {
                const float y0 = da.x * fa.x, ym = da.y * fb.x;           // fb: F[M]
                Q[0] = mk<float>((y0 + ym) * 0.5f, -(y0 - ym) * 0.5f);
            } else if (I.kind == 2) {
                Q[PXi(8)] = da * fa;
            } else {
                const C a = da * fa, b = conj(db * fb);
                const C E = scale(a + b, 0.5f);
                const C Od = mulc(scale(a - b, 0.5f), I.w);
                const C q = E + mul_pj(Od);
                const C qm = conj(E) + mul_pj(conj(Od));
                Q[PXi(I.p)] = conj(q);
                Q[PXi(I.pp)] = conj(qm);
            }
        };
        LDD_UNROLL
        for (int it0 = 0; it0 < 8; it0 += 4) {
            Item I[4];
            C fa1[4], fb1[4], fa2[4], fb2[4], da[4], db[4];
            const bool staged = set == 0 && pre_f;        // the tables sit in the slots of b2 / b0 this thread overwrites below
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                I[i] = pair_item(S, WN, tid, it0 + i);
                const int ip = PXi(I[i].p), iq = PXi(I[i].pp);
                if (staged) {
                    fa2[i] = b2[ip]; fb2[i] = b2[iq];
                    if (Fa) { fa1[i] = b0[ip]; fb1[i] = b0[iq]; }
                } else {
                    fa2[i] = Fb[ip]; fb2[i] = Fb[iq];
                    if (Fa) { fa1[i] = Fa[ip]; fb1[i] = Fa[iq]; }
                }
                if (I[i].kind == 1) {                      // k = 0 pairs with F[M], kept behind the permuted table
                    fb2[i] = Fb[SPAN];
                    if (Fa) fb1[i] = Fa[SPAN];
                }
            }
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) { da[i] = b1[PXi(I[i].p)]; db[i] = b1[PXi(I[i].pp)]; }
            LDD_UNROLL
            for (int i = 0; i < 4; ++i) {
                tangle_pair(I[i], da[i], db[i], fa2[i], fb2[i], b2);
                if (Fa) tangle_pair(I[i], da[i], db[i], fa1[i], fb1[i], b0);
            }
        }
        if (tid == 255) {
            const Item I = self_item();
            const C d = b1[PXi(8)];
            const bool staged = set == 0 && pre_f;
            const C f2v = staged ? b2[PXi(8)] : Fb[PXi(8)], f1v = Fa ? (staged ? b0[PXi(8)] : Fa[PXi(8)]) : f2v;
            tangle_pair(I, d, d, f2v, f2v, b2);
            if (Fa) tangle_pair(I, d, d, f1v, f1v, b0);
        }
        f2::half_sync(half);
        PHASE(10);
        dit_pair(b2, Fa ? b0 : nullptr, tw, tid);
        __syncthreads();
        PHASE(11);
        // stage 1 + store of the kept samples: y[2n] = r.x, y[2n+1] = -r.y (+ the filter's DC gain times ire0)
        LDD_UNROLL
        for (int which = 0; which < 2; ++which) {
            if (which == 1 && !have_a) continue;
            const int m = which == 0 ? mb : ma;
            C* g = which == 0 ? b2 : b0;
            static_assert(LDD_P_DEMOD == 0 && LDD_P_DEMOD05 == 1 && LDD_P_BURST == 3 && LDD_P_PILOT == 4, "plane order");
            float* out = (float*)(which == 0 ? (set == 0 ? p.plane[3] : p.plane[1]) : (set == 0 ? p.plane[0] : p.plane[4]));
            const float addc = (float)(which == 0 ? (set == 0 ? p.addc[2] : p.addc[1]) : (set == 0 ? p.addc[0] : p.addc[3]));
            const bool fast = ((keep0 | keep1) & 1) == 0 && ((o & 1) == 0) && ((((uintptr_t)out) & 7) == 0);
            float* ob = out + o - keep0;
            LDD_UNROLL
            for (int i0 = 0; i0 < 8; i0 += 4) {
                C a[4], b[4];
                LDD_UNROLL
                for (int i = 0; i < 4; ++i) { a[i] = g[ix0 + (i0 + i) * P512]; b[i] = g[ix0 + (i0 + i) * P512 + P4096]; }
                LDD_UNROLL
                for (int i = 0; i < 4; ++i) {
                    const C w = (i0 + i) == 0 ? tw.w1 : tw.w1 * w16<float>(i0 + i);
                    const C t = b[i] * w;
                    const C r1 = a[i] + t, r2 = a[i] - t;
                    const int s1 = 2 * (tid + (i0 + i) * NT), s2 = s1 + M;
                    if (fast) {
                        if (s1 >= keep0 && s1 < keep1) st_stream((float2*)(ob + s1), make_float2(r1.x + addc, -r1.y + addc));
                        if (s2 >= keep0 && s2 < keep1) st_stream((float2*)(ob + s2), make_float2(r2.x + addc, -r2.y + addc));
                    } else {
                        if (s1 >= keep0 && s1 < keep1) st_stream(ob + s1, r1.x + addc);
                        if (s1 + 1 >= keep0 && s1 + 1 < keep1) st_stream(ob + s1 + 1, -r1.y + addc);
                        if (s2 >= keep0 && s2 < keep1) st_stream(ob + s2, r2.x + addc);
                        if (s2 + 1 >= keep0 && s2 + 1 < keep1) st_stream(ob + s2 + 1, -r2.y + addc);
                    }
                    if (m == 1) {
                        // the whole filtered block feeds the sync scan: natural order, in place
                        g[ix0 + (i0 + i) * P512] = r1;
                        g[ix0 + (i0 + i) * P512 + P4096] = r2;
                    }
                }
            }
        }
        __syncthreads();
        PHASE(12);
    }

    // J. sync: s[n] = lo <= demod_05[n] <= hi (lddecode_core.py:308), decided on the float32 samples against float32
    //    thresholds (a sample within flag_margin of a threshold sends the block to the float64 lane, whose decision
    //    stands); recursion and store shared with every other lane
    {
        const double add = p.addc[1] + p.sync_ref;
        // |x - mid| <= hw  <=>  lo <= x <= hi;  | |x - mid| - hw | < margin  <=>  x within the margin of a threshold
        const float mid = (float)(0.5 * (p.sync_lo + p.sync_hi) - add), hw = (float)(0.5 * (p.sync_hi - p.sync_lo)), mg = (float)p.flag_margin;
        const C* e = b2 + (N / NT / 2 + PK) * tid;          // 16 elements = this thread's 32 samples
        unsigned mask = 0u;
        int near = 0;
        C v[16];
        LDD_UNROLL
        for (int j = 0; j < 16; ++j) v[j] = e[j];
        LDD_UNROLL
        for (int j = 0; j < 16; ++j) {
            const float t0 = fabsf(v[j].x - mid), t1 = fabsf(-v[j].y - mid);
            if (t0 <= hw) mask |= (1u << (2 * j));
            if (t1 <= hw) mask |= (2u << (2 * j));
            near |= (fabsf(t0 - hw) < mg) | (fabsf(t1 - hw) < mg);
        }
        if (p.flag_margin > 0.0) flagged = __syncthreads_or(near | (in0 + N > p.rf_limit));
        PHASE(13);
        sync_scan32(p, sc, stab, mask, (double*)b0, keep0, keep1, o, s_warp, s_total, s_last);
    }
    stage_fence();
    __syncthreads();
    PHASE(15);
    return flagged;
}

// ---- the mixed lane's float64 re-run of a flagged block (demod_05 -> demod_sync only) on the same in-place transforms --
// One float64 array (136 KB) lives in shared memory and every transform runs in place in it; the second sequence of the
// analytic-signal step (V) waits in this CTA's slice of the L2-resident scratch and the angles of the first (U) in a
// 64 KB shared array.  The generic float64 block (demod_block<double>) moves every second transform pass through L2; this
// one moves V out and back once: a re-run takes about a third of the time, and the tail it adds to the launch shrinks
// with it.  Arithmetic differs from the Stockham plan only in rounding (~1e-16 relative): a sync decision could differ
// from the exact lane's only for a sample within ~1e-9 Hz of a threshold.
typedef Cx<double> C2;
struct Consts64 {
    C2 w1[NT];       // W_8192^tid
    C2 w2[256];      // W_4096^j
    C2 wt[256];      // W_N^(32 q3 + 512 q4), j = 16 q3 + q4
    C2 w3[16];       // W_256^j
    C2 row[2][9];    // W_N^(h + 2 i)
};

__device__ inline void consts64_fill(Consts64& S, const C2* __restrict__ WM, const C2* __restrict__ WN, int tid) {
    S.w1[tid] = WM[tid];
    if (tid < 256) {
        S.w2[tid] = WM[2 * tid];
        S.wt[tid] = WN[32 * (tid >> 4) + 512 * (tid & 15)];
    }
    if (tid < 16) S.w3[tid] = WM[32 * tid];
    if (tid < 18) S.row[tid / 9][tid % 9] = WN[(tid / 9) + 2 * (tid % 9)];
}

struct Item64 {
    int p, pp, kind;
    C2 w;
};
__device__ inline Item64 pair_item64(const Consts64& S, const C2* __restrict__ WN, int tid, int it) {
    const int h = tid >> 8, th = tid & 255;
    Item64 I;
    I.kind = 0;
    if (h == 1 || it >= 1) {
        I.p = (h << 12) + 256 * it + th;
        I.pp = (h << 12) + 256 * (16 - h - it) + 255 - th;
        I.w = S.row[h][it] * S.wt[th];
    } else if (th < 128) {
        I.p = 2048 + th;
        I.pp = 2048 + 255 - th;
        I.w = S.row[0][8] * S.wt[th];
    } else {
        const int t2 = th - 128, q3 = t2 >> 3, q4 = t2 & 7;
        I.p = 16 * q3 + q4;
        if (t2 == 0) {
            I.kind = 1;
            I.pp = 0;
            I.w = mk<double>(1.0, 0.0);
        } else {
            I.pp = q3 ? 16 * (16 - q3) + (15 - q4) : (16 - q4);
            I.w = WN[32 * q3 + 512 * q4];
        }
    }
    return I;
}
__device__ inline void untangle_pair64(const Item64& I, C2 za, C2 zb, C2& xa, C2& xb) {
    if (I.kind == 0) {
        const C2 a = za, b = conj(zb);
        const C2 E = scale(a + b, 0.5);
        const C2 Od = scale(mul_mj(a - b), 0.5);
        const C2 Tw = I.w * Od;
        xa = E + Tw;
        xb = conj(E - Tw);
    } else if (I.kind == 1) {
        xa = xb = mk<double>(za.x + za.y, za.x - za.y);
    } else {
        xa = xb = conj(za);
    }
}

// smem: the kernel's dynamic shared memory (>= SPAN * 16 + M * 8 bytes); vscratch: M float64 complex in global memory;
// cb: shared-memory room for the float64 constants (the caller re-fills its own constants afterwards).  p: the float64
// parameter set (only05).  Ends with a barrier.
__device__ inline void rerun8k(const DemodParams& p, const int blk, char* smem, C2* __restrict__ vscratch, Consts64& cb, const ScanConsts& sc,
                               const ScanTab& stab, double* s_warp, double* s_total, unsigned* s_last) {
    const int tid = threadIdx.x, half = tid >> 8;
    C2* const Sx = (C2*)smem;
    double* const A1 = (double*)(smem + (size_t)SPAN * sizeof(C2));
    const C2* __restrict__ WN = (const C2*)p.WN;
    consts64_fill(cb, (const C2*)p.WM, WN, tid);
    __syncthreads();
    f2::TwT<double> tw;
    tw.w1 = cb.w1[tid];
    tw.w2 = cb.w2[tid & 255];
    tw.w3 = cb.w3[tid & 15];
    const int ix0 = PXi(tid);
    typedef double T;
    PHASE_BEGIN();

    const long long in0 = p.first_sample + (long long)blk * p.stride;
    const long long o = (long long)blk * p.stride;
    long long copylen = p.stride;
    if (o + (N - p.blockcut) > p.total_out) copylen = p.total_out - o;
    if (copylen > N - p.blockcut) copylen = N - p.blockcut;
    if (copylen < 0) copylen = 0;
    const int keep0 = p.blockcut, keep1 = p.blockcut + (int)copylen;

    // A. samples, stage 1
    {
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            const long long sa = in0 + 2 * (tid + i * NT), sb = sa + M;
            const int a0 = sa < p.rf_limit ? fetch_sample(p.rf, p.fmt, sa) : 0, a1 = sa + 1 < p.rf_limit ? fetch_sample(p.rf, p.fmt, sa + 1) : 0;
            const int c0 = sb < p.rf_limit ? fetch_sample(p.rf, p.fmt, sb) : 0, c1 = sb + 1 < p.rf_limit ? fetch_sample(p.rf, p.fmt, sb + 1) : 0;
            const C2 za = mk<double>((double)a0, (double)a1), zb = mk<double>((double)c0, (double)c1);
            const C2 w = i == 0 ? tw.w1 : tw.w1 * w16<double>(i);
            Sx[ix0 + i * P512] = za + zb;
            Sx[ix0 + i * P512 + P4096] = (za - zb) * w;
        }
    }
    __syncthreads();
    PHASE(0);
    f2::dif_234<PK>(Sx, tw, tid);
    f2::half_sync(half);
    PHASE(1);

    // C+E: untangle, RF filter, even/odd split: U in place, V to the scratch (by position)
    double dl = 0.0;
    if (p.mtf_period > 0.0) {
        const double centre = (double)o + 0.5 * (double)p.stride;
        double dlt = p.mtf_step * floor((centre - p.mtf_pos0) / p.mtf_period);
        if (dlt < -p.mtf_level0) dlt = -p.mtf_level0;
        if (centre < p.mtf_hold_until) dlt = p.mtf_hold_level - p.mtf_level0;
        dl = dlt;
    }
    const C2* __restrict__ HvP = (const C2*)p.HvP;
    const C2* __restrict__ LnP = (const C2*)p.lnMP;
    auto hv = [&](int pos) -> C2 {
        C2 h = HvP[pos];
        if (dl != 0.0) {
            const C2 z = scale(LnP[pos], dl);
            const C2 z2 = z * z;
            const C2 e = mk<double>(1.0 + z.x, z.y) + scale(z2, 0.5) + scale(z2 * z, 1.0 / 6.0);
            h = h * e;
        }
        return h;
    };
    auto e_pair = [&](const Item64& I, C2 xa, C2 xb, C2 h0, C2 h1, C2 h2, C2 h3) {
        if (I.kind == 1) {
            const C2 y0 = scale(h0, xa.x), y1 = scale(h1, xa.y);
            Sx[0] = conj(y0 + y1);
            vscratch[0] = conj(y0 - y1);
        } else {
            const C2 y0 = xa * h0, y1 = conj(xb) * h1;
            Sx[PXi(I.p)] = conj(y0 + y1);
            vscratch[I.p] = conj(mulc(y0 - y1, I.w));
            if (I.kind == 0) {
                const C2 z0 = xb * h2, z1 = conj(xa) * h3;
                Sx[PXi(I.pp)] = conj(z0 + z1);
                const C2 d = z0 - z1;
                vscratch[I.pp] = conj(mk<double>(-d.x, -d.y) * I.w);
            }
        }
    };
    LDD_UNROLL
    for (int it0 = 0; it0 < 8; it0 += 2) {
        Item64 I[2];
        C2 h0[2], h1[2], h2[2], h3[2];
        LDD_UNROLL
        for (int i = 0; i < 2; ++i) {
            I[i] = pair_item64(cb, WN, tid, it0 + i);
            h0[i] = hv(I[i].p); h1[i] = hv(M + I[i].p); h2[i] = hv(I[i].pp); h3[i] = hv(M + I[i].pp);
        }
        LDD_UNROLL
        for (int i = 0; i < 2; ++i) {
            C2 xa, xb;
            untangle_pair64(I[i], Sx[PXi(I[i].p)], Sx[PXi(I[i].pp)], xa, xb);
            e_pair(I[i], xa, xb, h0[i], h1[i], h2[i], h3[i]);
        }
    }
    if (tid == 255) {
        Item64 I;
        I.p = I.pp = 8; I.kind = 2; I.w = mk<double>(0.0, -1.0);
        C2 xa, xb;
        untangle_pair64(I, Sx[PXi(8)], Sx[PXi(8)], xa, xb);
        const C2 h0 = hv(8), h1 = hv(M + 8);
        e_pair(I, xa, xb, h0, h1, h0, h1);
    }
    f2::half_sync(half);
    PHASE(4);

    // F/G. inverse transform of U, its angles to A1; then V comes in from the scratch and gets the same
    f2::dit_432<PK>(Sx, tw, tid);
    __syncthreads();
    LDD_UNROLL
    for (int i = 0; i < 8; ++i) {
        const int ix = ix0 + i * P512;
        const C2 a = Sx[ix], b = Sx[ix + P4096];
        const C2 w = i == 0 ? tw.w1 : tw.w1 * w16<double>(i);
        const C2 t = b * w;
        const C2 r1 = a + t, r2 = a - t;
        A1[tid + i * NT] = Math<double>::atan2(-r1.y, r1.x);
        A1[tid + i * NT + M / 2] = Math<double>::atan2(-r2.y, r2.x);
    }
    __syncthreads();
    LDD_UNROLL
    for (int i = 0; i < 16; ++i) Sx[ix0 + i * P512] = vscratch[tid + i * NT];
    __syncthreads();
    f2::dit_432<PK>(Sx, tw, tid);
    __syncthreads();
    PHASE(5);
    LDD_UNROLL
    for (int i = 0; i < 8; ++i) {
        const int ix = ix0 + i * P512;
        const C2 a = Sx[ix], b = Sx[ix + P4096];
        const C2 w = i == 0 ? tw.w1 : tw.w1 * w16<double>(i);
        const C2 t = b * w;
        const C2 r1 = a + t, r2 = a - t;
        Sx[ix].x = Math<double>::atan2(-r1.y, r1.x);
        Sx[ix + P4096].x = Math<double>::atan2(-r2.y, r2.x);
    }
    __syncthreads();
    PHASE(6);
    // neighbour difference -> Hz - ire0, packed; all reads before the barrier, stage 1 of the next transform after it
    {
        const double twopi = 6.283185307179586476925286766559;
        const double hz = p.hz_per_rad, ire0 = p.ire0;
        const int ixm0 = PXi(tid - 1);
        C2 x1[8], x2[8];
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int ix = ix0 + i * P512, ixm = ixm0 + i * P512, n = tid + i * NT;
            const double au = A1[n], av = Sx[ix].x, au2 = A1[n + M / 2], av2 = Sx[ix + P4096].x;
            double d0 = 0.0;
            if (n > 0) {
                d0 = au - Sx[ixm].x;
                if (d0 < 0) d0 += twopi;
            }
            double d1 = av - au;
            if (d1 < 0) d1 += twopi;
            x1[i] = mk<double>(d0 * hz - ire0, d1 * hz - ire0);
            double e0 = au2 - Sx[ixm + P4096].x;
            if (e0 < 0) e0 += twopi;
            double e1 = av2 - au2;
            if (e1 < 0) e1 += twopi;
            x2[i] = mk<double>(e0 * hz - ire0, e1 * hz - ire0);
        }
        __syncthreads();
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int ix = ix0 + i * P512;
            const C2 w = i == 0 ? tw.w1 : tw.w1 * w16<double>(i);
            Sx[ix] = x1[i] + x2[i];
            Sx[ix + P4096] = (x1[i] - x2[i]) * w;
        }
    }
    __syncthreads();
    PHASE(7);

    // H/I. D = rfft(demod - ire0); untangle and FVideo05 in one step, in place; inverse transform
    f2::dif_234<PK>(Sx, tw, tid);
    f2::half_sync(half);
    PHASE(8);
    {
        const C2* __restrict__ F = (const C2*)p.FP[1];
        LDD_UNROLL
        for (int it0 = 0; it0 < 8; it0 += 2) {
            Item64 I[2];
            C2 fa[2], fb[2];
            LDD_UNROLL
            for (int i = 0; i < 2; ++i) {
                I[i] = pair_item64(cb, WN, tid, it0 + i);
                fa[i] = F[I[i].p];
                fb[i] = F[I[i].kind == 1 ? M : I[i].pp];
            }
            LDD_UNROLL
            for (int i = 0; i < 2; ++i) {
                C2 da, db;
                untangle_pair64(I[i], Sx[PXi(I[i].p)], Sx[PXi(I[i].pp)], da, db);
                if (I[i].kind == 1) {
                    const double y0 = da.x * fa[i].x, ym = da.y * fb[i].x;
                    Sx[0] = mk<double>((y0 + ym) * 0.5, -(y0 - ym) * 0.5);
                } else {
                    const C2 a = da * fa[i], b = conj(db * fb[i]);
                    const C2 E = scale(a + b, 0.5);
                    const C2 Od = mulc(scale(a - b, 0.5), I[i].w);
                    const C2 q = E + mul_pj(Od);
                    const C2 qm = conj(E) + mul_pj(conj(Od));
                    Sx[PXi(I[i].p)] = conj(q);
                    Sx[PXi(I[i].pp)] = conj(qm);
                }
            }
        }
        if (tid == 255) Sx[PXi(8)] = conj(Sx[PXi(8)]) * F[8];
    }
    f2::half_sync(half);
    PHASE(10);
    f2::dit_432<PK>(Sx, tw, tid);
    __syncthreads();
    PHASE(11);
    {
        float* out = (float*)p.plane[LDD_P_DEMOD05];
        const double addc = p.addc[1];
        float* ob = out + o - keep0;
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            const int ix = ix0 + i * P512;
            const C2 a = Sx[ix], b = Sx[ix + P4096];
            const C2 w = i == 0 ? tw.w1 : tw.w1 * w16<double>(i);
            const C2 t = b * w;
            const C2 r1 = a + t, r2 = a - t;
            const int s1 = 2 * (tid + i * NT), s2 = s1 + M;
            if (s1 >= keep0 && s1 < keep1) st_stream(ob + s1, (float)(r1.x + addc));
            if (s1 + 1 >= keep0 && s1 + 1 < keep1) st_stream(ob + s1 + 1, (float)(-r1.y + addc));
            if (s2 >= keep0 && s2 < keep1) st_stream(ob + s2, (float)(r2.x + addc));
            if (s2 + 1 >= keep0 && s2 + 1 < keep1) st_stream(ob + s2 + 1, (float)(-r2.y + addc));
            Sx[ix] = r1;
            Sx[ix + P4096] = r2;
        }
    }
    __syncthreads();
    PHASE(12);
    // J. decisions in float64 exactly as the generic block takes them, then the shared recursion
    {
        const double add = p.addc[1] + p.sync_ref;
        const C2* e = Sx + (N / NT / 2 + PK) * tid;
        unsigned mask = 0u;
        LDD_UNROLL
        for (int j = 0; j < 16; ++j) {
            const C2 v = e[j];
            const double v0 = v.x + add, v1 = -v.y + add;
            if (v0 >= p.sync_lo && v0 <= p.sync_hi) mask |= (1u << (2 * j));
            if (v1 >= p.sync_lo && v1 <= p.sync_hi) mask |= (2u << (2 * j));
        }
        PHASE(13);
        sync_scan32(p, sc, stab, mask, (double*)Sx, keep0, keep1, o, s_warp, s_total, s_last);
    }
    __syncthreads();
    PHASE(15);
}

}  // namespace d8

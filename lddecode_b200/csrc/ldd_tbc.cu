// Kernel (5): per-line time-base correction.  Restates Field.downscale + lddutils.scale +
// the final quantisation of FieldNTSC/FieldPAL.downscale (lddecode_core.py:789-812, 1023-1035,
// 1135-1159; lddutils.py:83-97): every output line is a cubic interpolating spline with
// not-a-knot ends (FITPACK splrep s=0,k=3) through the input samples int(b)..int(e) of the line,
// evaluated at outwidth equidistant points, times the line's wow factor, then scaled to uint16.
//
// The spline is solved in closed form instead of by a sequential tridiagonal sweep.  On unit
// spacing its second derivatives satisfy M[i-1] + 4 M[i] + M[i+1] = 6 (y[i-1] - 2 y[i] + y[i+1]).
// The inverse of the infinite (1,4,1) operator is g[k] = r^|k| / (2 sqrt 3), r = sqrt 3 - 2, so a
// particular solution P is a short FIR over the line's samples and a halo of real neighbours
// (|r|^24 = 2e-14, far below the float32 resolution of the input plane), computed by all threads in parallel; the two not-a-knot rows then fix the
// homogeneous part alpha r^i + beta r^(n-i) by a 2x2 solve.  One CTA per output line.
#include "ldd_internal.h"

namespace ldd {

constexpr int TBC_K = 24;                 // reach of the Green's function FIR: |r|^24 = 2e-14 (input is float32)
constexpr int TBC_H = TBC_K + 1;          // halo samples needed on each side
constexpr int TBC_MAXD = 4032;            // longest input line span supported
constexpr int TBC_THREADS = 256;
constexpr int TBC_NTAPS = 2 * TBC_H + 1;  // taps of the y -> P filter
constexpr int TBC_OPT = 9;                // outputs per thread and sweep of the register-tiled FIR (2271/9 < 256 threads)

// w[m] = 6 (g[m-1] - 2 g[m] + g[m+1]), g[k] = r^|k| / (2 sqrt 3) truncated to |k| <= K: set once per
// process.  Constant memory lets the fully unrolled FIR take its taps as immediate constant-bank operands.
__constant__ double c_tbc_taps[TBC_NTAPS];

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
    const float* burstlevel;  // NTSC final: [nfields][ll_stride] or NULL
    float clevel_k;           // float32(327.67 * clevel)
    int* status;              // [nfields]: OR of per-line error bits (1: window outside the plane / too long)
};

__global__ void __launch_bounds__(TBC_THREADS) tbc_kernel(const TbcParams p) {
    LDD_DYN_SMEM(smem_raw);
    double* ys = (double*)smem_raw;                       // y[-H .. dist+H]
    double* Ms = ys + (TBC_MAXD + 2 * TBC_H + TBC_OPT + 2);   // M[0 .. dist]
    __shared__ double s_ab[2];

    const int tid = threadIdx.x;
    const int field = blockIdx.y, line = blockIdx.x;
    const int linecount = p.linecount[field];
    if (line >= linecount) return;
    const double* ll = p.linelocs + (size_t)field * p.ll_stride;
    const double b = ll[p.lineoffset + line] + p.lineloc_add, e = ll[p.lineoffset + line + 1] + p.lineloc_add;
    const long long ib = (long long)b, ie = (long long)e;
    const int dist = (int)(ie - ib);
    const int W = p.outwidth;
    char* outbase = (char*)p.out;
    const long long base = p.base ? p.base[field] : 0;
    if (!(b >= 0.0) || dist < 3 || dist > TBC_MAXD || base + ib + dist + 1 > p.n || base + ib < 0) {
        if (tid == 0) atomicOr(&p.status[field], 1);
        return;
    }
    const double r = -0.26794919243112270647;      // sqrt(3) - 2
    const double c = 0.28867513459481288225;       // 1 / (2 sqrt 3)
    // stage the samples (relative values; the spline is linear so plane_add is added at the end)
    for (int i = tid; i < dist + 1 + 2 * TBC_H + TBC_OPT; i += TBC_THREADS) {
        long long s = base + ib - TBC_H + i;
        s = s < 0 ? 0 : (s >= p.n ? p.n - 1 : s);
        ys[i] = (double)p.plane[s];
    }
    __syncthreads();
    // P = w * y, register tiled: one shared-memory load feeds up to TBC_OPT accumulators, the taps are
    // constant-bank operands (a plain tap-by-tap loop is bound by shared-memory bandwidth: 2 loads per FMA)
    for (int i0 = tid * TBC_OPT; i0 <= dist; i0 += TBC_THREADS * TBC_OPT) {
        double acc[TBC_OPT];
        LDD_UNROLL
        for (int o = 0; o < TBC_OPT; ++o) acc[o] = 0.0;
        const double* y = ys + i0;                        // y[i0-H] ... ; output o uses y[o + m], m = 0..NTAPS-1
        LDD_UNROLL
        for (int t = 0; t < TBC_NTAPS + TBC_OPT - 1; ++t) {
            const double yv = y[t];
            LDD_UNROLL
            for (int o = 0; o < TBC_OPT; ++o) {
                const int m = t - o;
                if (m >= 0 && m < TBC_NTAPS) acc[o] = fma(c_tbc_taps[m], yv, acc[o]);
            }
        }
        LDD_UNROLL
        for (int o = 0; o < TBC_OPT; ++o)
            if (i0 + o <= dist) Ms[i0 + o] = acc[o];
    }
    __syncthreads();
    if (tid == 0) {
        double L = Ms[0] - 2.0 * Ms[1] + Ms[2];
        double R = Ms[dist] - 2.0 * Ms[dist - 1] + Ms[dist - 2];
        double A = (1.0 - r) * (1.0 - r);
        double q = pow(r, (double)(dist - 2));
        double den = A * (1.0 - q * q);
        s_ab[0] = (-L + q * R) / den;
        s_ab[1] = (-R + q * L) / den;
    }
    __syncthreads();
    {
        const double alpha = s_ab[0], beta = s_ab[1];
        for (int i = tid; i <= dist; i += TBC_THREADS) {
            double corr = 0.0;
            if (i < 64) corr += alpha * pow(r, (double)i);
            if (dist - i < 64) corr += beta * pow(r, (double)(dist - i));
            Ms[i] += corr;
        }
    }
    __syncthreads();
    // evaluate (np.linspace(fb, fb + (e - b), W + 1)[:-1]: x_j = j * step + fb)
    const double fb = b - (double)ib;
    const double stop = (e - b) + fb;
    const double step = (stop - fb) / (double)W;
    const double wowf = p.wow ? (e - b) / (double)p.linelen : 1.0;
    const double* y0 = ys + TBC_H;
    const double sixth = 1.0 / 6.0;
    for (int j = tid; j < W; j += TBC_THREADS) {
        double x = (double)j * step + fb;
        int i = (int)x;
        if (i > dist - 1) i = dist - 1;
        double t = x - (double)i, u = 1.0 - t;
        double Mi = Ms[i], Mj = Ms[i + 1];
        double S = Mi * u * u * u * sixth + Mj * t * t * t * sixth + (y0[i] - Mi * sixth) * u + (y0[i + 1] - Mj * sixth) * t;
        double hz = (S + p.plane_add) * wowf;
        size_t o = (size_t)field * (size_t)p.out_stride + (size_t)line * W + j;
        if (p.mode == 0) {
            ((double*)outbase)[o] = hz;
        } else {
            double red = (hz - p.ire0) / p.hz_ire;
            red -= p.vsync_ire;
            double v = red * p.out_scale + p.out_off;
            v = v < 0.0 ? 0.0 : (v > 65535.0 ? 65535.0 : v);
            unsigned short q16 = (unsigned short)(v + 0.5);
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

}  // namespace ldd

using namespace ldd;

extern "C" int ldd_tbc_fields(ldd_handle* h, const float* plane_dev, long long n, double plane_add,
                              const long long* base_dev, const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                              int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                              void* out_dev, long long out_stride, const float* burstlevel_dev, double colorlevel,
                              int* status_dev, void* stream) {
    if (!h || !plane_dev || !linelocs_dev || !linecount_dev || !out_dev || !status_dev) return LDD_EINVAL;
    if (nfields <= 0 || max_linecount <= 0) return LDD_OK;
    if (outwidth < 1 || (mode != 0 && mode != 1)) return LDD_EINVAL;
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
    p.burstlevel = burstlevel_dev;
    // np.uint16(327.67 * clevel * np.abs(float32 burstlevel)): python-float product, then float32 arithmetic
    double clevel = (1.0 / colorlevel) / (1700000.0 / 140.0);
    p.clevel_k = (float)(327.67 * clevel);
    p.status = status_dev;
    size_t smem = (size_t)(TBC_MAXD + 2 * TBC_H + TBC_OPT + 2 + TBC_MAXD + TBC_OPT + 2) * sizeof(double);
    if (!h->tbc_taps_set) {
        const double r = -0.26794919243112270647, c = 0.28867513459481288225;
        double taps[TBC_NTAPS];
        auto g = [&](int q) -> double { int a = q < 0 ? -q : q; return a > TBC_K ? 0.0 : c * pow(r, (double)a); };
        for (int m = 0; m < TBC_NTAPS; ++m) { int k = m - TBC_H; taps[m] = 6.0 * (g(k - 1) - 2.0 * g(k) + g(k + 1)); }
        cudaMemcpyToSymbol(c_tbc_taps, taps, sizeof taps);
        h->tbc_taps_set = true;
    }
    cudaFuncSetAttribute(tbc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaStream_t st = (cudaStream_t)stream;
    LDD_LAUNCH(tbc_kernel, dim3(max_linecount, nfields), dim3(TBC_THREADS), smem, st, p);
    return launch_status(h, "tbc_kernel");
}

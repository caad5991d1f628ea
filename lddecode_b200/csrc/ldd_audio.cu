// Second audio stage, RFDecode.audio_phase2 / runfilter_audio_phase2 (lddecode_core.py:335-371), and the
// 48 kHz PCM resample along the line positions, downscale_audio (lddecode_core.py:431-484).
//
// audio_phase2: blocks of N phase-1 samples -> FFT(N) -> keep the lowest and highest N/8 bins -> * audio_lpf2
// -> IFFT(N/4).real / 4, hop N-256, first 64 outputs of every later block dropped, last block re-anchored at
// len-N-1.  float64 throughout (the audio rate is fs/8 or fs/16, so this is a fraction of a percent of the
// work).  Both channels of a block ride through ONE complex transform each way: z = left + j right forward
// (the two spectra are separated on the N/4 bins that are kept), and the two real outputs come back as the real
// and imaginary part of one inverse transform of YL + j YR (after making the two self-paired bins real, which is
// what taking .real of the reference's inverse does to them).  The block geometry is closed form in the block
// index, so the call only enqueues a memset and one kernel: no job table, no allocation, no synchronisation.
#include "ldd_internal.h"

namespace ldd {

struct Audio2Geom {
    long long len, out_len;
    long long hop;          // N - 4 * 64
    int nmid;               // middle blocks (in_start = j * hop, j = 1 .. nmid)
    int N;
};

constexpr int A2_SKIP = 64;

__global__ void __launch_bounds__(512, 1)
audio2_kernel(const double* __restrict__ in_l, const double* __restrict__ in_r, double* out_l, double* out_r, Audio2Geom g,
              FftPlan plan_n, FftPlan plan_q, const Cx<double>* __restrict__ WN, const Cx<double>* __restrict__ lpf2,
              void* scratch, size_t per_cta) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    Cx<double>* b0 = (Cx<double>*)((char*)scratch + (size_t)blockIdx.x * per_cta);
    const int N = g.N;
    Cx<double>* b1 = b0 + N;
    const int Q = N / 4, E = N / 8;
    const int njobs = g.nmid + 2;
    const long long tail0 = g.out_len - (Q - A2_SKIP);        // the re-anchored last block owns out[tail0 ..)
    for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
        long long in_start, out_pos;
        int skip;
        const bool last = j == njobs - 1;
        if (j == 0) { in_start = 0; out_pos = 0; skip = 0; }
        else if (last) { in_start = g.len - N - 1; out_pos = tail0; skip = A2_SKIP; }
        else { in_start = (long long)j * g.hop; out_pos = Q + (long long)(j - 1) * (Q - A2_SKIP); skip = A2_SKIP; }
        for (int i = tid; i < N; i += nthr) {
            const long long s = in_start + i;
            b0[i] = s < g.len ? mk<double>(in_l[s], in_r[s]) : mk<double>(0.0, 0.0);
        }
        __syncthreads();
        Cx<double>* spec = fft_run<double, false>(b0, b1, plan_n, WN, 1, tid, nthr);
        Cx<double>* fr = (spec == b0) ? b1 : b0;
        for (int i = tid; i < Q; i += nthr) {
            const int k = (i < E) ? i : N - Q + i;
            const Cx<double> zk = spec[k], zm = conj(spec[(N - k) & (N - 1)]);
            Cx<double> yl = scale(zk + zm, 0.5) * lpf2[i];
            Cx<double> yr = scale(mul_mj(zk - zm), 0.5) * lpf2[i];
            if (i == 0 || i == E) { yl.y = 0.0; yr.y = 0.0; }
            // W = YL + j YR, stored conjugated for the inverse-by-forward transform
            fr[i] = mk<double>(yl.x - yr.y, -(yl.y + yr.x));
        }
        __syncthreads();
        Cx<double>* r = fft_run<double, false>(fr, fr + Q, plan_q, WN, 4, tid, nthr);
        const double sc = 1.0 / (double)N;       // 1/(N/4) of the inverse transform, / audio_fdiv2 = 4
        for (int i = skip + tid; i < Q; i += nthr) {
            const long long o = out_pos + (i - skip);
            // the last block overwrites the tail after everything else in the reference (lddecode_core.py:368-369):
            // here the other blocks simply do not write there
            if (o >= 0 && o < g.out_len && (last || o < tail0)) {
                out_l[o] = r[i].x * sc;
                out_r[o] = -r[i].y * sc;
            }
        }
        __syncthreads();
    }
}

// downscale_audio (lddecode_core.py:431-484), batched over fields.  Output sample i of field f sits at time
// t = np.arange(t0, ..)[i], on line linenum = t * 1e6 / line_period + 1; its input position is
// interpolated between the two neighbouring line positions, divided by `scale` and truncated to an index into the
// phase-2 audio of the field's window.
struct PcmParams {
    const double* audio_l;
    const double* audio_r;
    long long audio_len;
    const long long* abase;      // [nfields] index of the window's audio sample 0 in audio_l/r (NULL: 0)
    const double* linelocs;      // [nfields][ll_stride]
    const int* nll;              // [nfields] entries of the field's line table (linecount + 4), less nll_add
    int nll_add;
    const double* fbase;         // range-wide audio (ldd_pipe_pcm): [nfields] position of the window's audio sample 0 in
                                 // audio_l/r, in audio samples and in general fractional; NULL: the audio is the window's own
    const double* t0;            // [nfields] arange start
    const double* t1;            // [nfields] arange's second value, t0 + 1/freq (np.arange fills start + i * (t1 - t0) from i = 2 on)
    const int* nout;             // [nfields] stereo samples to produce
    const long long* out_off;    // [nfields] first int16 of the field in out
    int ll_stride;
    double lineloc_add, line_period, linelen, scale, lfreq, rfreq;
    short* out;
    int* status;                 // bit 4 (16): an index left the audio array (the reference raises IndexError)
};

__global__ void __launch_bounds__(256) pcm_kernel(const PcmParams p) {
    const int f = blockIdx.y;
    const int n = p.nout[f];
    const double* ll = p.linelocs + (size_t)f * p.ll_stride;
    const int nll = p.nll[f] + p.nll_add;
    const long long ab = p.abase ? p.abase[f] : 0;
    const double fb = p.fbase ? p.fbase[f] : 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double t0 = p.t0[f], t1 = p.t1[f];
        const double t = i == 0 ? t0 : (i == 1 ? t1 : t0 + (double)i * (t1 - t0));
        const double linenum = ((t * 1000000.0) / p.line_period) + 1.0;
        const int li = (int)linenum;
        short ol = 0, orr = 0;
        if (li < 0 || li >= nll) {
            atomicOr(&p.status[f], 16);
        } else {
            const double cur = ll[li] + p.lineloc_add;
            const double nxt = (li + 1 < nll) ? ll[li + 1] + p.lineloc_add : cur + p.linelen;
            double sampleloc = cur;
            sampleloc += (nxt - cur) * (linenum - floor(linenum));
            const double swow = (nxt - cur) / p.linelen;
            long long idx = (long long)(sampleloc / p.scale) + ab;
            double l = 0.0, r = 0.0;
            bool ok;
            if (p.fbase) {
                // The reference reads sample idx of the window's own phase-2 audio.  On the range-wide audio that instant
                // lies at x = fbase + idx, between two samples when the window does not start on the range's audio
                // grid: four-point Lagrange interpolation there (weights 0, 1, 0, 0 when x is whole, i.e. the very
                // sample the reference reads).
                const double x = fb + (double)idx;
                const double xf = floor(x), u = x - xf;
                idx = (long long)xf;
                ok = idx >= 1 && idx + 2 < p.audio_len;
                if (ok) {
                    const double w0 = -u * (u - 1.0) * (u - 2.0) * (1.0 / 6.0), w1 = (u + 1.0) * (u - 1.0) * (u - 2.0) * 0.5;
                    const double w2 = -(u + 1.0) * u * (u - 2.0) * 0.5, w3 = (u + 1.0) * u * (u - 1.0) * (1.0 / 6.0);
                    if (u == 0.0) {
                        l = p.audio_l[idx]; r = p.audio_r[idx];
                    } else {
                        l = w0 * p.audio_l[idx - 1] + w1 * p.audio_l[idx] + w2 * p.audio_l[idx + 1] + w3 * p.audio_l[idx + 2];
                        r = w0 * p.audio_r[idx - 1] + w1 * p.audio_r[idx] + w2 * p.audio_r[idx + 1] + w3 * p.audio_r[idx + 2];
                    }
                }
            } else {
                ok = idx >= 0 && idx < p.audio_len;
                if (ok) { l = p.audio_l[idx]; r = p.audio_r[idx]; }
            }
            if (!ok) {
                atomicOr(&p.status[f], 16);
            } else {
                l *= swow; r *= swow;
                l -= p.lfreq; r -= p.rfreq;
                // int(np.round(x * 32767 / 150000)): round half to even; np.clip(-32766, 32766)
                double vl = rint(l * 32767.0 / 150000.0), vr = rint(r * 32767.0 / 150000.0);
                vl = vl < -32766.0 ? -32766.0 : (vl > 32766.0 ? 32766.0 : vl);
                vr = vr < -32766.0 ? -32766.0 : (vr > 32766.0 ? 32766.0 : vr);
                ol = (short)vl; orr = (short)vr;
            }
        }
        short* o = p.out + p.out_off[f] + 2 * (long long)i;
        o[0] = ol;
        o[1] = orr;
    }
}

}  // namespace ldd

using namespace ldd;

// ldd_pipe_pcm's launch: the audio is the range's own (audio sample 0 <-> plane sample 0), fbase_dev[f] the field window's
// first plane sample divided by the total audio decimation, linecount_dev the fields' line counts (tables hold 4 more).
int ldd::pcm_range_launch(ldd_handle* h, const double* audio_l, const double* audio_r, long long audio_len, const double* fbase_dev,
                          const double* linelocs_dev, int ll_stride, const int* linecount_dev, const double* t0_dev,
                          const double* t1_dev, const int* nout_dev, const long long* out_off_dev, int nfields, int max_nout,
                          double lineloc_add, double scale, double line_period_us, double lfreq, double rfreq, short* out_dev,
                          int* status_dev, cudaStream_t st) {
    if (nfields <= 0 || max_nout <= 0) return LDD_OK;
    PcmParams p;
    p.audio_l = audio_l; p.audio_r = audio_r; p.audio_len = audio_len; p.abase = nullptr; p.fbase = fbase_dev;
    p.linelocs = linelocs_dev; p.nll = linecount_dev; p.nll_add = 4; p.t0 = t0_dev; p.t1 = t1_dev; p.nout = nout_dev;
    p.out_off = out_off_dev; p.ll_stride = ll_stride; p.lineloc_add = lineloc_add; p.line_period = line_period_us;
    p.linelen = (double)h->cfg.linelen; p.scale = scale; p.lfreq = lfreq; p.rfreq = rfreq; p.out = out_dev; p.status = status_dev;
    LDD_LAUNCH(pcm_kernel, dim3((max_nout + 255) / 256, nfields), dim3(256), 0, st, p);
    return launch_status(h, "pcm_kernel");
}

extern "C" int ldd_audio_phase2(ldd_handle* h, const double* in_l_dev, const double* in_r_dev, long long len,
                                double* out_l_dev, double* out_r_dev, void* stream) {
    if (!h || !in_l_dev || !in_r_dev || !out_l_dev || !out_r_dev) return LDD_EINVAL;
    if (!h->have_filter[LDD_F_AUDIO_LPF2]) { h->err = "audio_lpf2 not set"; return LDD_EINVAL; }
    const int N = h->cfg.blocklen;
    if (len < (long long)N + 1) { h->err = "audio_phase2 needs at least blocklen+1 samples (the reference raises)"; return LDD_EINVAL; }
    Audio2Geom g;
    g.len = len; g.out_len = len / 4; g.hop = N - A2_SKIP * 4; g.N = N;
    // range(hop, len - hop, hop), stopped where a block would leave the array: the reference raises on such a short
    // block (shape mismatch in runfilter_audio_phase2); the re-anchored last block covers that tail anyway
    long long nmid = 0;
    for (long long s = g.hop; s < len - g.hop && s + N <= len; s += g.hop) ++nmid;
    g.nmid = (int)nmid;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per_cta = (size_t)2 * N * sizeof(Cx<double>);
    const int maxgrid = (int)(h->scratch_bytes / per_cta);
    if (maxgrid < 1) { h->err = "audio_phase2: no scratch"; return LDD_ENOMEM; }
    const int njobs = g.nmid + 2;
    const int grid = njobs < maxgrid ? njobs : maxgrid;
    cudaMemsetAsync(out_l_dev, 0, (size_t)g.out_len * sizeof(double), st);
    cudaMemsetAsync(out_r_dev, 0, (size_t)g.out_len * sizeof(double), st);
    LDD_LAUNCH(audio2_kernel, dim3(grid), dim3(512), 0, st, in_l_dev, in_r_dev, out_l_dev, out_r_dev, g, make_plan(N), make_plan(N / 4),
               (const Cx<double>*)h->d_WNfull, (const Cx<double>*)h->d_lpf2, h->scratch, per_cta);
    return launch_status(h, "audio2_kernel");
}

extern "C" int ldd_downscale_audio(ldd_handle* h, const double* audio_l_dev, const double* audio_r_dev, long long audio_len,
                                   const long long* audio_base_dev, const double* linelocs_dev, int ll_stride,
                                   const int* nll_dev, const double* t0_dev, const double* t1_dev, const int* nout_dev,
                                   const long long* out_off_dev, int nfields, int max_nout, double lineloc_add, double scale,
                                   double line_period_us, double audio_lfreq, double audio_rfreq,
                                   short* out_dev, int* status_dev, void* stream) {
    if (!h || !audio_l_dev || !audio_r_dev || !linelocs_dev || !nll_dev || !t0_dev || !t1_dev || !nout_dev || !out_off_dev ||
        !out_dev || !status_dev || scale <= 0) return LDD_EINVAL;
    if (nfields <= 0 || max_nout <= 0) return LDD_OK;
    PcmParams p;
    p.audio_l = audio_l_dev; p.audio_r = audio_r_dev; p.audio_len = audio_len; p.abase = audio_base_dev;
    p.linelocs = linelocs_dev; p.nll = nll_dev; p.nll_add = 0; p.fbase = nullptr; p.t0 = t0_dev; p.t1 = t1_dev; p.nout = nout_dev; p.out_off = out_off_dev;
    p.ll_stride = ll_stride; p.lineloc_add = lineloc_add;
    // SysParams line_period / audio_lfreq / audio_rfreq (lddecode_core.py:43-44, 51, 62, 72-73) come from the caller
    p.line_period = line_period_us;
    p.linelen = (double)h->cfg.linelen;
    p.scale = scale;
    p.lfreq = audio_lfreq;
    p.rfreq = audio_rfreq;
    p.out = out_dev; p.status = status_dev;
    LDD_LAUNCH(pcm_kernel, dim3((max_nout + 255) / 256, nfields), dim3(256), 0, (cudaStream_t)stream, p);
    return launch_status(h, "pcm_kernel");
}

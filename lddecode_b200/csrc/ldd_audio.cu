// Second audio stage, RFDecode.audio_phase2 / runfilter_audio_phase2 (lddecode_core.py:335-371):
// blocks of N phase-1 samples -> FFT(N) -> keep the lowest and highest N/8 bins -> * audio_lpf2
// -> IFFT(N/4).real / 4, hop N-256, first 64 outputs of every later block dropped, last block
// re-anchored at len-N-1.  float64 throughout (the audio rate is fs/8 or fs/16, so this is a
// fraction of a percent of the work); one CTA per (channel, block) out of the L2-resident scratch.
#include "ldd_internal.h"

namespace ldd {

struct Audio2Job {
    long long in_start;   // first phase-1 sample of the block
    long long out_pos;    // where out[skip] lands in the output
    int skip;             // 0 for the first block, 64 afterwards
    int ch;               // 0 left, 1 right
};

__global__ void __launch_bounds__(256, 1)
audio2_kernel(const double* __restrict__ in_l, const double* __restrict__ in_r, double* out_l, double* out_r,
              long long out_len, const Audio2Job* __restrict__ jobs, int njobs, int N, FftPlan plan_n, FftPlan plan_q,
              const Cx<double>* __restrict__ WN, const Cx<double>* __restrict__ lpf2, void* scratch, size_t per_cta) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    Cx<double>* b0 = (Cx<double>*)((char*)scratch + (size_t)blockIdx.x * per_cta);
    Cx<double>* b1 = b0 + N;
    const int Q = N / 4, E = N / 8;
    for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
        const Audio2Job job = jobs[j];
        const double* in = job.ch ? in_r : in_l;
        double* out = job.ch ? out_r : out_l;
        for (int i = tid; i < N; i += nthr) b0[i] = mk<double>(in[job.in_start + i], 0.0);
        __syncthreads();
        Cx<double>* spec = fft_run<double, false>(b0, b1, plan_n, WN, 1, tid, nthr);
        Cx<double>* fr = (spec == b0) ? b1 : b0;
        for (int i = tid; i < Q; i += nthr) {
            Cx<double> s = (i < E) ? spec[i] : spec[N - Q + i];
            fr[i] = conj(s * lpf2[i]);
        }
        __syncthreads();
        Cx<double>* r = fft_run<double, false>(fr, fr + Q, plan_q, WN, 4, tid, nthr);
        const double sc = 1.0 / (double)N;       // 1/(N/4) of the inverse transform, / audio_fdiv2 = 4
        for (int i = job.skip + tid; i < Q; i += nthr) {
            long long o = job.out_pos + (i - job.skip);
            if (o >= 0 && o < out_len) out[o] = r[i].x * sc;
        }
        __syncthreads();
    }
}

}  // namespace ldd

using namespace ldd;

extern "C" int ldd_audio_phase2(ldd_handle* h, const double* in_l_dev, const double* in_r_dev, long long len,
                                double* out_l_dev, double* out_r_dev, void* stream) {
    if (!h || !in_l_dev || !in_r_dev || !out_l_dev || !out_r_dev) return LDD_EINVAL;
    if (!h->have_filter[LDD_F_AUDIO_LPF2]) { h->err = "audio_lpf2 not set"; return LDD_EINVAL; }
    const int N = h->cfg.blocklen, Q = N / 4, askip = 64;
    if (len < (long long)N + 1) { h->err = "audio_phase2 needs at least blocklen+1 samples (the reference raises)"; return LDD_EINVAL; }
    const long long out_len = len / 4;
    const long long hop = N - askip * 4;
    std::vector<Audio2Job> jobs, last;
    for (int ch = 0; ch < 2; ++ch) {
        jobs.push_back({0, 0, 0, ch});
        long long pos = Q;
        for (long long s = hop; s < len - hop; s += hop) {
            jobs.push_back({s, pos, askip, ch});
            pos += Q - askip;
        }
        last.push_back({len - N - 1, out_len - (Q - askip), askip, ch});
    }
    cudaStream_t st = (cudaStream_t)stream;
    size_t nj = jobs.size() + last.size();
    Audio2Job* d_jobs = nullptr;
    if (cudaMalloc((void**)&d_jobs, nj * sizeof(Audio2Job)) != cudaSuccess) { h->err = "cudaMalloc jobs"; return LDD_ECUDA; }
    std::vector<Audio2Job> all(jobs);
    all.insert(all.end(), last.begin(), last.end());
    cudaMemcpyAsync(d_jobs, all.data(), nj * sizeof(Audio2Job), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);         // `all` is pageable host memory
    size_t per_cta = (size_t)2 * N * sizeof(Cx<double>);
    int maxgrid = (int)(h->scratch_bytes / per_cta);
    FftPlan pn = make_plan(N), pq = make_plan(Q);
    cudaMemsetAsync(out_l_dev, 0, (size_t)out_len * sizeof(double), st);
    cudaMemsetAsync(out_r_dev, 0, (size_t)out_len * sizeof(double), st);
    int g1 = (int)jobs.size() < maxgrid ? (int)jobs.size() : maxgrid;
    LDD_LAUNCH(audio2_kernel, dim3(g1), dim3(256), 0, st, in_l_dev, in_r_dev, out_l_dev, out_r_dev, out_len,
               (const Audio2Job*)d_jobs, (int)jobs.size(), N, pn, pq, (const Cx<double>*)h->d_WNfull,
               (const Cx<double>*)h->d_lpf2, h->scratch, per_cta);
    // the re-anchored last block overwrites the tail, after everything else (lddecode_core.py:368-369)
    LDD_LAUNCH(audio2_kernel, dim3(2), dim3(256), 0, st, in_l_dev, in_r_dev, out_l_dev, out_r_dev, out_len,
               (const Audio2Job*)(d_jobs + jobs.size()), 2, N, pn, pq, (const Cx<double>*)h->d_WNfull,
               (const Cx<double>*)h->d_lpf2, h->scratch, per_cta);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(st);
    cudaFree(d_jobs);
    if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return LDD_ECUDA; }
    return LDD_OK;
}

// C-ABI glue of libldd_b200.so: handle, table upload, range geometry, launches.
#include "ldd_internal.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace ldd;

namespace {

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int fail(ldd_handle* h, int code, const char* fmt, ...) {
    if (h) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}

#define CUDA_TRY(h, expr)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) return fail((h), LDD_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// Upload a complex table given in double to a fp64 and a fp32 device copy.
int upload_both(ldd_handle* h, const std::vector<Cx<double>>& t, void** slot /*[2]*/) {
    size_t n = t.size();
    for (int i = 0; i < 2; ++i)
        if (slot[i]) { cudaFree(slot[i]); slot[i] = nullptr; }
    CUDA_TRY(h, cudaMalloc(&slot[0], n * sizeof(Cx<double>)));
    CUDA_TRY(h, cudaMalloc(&slot[1], n * sizeof(Cx<float>)));
    std::vector<Cx<float>> f(n);
    for (size_t i = 0; i < n; ++i) f[i] = mk<float>((float)t[i].x, (float)t[i].y);
    CUDA_TRY(h, cudaMemcpy(slot[0], t.data(), n * sizeof(Cx<double>), cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(slot[1], f.data(), n * sizeof(Cx<float>), cudaMemcpyHostToDevice));
    return LDD_OK;
}

std::vector<Cx<double>> roots(int n, int count) {
    std::vector<Cx<double>> t(count);
    for (int k = 0; k < count; ++k) {
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        t[k] = mk<double>((double)cosl(a), (double)sinl(a));
    }
    return t;
}

// Hv = RFVideo * MTF ** level (lddecode_core.py:290-293) on the device, float64 and float32 copies: a level change
// (the reference adapts it per CAV frame, :1300-1306) is one tiny stream-ordered kernel instead of a host-side
// complex power over N entries and a blocking upload.  level 0 / 1 are exact like numpy's integer-power path;
// other levels are exp(level * log z) as libm's cpow.
__global__ void hv_kernel(const Cx<double>* __restrict__ rf, const Cx<double>* __restrict__ mtf, double level,
                          Cx<double>* hv64, Cx<float>* hv32, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Cx<double> m = mk<double>(1.0, 0.0);
    if (level == 1.0) m = mtf[k];
    else if (level != 0.0) {
        const Cx<double> z = mtf[k];
        const double mag = exp(level * log(hypot(z.x, z.y))), ang = level * atan2(z.y, z.x);
        m = mk<double>(mag * cos(ang), mag * sin(ang));
    }
    const Cx<double> v = rf[k] * m;
    hv64[k] = v;
    hv32[k] = mk<float>((float)v.x, (float)v.y);
}

// dst[part stride + pos(k)] = src[part M + k] for nparts parts of M = 8192 entries (ldd_fft2.cuh's digit-permuted order),
// the `tail` entries behind them copied as they are.  padded: positions in the block arrays' padded layout (one slot after
// every 16, stride = 8704 per part), so that a part is ONE contiguous range in global and in shared memory and a single
// bulk copy brings it in (ldd_demod8k.cuh).
template <class T>
__global__ void permute_kernel(const Cx<T>* __restrict__ src, Cx<T>* __restrict__ dst, int nparts, int tail, int padded) {
    const int M = f2::M, stride = padded ? f2::span<1>() : M;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nparts * M) {
        const int pos = f2::pos_of_idx(k % M);
        dst[(k / M) * stride + (padded ? f2::pix<1>(pos) : pos)] = src[k];
    } else if (k < nparts * M + tail) {
        dst[nparts * stride + (k - nparts * M)] = src[k];
    }
}

// (re)builds the permuted float32 copy (padded layout) and, for the mixed lane, the float64 one (plain) on `stream`; no-op
// for other block lengths
int permuted_copy(ldd_handle* h, void* const* src /*[2]: float64, float32*/, void** dst32, void** dst64, int nparts, int tail, cudaStream_t stream) {
    if (h->cfg.blocklen != 2 * f2::M || !src[1]) return LDD_OK;
    const int n = nparts * f2::M + tail;
    if (!*dst32) {
        const size_t bytes = ((size_t)nparts * f2::span<1>() + tail) * sizeof(Cx<float>);
        CUDA_TRY(h, cudaMalloc(dst32, bytes));
        CUDA_TRY(h, cudaMemset(*dst32, 0, bytes));
    }
    LDD_LAUNCH(permute_kernel<float>, dim3((n + 255) / 256), dim3(256), 0, stream, (const Cx<float>*)src[1], (Cx<float>*)*dst32, nparts, tail, 1);
    if (dst64 && h->cfg.precision == LDD_PREC_MIXED && src[0]) {
        if (!*dst64) CUDA_TRY(h, cudaMalloc(dst64, (size_t)n * sizeof(Cx<double>)));
        LDD_LAUNCH(permute_kernel<double>, dim3((n + 255) / 256), dim3(256), 0, stream, (const Cx<double>*)src[0], (Cx<double>*)*dst64, nparts, tail, 0);
    }
    return launch_status(h, "permute_kernel");
}

}  // namespace

extern "C" {

int ldd_abi_version(void) { return LDD_ABI_VERSION; }

int ldd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* ldd_last_error(ldd_handle* h) { return h ? h->err.c_str() : "null handle"; }

int ldd_create(const ldd_config* cfg, ldd_handle** out) {
    if (!cfg || !out) return LDD_EINVAL;
    *out = nullptr;
    if (cfg->abi_version != LDD_ABI_VERSION) return LDD_EINVAL;
    if (!is_pow2(cfg->blocklen) || cfg->blocklen < 4096 || cfg->blocklen > 262144) return LDD_EINVAL;
    if (cfg->blockcut < 0 || cfg->blockcut_end < 0 || cfg->blockcut + cfg->blockcut_end >= cfg->blocklen / 2) return LDD_EINVAL;
    if (cfg->system != LDD_SYSTEM_NTSC && cfg->system != LDD_SYSTEM_PAL) return LDD_EINVAL;
    // The refinement / TBC kernels stage fixed windows per line (4.7 us of pilot, 4 us around an hsync edge, the 40 burst
    // samples, a line of up to 1.25 x nominal length): sized for 21 .. 40.5 MSPS (8fsc NTSC / PAL and the reference's
    // 40 MSPS default).  Outside that range fields would be flagged invalid at run time: refuse the configuration instead.
    if (!(cfg->freq_hz >= 21.0e6 && cfg->freq_hz <= 40.5e6) || cfg->linelen < 1200 || cfg->linelen > 2700 || cfg->outlinelen < 600) return LDD_EINVAL;
    if (ldd_device_count() <= cfg->device) return LDD_ECUDA;
    ldd_handle* h = new (std::nothrow) ldd_handle();
    if (!h) return LDD_ENOMEM;
    h->cfg = *cfg;
    h->device = cfg->device;
    memset(h->d_WM, 0, sizeof h->d_WM);
    memset(h->d_WN, 0, sizeof h->d_WN);
    memset(h->d_Hv, 0, sizeof h->d_Hv);
    memset(h->d_F, 0, sizeof h->d_F);
    memset(h->d_AL, 0, sizeof h->d_AL);
    memset(h->d_AR, 0, sizeof h->d_AR);
    memset(h->have_filter, 0, sizeof h->have_filter);
    h->scratch = nullptr;
    h->scratch_bytes = 0;
    h->d_lpf2 = nullptr;
    h->d_WNfull = nullptr;
    *out = h;                       // from here on the caller can read the error text
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(h, LDD_ECUDA, "cudaSetDevice(%d) failed", cfg->device);
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    const int N = cfg->blocklen, M = N / 2;
    h->A = 0;
    if (cfg->decode_analog_audio) {
        int w = cfg->audio_slice_hi - cfg->audio_slice_lo;
        if (w <= 0 || !is_pow2(2 * w) || cfg->audio_slice_lo < 1 || cfg->audio_slice_hi >= M || 2 * w > M / 2)
            return fail(h, LDD_EINVAL, "audio slice [%d,%d) unsupported", cfg->audio_slice_lo, cfg->audio_slice_hi);
        h->A = 2 * w;
        int ds = N / h->A;
        long long stride = N - cfg->blockcut - cfg->blockcut_end;
        if (cfg->blockcut % ds || stride % ds)
            return fail(h, LDD_EINVAL, "blockcut/stride not a multiple of the audio decimation %d", ds);
    }
    int rc = upload_both(h, roots(M, M), h->d_WM);
    if (rc) return rc;
    rc = upload_both(h, roots(N, M), h->d_WN);
    if (rc) return rc;
    {
        std::vector<Cx<double>> t = roots(N, N);
        CUDA_TRY(h, cudaMalloc(&h->d_WNfull, (size_t)N * sizeof(Cx<double>)));
        CUDA_TRY(h, cudaMemcpy(h->d_WNfull, t.data(), (size_t)N * sizeof(Cx<double>), cudaMemcpyHostToDevice));
    }
    // launch geometry / scratch of the demodulation kernel.  Tunables (development): LDD_THREADS,
    // LDD_RADIX_MAX, LDD_CTAS_PER_SM.
    const bool f64 = cfg->precision == LDD_PREC_F64;
    const bool mixed = cfg->precision == LDD_PREC_MIXED;
    if (cfg->precision < LDD_PREC_F64 || cfg->precision > LDD_PREC_MIXED) return fail(h, LDD_EINVAL, "bad precision %d", cfg->precision);
    const char* env = getenv("LDD_CTAS_PER_SM");
    int per_sm = env ? atoi(env) : 1;
    if (per_sm < 1) per_sm = 1;
    env = getenv("LDD_THREADS");
    h->threads = env ? atoi(env) : 512;
    if (f64) { if (h->threads != 256 && h->threads != 512 && h->threads != 1024 && h->threads != 2256) h->threads = 512; }
    else { if (h->threads != 512 && h->threads != 1024) h->threads = 512; }
    env = getenv("LDD_RADIX_MAX");
    h->radix_max = env ? atoi(env) : 16;
    if (h->radix_max != 4 && h->radix_max != 8 && h->radix_max != 16) h->radix_max = 16;
    size_t esz = f64 ? sizeof(Cx<double>) : sizeof(Cx<float>);
    size_t per_cta = (size_t)3 * M * esz;
    size_t per_cta_padded = (size_t)3 * pspan<true>(M) * esz;
    bool smem_lane = !f64 && per_cta_padded + 2048 <= h->smem_optin;
    h->smem_bytes = smem_lane ? per_cta_padded : 0;
    h->grid = smem_lane ? h->sm_count : h->sm_count * per_sm;
    h->scratch_per_cta = smem_lane ? 0 : per_cta;
    // float64 lane(s): a padded shared-memory buffer as the ping-pong partner of the length-M transforms when it fits
    {
        size_t want = (size_t)pspan<true>(M) * sizeof(Cx<double>);
        bool uses_f64 = f64 || mixed;
        if (uses_f64 && want + 2048 <= h->smem_optin && !getenv("LDD_NO_SMEM_PARTNER") && (f64 ? h->threads == 512 : true))
            h->sp_bytes = want;
    }
    // audio phase 2 runs out of the same scratch: two length-N complex128 buffers per CTA
    size_t per_cta_a2 = (size_t)2 * N * sizeof(Cx<double>);
    size_t need = (size_t)h->grid * (h->scratch_per_cta > per_cta_a2 ? h->scratch_per_cta : per_cta_a2);
    CUDA_TRY(h, cudaMalloc(&h->scratch, need));
    h->scratch_bytes = need;
    if (mixed) {
        // second pass of the mixed lane: float64, global scratch, one CTA per SM
        h->scratch64_per_cta = (size_t)3 * M * sizeof(Cx<double>);
        CUDA_TRY(h, cudaMalloc(&h->scratch64, (size_t)h->sm_count * h->scratch64_per_cta));
        CUDA_TRY(h, cudaMalloc((void**)&h->d_queue, 2 * sizeof(int)));
        const char* es = getenv("LDD_SPARE_SMS");
        if (es) h->spare_sms = atoi(es);
        if (h->spare_sms < 0 || h->spare_sms > 64) h->spare_sms = 4;
        const char* em = getenv("LDD_FLAG_MARGIN_HZ");
        if (em) h->flag_margin = atof(em);
    }
#ifndef LDD_EMU
    // keep the scratch slices resident in L2 (persisting lines) while the planes stream through
    h->l2_window = 0;
    if (h->scratch_per_cta && !getenv("LDD_NO_L2_PERSIST")) {
        size_t want = (size_t)h->grid * h->scratch_per_cta;
        size_t cap = (size_t)prop.persistingL2CacheMaxSize;
        size_t win = (size_t)prop.accessPolicyMaxWindowSize;
        if (cap > 0 && win > 0) {
            size_t lim = want < cap ? want : cap;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess) h->l2_window = want < win ? want : win;
            h->l2_ratio = want <= lim ? 1.0f : (float)lim / (float)want;
        }
    }
#endif
    return LDD_OK;
}

void ldd_destroy(ldd_handle* h) {
    if (!h) return;
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->d_WM[i]); cudaFree(h->d_WN[i]); cudaFree(h->d_Hv[i]); cudaFree(h->d_AL[i]); cudaFree(h->d_AR[i]);
        for (int m = 0; m < 4; ++m) cudaFree(h->d_F[m][i]);
    }
    cudaFree(h->scratch);
    cudaFree(h->d_lpf2);
    cudaFree(h->d_WNfull);
    cudaFree(h->d_rfbase);
    cudaFree(h->d_mtf);
    cudaFree(h->d_lnM[0]); cudaFree(h->d_lnM[1]);
    cudaFree(h->d_HvP); cudaFree(h->d_lnMP);
    cudaFree(h->d_HvP64); cudaFree(h->d_lnMP64); cudaFree(h->d_FP64_05);
    for (int m = 0; m < 4; ++m) cudaFree(h->d_FP[m]);
    cudaFree(h->scratch64);
    cudaFree(h->d_flags);
    cudaFree(h->d_queue);
    cudaFree(h->peak_ws);
    cudaFree(h->pilot_ws);
    delete h;
}

int ldd_set_filter(ldd_handle* h, int id, const double* table, int n) {
    if (!h || !table) return LDD_EINVAL;
    const int N = h->cfg.blocklen, M = N / 2;
    const Cx<double>* src = (const Cx<double>*)table;
    std::vector<Cx<double>> t;
    switch (id) {
        case LDD_F_RFVIDEO: {
            if (n != N) return fail(h, LDD_EINVAL, "RFVideo table needs %d entries", N);
            t.assign(src, src + N);
            int rc = upload_both(h, t, h->d_Hv);
            if (rc) return rc;
            rc = permuted_copy(h, h->d_Hv, &h->d_HvP, &h->d_HvP64, 2, 0, 0);
            if (rc) return rc;
            // kept as the base of ldd_set_mtf_level
            h->mtf_level_set = 0.0;
            if (!h->d_rfbase) CUDA_TRY(h, cudaMalloc(&h->d_rfbase, (size_t)N * sizeof(Cx<double>)));
            CUDA_TRY(h, cudaMemcpy(h->d_rfbase, table, (size_t)N * sizeof(Cx<double>), cudaMemcpyHostToDevice));
            break;
        }
        case LDD_F_MTF: {
            if (n != N) return fail(h, LDD_EINVAL, "MTF table needs %d entries", N);
            h->mtf_level_set = -1e300;       // forces the next ldd_set_mtf_level to run
            if (!h->d_mtf) CUDA_TRY(h, cudaMalloc(&h->d_mtf, (size_t)N * sizeof(Cx<double>)));
            CUDA_TRY(h, cudaMemcpy(h->d_mtf, table, (size_t)N * sizeof(Cx<double>), cudaMemcpyHostToDevice));
            // principal logarithm, for the per-block level ramp (numpy's complex power is exp(level * log z) as well)
            t.resize(N);
            for (int k = 0; k < N; ++k) t[k] = mk<double>(std::log(std::hypot(src[k].x, src[k].y)), std::atan2(src[k].y, src[k].x));
            int rc = upload_both(h, t, h->d_lnM);
            if (rc) return rc;
            rc = permuted_copy(h, h->d_lnM, &h->d_lnMP, &h->d_lnMP64, 2, 0, 0);
            if (rc) return rc;
            break;
        }
        case LDD_F_VIDEO: case LDD_F_VIDEO05: case LDD_F_BURST: case LDD_F_PILOT: {
            if (n != N) return fail(h, LDD_EINVAL, "post filter table needs %d entries", N);
            int m = id - LDD_F_VIDEO;
            t.resize(M + 1);
            const long double twopi = 2.0L * 3.14159265358979323846264338327950288L;
            for (int k = 0; k <= M; ++k) {
                Cx<double> v = src[k];
                if (id == LDD_F_VIDEO05 && h->cfg.f05_offset) {
                    // np.roll(x, -off)  <=>  X[k] * e^{+2 pi i off k / N}   (lddecode_core.py:303)
                    long long kk = ((long long)k * h->cfg.f05_offset) % N;
                    long double a = twopi * (long double)kk / (long double)N;
                    v = v * mk<double>((double)cosl(a), (double)sinl(a));
                }
                t[k] = scale(v, 1.0 / (double)M);
            }
            h->dc[m] = src[0].x;
            int rc = upload_both(h, t, h->d_F[m]);
            if (rc) return rc;
            rc = permuted_copy(h, h->d_F[m], &h->d_FP[m], m == 1 ? &h->d_FP64_05 : nullptr, 1, 1, 0);
            if (rc) return rc;
            break;
        }
        case LDD_F_AUDIO_L: case LDD_F_AUDIO_R: {
            if (n != h->A || h->A == 0) return fail(h, LDD_EINVAL, "audio filter needs %d entries", h->A);
            t.assign(src, src + n);
            int rc = upload_both(h, t, id == LDD_F_AUDIO_L ? h->d_AL : h->d_AR);
            if (rc) return rc;
            break;
        }
        case LDD_F_AUDIO_LPF2: {
            if (n != N / 4) return fail(h, LDD_EINVAL, "audio_lpf2 needs %d entries", N / 4);
            if (h->d_lpf2) cudaFree(h->d_lpf2);
            CUDA_TRY(h, cudaMalloc(&h->d_lpf2, (size_t)n * sizeof(Cx<double>)));
            CUDA_TRY(h, cudaMemcpy(h->d_lpf2, table, (size_t)n * sizeof(Cx<double>), cudaMemcpyHostToDevice));
            break;
        }
        default:
            return fail(h, LDD_EINVAL, "unknown filter id %d", id);
    }
    // The copies above are staged from pageable memory and the permuted copies are built on the legacy stream; the
    // kernels that read the tables run on the caller's (non-blocking) streams, which the legacy stream does not order.
    // A table upload is set-up work: wait until it is in place.
    CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)0));
    h->have_filter[id] = true;
    return LDD_OK;
}

int ldd_set_mtf_level(ldd_handle* h, double level, void* stream) {
    if (!h) return LDD_EINVAL;
    if (!h->have_filter[LDD_F_RFVIDEO] || !h->have_filter[LDD_F_MTF] || !h->d_rfbase || !h->d_mtf)
        return fail(h, LDD_EINVAL, "ldd_set_mtf_level needs the RFVideo (level 0) and MTF tables");
    if (h->mtf_level_set == level) return LDD_OK;          // a handle serves one stream user: nothing to order against
    const int N = h->cfg.blocklen;
    h->mtf_level_set = level;
    LDD_LAUNCH(hv_kernel, dim3((N + 255) / 256), dim3(256), 0, (cudaStream_t)stream, (const Cx<double>*)h->d_rfbase,
               (const Cx<double>*)h->d_mtf, level, (Cx<double>*)h->d_Hv[0], (Cx<float>*)h->d_Hv[1], N);
    int rc = launch_status(h, "hv_kernel");
    if (rc) return rc;
    return permuted_copy(h, h->d_Hv, &h->d_HvP, &h->d_HvP64, 2, 0, (cudaStream_t)stream);
}

int ldd_set_mtf_ramp(ldd_handle* h, double pos0_sample, double period_samples, double step_per_period,
                     double hold_until_sample, double hold_level) {
    if (!h || period_samples < 0.0) return LDD_EINVAL;
    if (period_samples > 0.0 && !h->d_lnM[0]) return fail(h, LDD_EINVAL, "ldd_set_mtf_ramp needs the MTF table (LDD_F_MTF)");
    h->ramp_pos0 = pos0_sample; h->ramp_period = period_samples; h->ramp_step = step_per_period;
    h->ramp_hold_until = hold_until_sample; h->ramp_hold_level = hold_level;
    return LDD_OK;
}

int ldd_demod_range_query(ldd_handle* h, long long start, long long length, ldd_range* out) {
    if (!h || !out || length < 0 || start < 0) return LDD_EINVAL;
    const int N = h->cfg.blocklen;
    const long long S = N - h->cfg.blockcut - h->cfg.blockcut_end;
    long long end = start + length + 1;                              // lddecode_core.py:374
    long long s0 = start > h->cfg.blockcut ? start - h->cfg.blockcut : 0;   // :376-379
    out->first_sample = s0;
    out->nblocks = (end - s0 + S - 1) / S;                           // range(start, end, S), :385
    out->total_out = end - s0 + 1;                                   // :400
    out->audio1_len = 0;
    out->audio2_len = 0;
    if (h->A) {
        int ds = N / h->A;
        out->audio1_len = (end - s0) / ds + 1;                       // :417
        out->audio2_len = out->audio1_len / 4;                       // :350
    }
    out->last_needed = s0 + (out->nblocks - 1) * S + N;
    return LDD_OK;
}

static int run_demod(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                     long long first_sample, long long nblocks, long long total_out, int blockcut, long long S,
                     void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev,
                     long long audio1_len, void* stream, bool pad_last = false) {
    if (!h || !rf_dev || !planes_dev || nblocks < 0 || total_out < 0) return LDD_EINVAL;
    const ldd_config& c = h->cfg;
    const int N = c.blocklen, M = N / 2;
    const bool pal = c.system == LDD_SYSTEM_PAL;
    const int nfilt = pal ? 4 : 3;
    for (int m = 0; m < nfilt; ++m)
        if (!h->have_filter[LDD_F_VIDEO + m]) return fail(h, LDD_EINVAL, "filter %d not set", LDD_F_VIDEO + m);
    if (!h->have_filter[LDD_F_RFVIDEO]) return fail(h, LDD_EINVAL, "RFVideo filter not set");
    const bool audio = h->A > 0 && audio1_l_dev && audio1_r_dev;
    if (audio && (!h->have_filter[LDD_F_AUDIO_L] || !h->have_filter[LDD_F_AUDIO_R]))
        return fail(h, LDD_EINVAL, "audio filters not set");
    if (fmt < LDD_FMT_U8 || fmt > LDD_FMT_LDS40) return fail(h, LDD_EINVAL, "bad format %d", fmt);
    int group = fmt == LDD_FMT_R30 ? 3 : (fmt == LDD_FMT_LDS40 ? 4 : 1);
    if (rf_base % group) return fail(h, LDD_EINVAL, "rf_base must be a multiple of %d for this format", group);
    if (nblocks == 0) return LDD_OK;
    long long last_needed = first_sample + (nblocks - 1) * S + N;
    if (first_sample < rf_base || (pad_last ? last_needed - N >= rf_base + rf_len : last_needed > rf_base + rf_len))
        return fail(h, LDD_ESHORT, "capture too short: need [%lld,%lld), have [%lld,%lld)", first_sample, last_needed,
                    rf_base, rf_base + rf_len);
    for (int pidx = 0; pidx < (pal ? 5 : 4); ++pidx)
        if (!planes_dev[pidx]) return fail(h, LDD_EINVAL, "plane %d is NULL", pidx);

    const int lane = c.precision == LDD_PREC_F64 ? 0 : 1;
    const bool mixed = c.precision == LDD_PREC_MIXED;
    DemodParams p;
    memset(&p, 0, sizeof p);
    p.N = N; p.M = M; p.A = audio ? h->A : 0;
    p.blockcut = blockcut;
    p.nfilt = nfilt;
    p.plan_m = make_plan(M, h->radix_max);
    if (p.A) { p.plan_a = make_plan(p.A, h->radix_max); p.wstride_a = M / p.A; p.audio_ds = N / p.A; }
    p.rf = rf_dev; p.fmt = fmt;
    p.first_sample = first_sample - rf_base;
    p.rf_limit = rf_len;
    p.stride = S;
    p.nblocks = (int)nblocks;
    p.WM = h->d_WM[lane]; p.WN = h->d_WN[lane]; p.Hv = h->d_Hv[lane];
    p.lnM = h->d_lnM[lane];
    // the in-place float32 block (ldd_demod8k.cuh) runs when its permuted tables exist; LDD_STOCKHAM_BLOCK=1 keeps the
    // out-of-place plan (comparison runs)
    if (!getenv("LDD_STOCKHAM_BLOCK")) {
        p.HvP = h->d_HvP; p.lnMP = h->d_lnMP;
        for (int m = 0; m < nfilt; ++m) p.FP[m] = h->d_FP[m];
        for (int m = 0; m < nfilt; ++m) if (!p.FP[m]) p.HvP = nullptr;
    }
    if (h->ramp_period > 0.0 && h->d_lnM[0] && blockcut == c.blockcut) {
        // plane sample k of this launch <-> capture sample first_sample + blockcut + k
        p.mtf_pos0 = h->ramp_pos0 - (double)(first_sample + blockcut);
        p.mtf_period = h->ramp_period; p.mtf_step = h->ramp_step; p.mtf_level0 = h->mtf_level_set;
        p.mtf_hold_until = h->ramp_hold_until - (double)(first_sample + blockcut); p.mtf_hold_level = h->ramp_hold_level;
    }
    const double rel[4] = {1.0, 1.0, 0.0, 0.0};
    for (int m = 0; m < nfilt; ++m) {
        p.F[m] = h->d_F[m][lane];
        p.addc[m] = c.ire0 * (h->dc[m] - rel[m]);
    }
    p.AL = h->d_AL[lane]; p.AR = h->d_AR[lane];
    p.a_lo = c.audio_slice_lo; p.a_hi = c.audio_slice_hi;
    const double twopi = 6.283185307179586476925286766559;
    p.audio_scale = c.freq_arf / twopi;
    p.audio_lowfreq = c.audio_lowfreq;
    p.hz_per_rad = c.freq_hz / twopi;
    p.ire0 = c.ire0;
    p.sync_lo = c.sync_lo_hz; p.sync_hi = c.sync_hi_hz;
    p.sync_ref = c.ire0;
    p.fp_b0 = c.fpsync_b0; p.fp_b1 = c.fpsync_b1; p.fp_c = -c.fpsync_a1;
    for (int i = 0; i < 5; ++i) p.plane[i] = planes_dev[i];
    p.total_out = total_out;
    p.audio_l = audio1_l_dev; p.audio_r = audio1_r_dev; p.audio_total = audio1_len;
    p.scratch = h->scratch_per_cta ? h->scratch : nullptr;
    p.scratch_per_cta = h->scratch_per_cta;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (int)(nblocks < h->grid ? nblocks : h->grid);
    if (grid > 0 && !getenv("LDD_FULL_GRID")) {
        // Balanced persistent grid: the kernel takes ceil(nblocks / grid) rounds whatever the last round's fill, so the
        // smallest grid with the same number of rounds costs nothing and leaves the remaining SMs to whatever else is
        // enqueued (the refine / TBC kernels of the previous capture on their side stream, NCCL's gather kernels).
        const long long rounds = (nblocks + grid - 1) / grid;
        grid = (int)((nblocks + rounds - 1) / rounds);
    }
    if (audio) {
        // the reference leaves the unwritten tail of output_audio at zero (lddecode_core.py:417-422)
        CUDA_TRY(h, cudaMemsetAsync(audio1_l_dev, 0, (size_t)audio1_len * sizeof(double), st));
        CUDA_TRY(h, cudaMemsetAsync(audio1_r_dev, 0, (size_t)audio1_len * sizeof(double), st));
    }
#ifndef LDD_EMU
    cudaStreamAttrValue l2attr;
    if (h->l2_window) {
        memset(&l2attr, 0, sizeof l2attr);
        l2attr.accessPolicyWindow.base_ptr = h->scratch;
        l2attr.accessPolicyWindow.num_bytes = h->l2_window;
        l2attr.accessPolicyWindow.hitRatio = h->l2_ratio;
        l2attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        l2attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &l2attr);
    }
#endif
    h->last_nblocks = nblocks;
    if (mixed) {
        if ((size_t)nblocks + 1 > h->flags_cap) {
            if (h->d_flags) { cudaStreamSynchronize(st); cudaFree(h->d_flags); h->d_flags = nullptr; }
            size_t cap = (size_t)nblocks + 1024;
            CUDA_TRY(h, cudaMalloc((void**)&h->d_flags, cap * sizeof(int)));
            h->flags_cap = cap;
        }
        CUDA_TRY(h, cudaMemsetAsync(h->d_flags, 0, sizeof(int), st));
        p.flag_count = h->d_flags;
        p.flag_list = h->d_flags + 1;
        p.flag_margin = h->flag_margin;
    }
    int rc;
    h->last_fused = false;
    DemodParams q = p;
    if (mixed) {
        // float64 parameter set of the re-run: only the sync decision (demod_05 -> demod_sync) is redone -- the float32
        // planes of a flagged block are as good as those of its neighbours
        q.WM = h->d_WM[0]; q.WN = h->d_WN[0]; q.Hv = h->d_Hv[0]; q.lnM = h->d_lnM[0];
        for (int m = 0; m < nfilt; ++m) q.F[m] = h->d_F[m][0];
        q.AL = h->d_AL[0]; q.AR = h->d_AR[0];
        q.scratch = h->scratch64; q.scratch_per_cta = h->scratch64_per_cta;
        q.flag_list = nullptr; q.flag_count = nullptr; q.flag_margin = 0.0;
        q.only05 = 1; q.A = 0;
        // float64 permuted tables: the re-run takes the in-place block too (LDD_STOCKHAM_RERUN=1: the generic float64 block)
        q.HvP = nullptr; q.lnMP = nullptr;
        for (int m = 0; m < 4; ++m) q.FP[m] = nullptr;
        if (p.HvP && h->d_HvP64 && h->d_FP64_05 && !getenv("LDD_STOCKHAM_RERUN")) {
            q.HvP = h->d_HvP64; q.lnMP = h->d_lnMP64; q.FP[1] = h->d_FP64_05;
        }
    }
    if (mixed && h->d_queue && demod_mixed_fused_ok(p, h->threads, h->smem_bytes, h->sp_bytes)) {
        // one launch: dynamic block queue, flagged blocks re-run in float64 by the CTA that found them
        CUDA_TRY(h, cudaMemsetAsync(h->d_queue, 0, 2 * sizeof(int), st));
        p.flag_list = nullptr; p.flag_count = nullptr;
        int g = (int)(nblocks < h->sm_count ? nblocks : h->sm_count);
        if (!getenv("LDD_FULL_GRID") && g > 8) g -= h->spare_sms;      // a few SMs stay free for the side-stream kernels / NCCL
        rc = launch_demod_mixed(p, q, h->d_queue, g, st, h->smem_bytes);
        h->last_fused = true;
    } else {
        if (lane == 0) rc = launch_demod_f64(p, grid, h->threads, st, h->sp_bytes);
        else rc = launch_demod_f32(p, grid, h->threads, st, h->smem_bytes);
        if (mixed && rc == LDD_OK) {
            // second pass: float64 over the flagged blocks only; the list is read on the device, so there
            // is no host round trip -- an idle launch costs a few microseconds when nothing was flagged
            q.block_count = h->d_flags; q.block_list = h->d_flags + 1;
            int g64 = (int)(nblocks < h->sm_count ? nblocks : h->sm_count);
            rc = launch_demod_f64(q, g64, 512, st, h->sp_bytes);
        }
    }
#ifndef LDD_EMU
    if (h->l2_window) {
        l2attr.accessPolicyWindow.num_bytes = 0;        // later kernels on this stream are not affected
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &l2attr);
    }
#endif
    if (rc) return fail(h, rc, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return LDD_OK;
}

int ldd_mixed_stats(ldd_handle* h, long long* flagged_blocks, long long* total_blocks) {
    if (!h || !flagged_blocks || !total_blocks) return LDD_EINVAL;
    *flagged_blocks = 0;
    *total_blocks = h->last_nblocks;
    if (h->cfg.precision != LDD_PREC_MIXED || !h->d_flags) return LDD_OK;
    int n = 0;
    CUDA_TRY(h, cudaMemcpy(&n, h->last_fused ? h->d_queue + 1 : h->d_flags, sizeof(int), cudaMemcpyDeviceToHost));
    *flagged_blocks = n;
    return LDD_OK;
}

int ldd_demod_blocks(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                     long long first_sample, long long nblocks, long long total_out,
                     void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev,
                     long long audio1_len, void* stream) {
    if (!h) return LDD_EINVAL;
    const long long S = h->cfg.blocklen - h->cfg.blockcut - h->cfg.blockcut_end;
    return run_demod(h, rf_dev, fmt, rf_base, rf_len, first_sample, nblocks, total_out, h->cfg.blockcut, S,
                     planes_dev, audio1_l_dev, audio1_r_dev, audio1_len, stream);
}

}  // extern "C"

// ldd_demod_blocks whose LAST block may reach past the end of the capture (it reads zeros there): the range planner of
// ldd_pipe_launch uses it where a capture does not end on the block grid, so that the planes cover every window the
// reference can still read (its windows end at least blockcut_end samples before the end of the capture).
int ldd::demod_blocks_padded(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                             long long first_sample, long long nblocks, long long total_out,
                             void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev,
                             long long audio1_len, void* stream) {
    if (!h) return LDD_EINVAL;
    const long long S = h->cfg.blocklen - h->cfg.blockcut - h->cfg.blockcut_end;
    return run_demod(h, rf_dev, fmt, rf_base, rf_len, first_sample, nblocks, total_out, h->cfg.blockcut, S,
                     planes_dev, audio1_l_dev, audio1_r_dev, audio1_len, stream, true);
}

extern "C" {

int ldd_demodblock(ldd_handle* h, const void* rf_dev, int fmt, long long rf_len,
                   void* const* planes_dev, double* audio_l_dev, double* audio_r_dev, void* stream) {
    if (!h) return LDD_EINVAL;
    const int N = h->cfg.blocklen;
    return run_demod(h, rf_dev, fmt, 0, rf_len, 0, 1, N, 0, N, planes_dev, audio_l_dev, audio_r_dev,
                     h->A, stream);
}

int ldd_demod_range(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                    long long start, long long length,
                    void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev, void* stream) {
    ldd_range r;
    int rc = ldd_demod_range_query(h, start, length, &r);
    if (rc) return rc;
    return ldd_demod_blocks(h, rf_dev, fmt, rf_base, rf_len, r.first_sample, r.nblocks, r.total_out,
                            planes_dev, audio1_l_dev, audio1_r_dev, r.audio1_len, stream);
}

}  // extern "C"

// ---- peer memory (one process per GPU): the root rank exposes its gather buffer, the other ranks map it and let their
// TBC kernels store fields straight into it over NVLink; flags in the same buffer order producers and consumer ----------
namespace {

__global__ void peer_signal_kernel(volatile int* flag, int value) {
    __threadfence_system();                 // everything this stream wrote before is visible system-wide first
    *flag = value;
    __threadfence_system();
}

__global__ void peer_wait_kernel(const volatile int* flags, int n, int stride, int value) {
    // one thread per flag; spins until flags[i * stride] >= value
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    while (flags[(size_t)i * stride] < value) {
#ifndef LDD_EMU
        __nanosleep(200);
#endif
    }
    __threadfence_system();
}

}  // namespace

extern "C" {

int ldd_peer_alloc(size_t nbytes, void** dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64 || nbytes == 0) return LDD_EINVAL;
#ifdef LDD_EMU
    return LDD_ECUDA;
#else
    if (cudaMalloc(dev_ptr, nbytes) != cudaSuccess) return LDD_ENOMEM;
    cudaMemset(*dev_ptr, 0, nbytes);
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, *dev_ptr) != cudaSuccess) { cudaFree(*dev_ptr); *dev_ptr = nullptr; return LDD_ECUDA; }
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &hd, 64);
    return LDD_OK;
#endif
}

int ldd_peer_open(const unsigned char* handle64, void** dev_ptr) {
    if (!dev_ptr || !handle64) return LDD_EINVAL;
#ifdef LDD_EMU
    return LDD_ECUDA;
#else
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    return cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? LDD_OK : LDD_ECUDA;
#endif
}

int ldd_peer_close(void* dev_ptr) {
#ifdef LDD_EMU
    return LDD_ECUDA;
#else
    return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? LDD_OK : LDD_ECUDA;
#endif
}

int ldd_peer_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? LDD_OK : LDD_ECUDA; }

int ldd_peer_read(void* host_dst, const void* dev_src, size_t nbytes) {
    if (!host_dst || !dev_src) return LDD_EINVAL;
    return cudaMemcpy(host_dst, dev_src, nbytes, cudaMemcpyDeviceToHost) == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

int ldd_peer_copy(void* dst_dev, const void* src_dev, size_t nbytes, void* stream) {
    // one DMA transfer between two device allocations (either may be a peer mapping): the copy engines move it over
    // NVLink, no SM of either GPU is involved
    if (!dst_dev || !src_dev) return LDD_EINVAL;
    if (nbytes == 0) return LDD_OK;
#ifdef LDD_EMU
    memcpy(dst_dev, src_dev, nbytes);
    return LDD_OK;
#else
    return cudaMemcpyAsync(dst_dev, src_dev, nbytes, cudaMemcpyDefault, (cudaStream_t)stream) == cudaSuccess ? LDD_OK : LDD_ECUDA;
#endif
}

int ldd_peer_signal(int* flag_dev, int value, void* stream) {
    if (!flag_dev) return LDD_EINVAL;
    LDD_LAUNCH(peer_signal_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, (volatile int*)flag_dev, value);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

int ldd_peer_wait(const int* flags_dev, int n, int stride, int value, void* stream) {
    if (!flags_dev || n < 1 || stride < 1) return LDD_EINVAL;
    LDD_LAUNCH(peer_wait_kernel, dim3((n + 31) / 32), dim3(32), 0, (cudaStream_t)stream, (const volatile int*)flags_dev, n, stride, value);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

}  // extern "C"

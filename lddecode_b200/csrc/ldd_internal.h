// Internal declarations shared by the translation units of libldd_b200.so.
#pragma once
#include "ldd_fft.cuh"
#include "ldd_fft2.cuh"
#include "ldd_async.cuh"
#include "../../include/ldd_b200.h"

#include <string>
#include <vector>

struct ldd_handle;

namespace ldd {

// Everything the fused demodulation kernel needs, passed by value.
struct DemodParams {
    // geometry
    int N, M, A;                 // block length, N/2, audio transform length (0: audio off)
    int blockcut;                // leading samples of every block that are discarded (1024)
    int nfilt;                   // 3 NTSC (video, video05, burst), 4 PAL (+pilot)
    FftPlan plan_m, plan_a;
    int wstride_a;               // M / A
    // input
    const void* rf;              // device pointer to the capture (format fmt)
    int fmt;
    long long rf_limit;          // samples available in rf (relative to rf[0]); a block reaching past it reads zeros there
    long long first_sample;      // capture sample index of block 0
    long long stride;            // N - blockcut - blockcut_end
    int nblocks;
    // tables (Cx<T>, device)
    const void* WM;              // e^{-2 pi i k/M}, k<M
    const void* WN;              // e^{-2 pi i k/N}, k<M
    const void* Hv;              // RFVideo * MTF^level, N entries
    const void* lnM;             // log(MTF), N entries (per-block level ramp; may be NULL when mtf_period == 0)
    double mtf_pos0;             // plane sample where the ramp's origin frame starts
    double mtf_period;           // samples per frame; 0: every block uses the table as it is
    double mtf_step;             // level change per frame (-1e-4)
    double mtf_level0;           // level baked into Hv (the level of the origin frame)
    double mtf_hold_until;       // blocks centred before this plane sample use mtf_hold_level (the decode's first frame)
    double mtf_hold_level;
    const void* F[4];            // FVideo, FVideo05 (pre-rolled), FVideoBurst, FVideoPilot; k<=M; scaled 1/M
    // float32 copies in the digit-permuted order of the in-place transforms (ldd_fft2.cuh; N = 16384 only, else NULL):
    // in the block arrays' padded layout PX(p) = p + p/16, SPAN = 8704 entries per part (one bulk copy per part):
    // HvP[PX(p)] = Hv[idx(p)], HvP[SPAN + PX(p)] = Hv[M + idx(p)]; lnMP likewise; FP[m][PX(p)] = F[m][idx(p)],
    // FP[m][SPAN] = F[m][M].  (The float64 parameter set of the mixed lane's re-run: plain, index p, F[M] at index M.)
    const void* HvP;
    const void* lnMP;
    const void* FP[4];
    const void* AL;              // audio_lfilt / audio_rfilt, A entries
    const void* AR;
    double addc[4];              // added to the inverse transform of filter m before the store
    int a_lo, a_hi;              // audio_fdslice_lo = [a_lo, a_hi)
    double audio_scale;          // freq_arf / 2 pi
    double audio_lowfreq;
    // constants
    double hz_per_rad;           // freq_hz / 2 pi
    double ire0;
    double sync_lo, sync_hi;     // iretohz(-55), iretohz(-25), compared against (plane05 + sync_ref)
    double sync_ref;             // value to add to the stored demod_05 plane to get absolute Hz
    double fp_b0, fp_b1, fp_c;   // FPsync: y[n] = b0 s[n] + b1 s[n-1] + c y[n-1]
    // outputs (device)
    void* plane[5];              // demod, demod_05, demod_sync (float64!), demod_burst, demod_pilot (float32)
    long long total_out;         // length of each plane
    double* audio_l;
    double* audio_r;
    long long audio_total;
    int audio_ds;                // N / A
    // scratch for the global-memory lane
    void* scratch;
    size_t scratch_per_cta;      // bytes
    // mixed-precision lane: the float32 pass appends to flag_list the blocks that hold a demod_05 sample
    // within flag_margin Hz of a sync threshold; the float64 pass then works through block_list instead
    // of 0..nblocks-1 (both NULL otherwise)
    int* flag_list;
    int* flag_count;
    double flag_margin;
    int only05;                // re-run pass of the mixed lane: only demod_05 and the sync plane are produced
    const int* block_list;
    const int* block_count;
};

// cudaGetLastError() after a launch -> LDD_OK / LDD_ECUDA with the CUDA error text in h->err
inline int launch_status(ldd_handle* h, const char* what);

int launch_demod_f64(const DemodParams& p, int grid, int threads, cudaStream_t st, size_t sp_bytes);
int launch_demod_f32(const DemodParams& p, int grid, int threads, cudaStream_t st, size_t smem_bytes);
int demod_blocks_padded(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                        long long first_sample, long long nblocks, long long total_out,
                        void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev,
                        long long audio1_len, void* stream);
int pcm_range_launch(ldd_handle* h, const double* audio_l, const double* audio_r, long long audio_len, const double* fbase_dev,
                     const double* linelocs_dev, int ll_stride, const int* linecount_dev, const double* t0_dev,
                     const double* t1_dev, const int* nout_dev, const long long* out_off_dev, int nfields, int max_nout,
                     double lineloc_add, double scale, double line_period_us, double lfreq, double rfreq, short* out_dev,
                     int* status_dev, cudaStream_t st);
bool demod_mixed_fused_ok(const DemodParams& p, int threads, size_t smem_bytes, size_t sp_bytes);
int launch_demod_mixed(const DemodParams& pf, const DemodParams& pq, int* queue, int grid, cudaStream_t st, size_t smem_bytes);

}  // namespace ldd

struct ldd_handle {
    ldd_config cfg;
    int device;
    int sm_count;
    size_t smem_optin;
    std::string err;
    // device tables, fp64 and fp32 copies
    void* d_WM[2];
    void* d_WN[2];
    void* d_Hv[2];
    void* d_F[4][2];
    void* d_HvP = nullptr;       // permuted float32 copies (block length 16384): see DemodParams
    void* d_lnMP = nullptr;
    void* d_FP[4] = {nullptr, nullptr, nullptr, nullptr};
    void* d_HvP64 = nullptr;     // the same in float64 for the mixed lane's re-run (RF filter, log MTF, FVideo05)
    void* d_lnMP64 = nullptr;
    void* d_FP64_05 = nullptr;
    void* d_AL[2];
    void* d_AR[2];
    double dc[4];
    bool have_filter[16];
    int A;
    void* scratch;
    size_t scratch_bytes;
    size_t scratch_per_cta;
    int grid;
    int threads;      // CTA size of the demodulation kernel
    int radix_max;    // largest Stockham radix used
    size_t smem_bytes;
    size_t sp_bytes = 0;    // float64 lane: bytes of the shared-memory ping-pong partner (0: not used)
    void* scratch64 = nullptr;       // float64 scratch of the mixed lane's second pass
    size_t scratch64_per_cta = 0;
    int* d_flags = nullptr;          // [0] = count, [1..] = block indices
    int* d_queue = nullptr;          // fused mixed kernel: [0] = next block to hand out, [1] = blocks re-run in float64
    bool last_fused = false;
    int spare_sms = 4;               // SMs the fused kernel leaves to concurrent streams
    size_t flags_cap = 0;
    double flag_margin = 16.0;       // Hz
    long long last_nblocks = 0;
    bool tbc_taps_set = false;
    bool burst_taps_set = false;     // __constant__ spline taps of the burst kernels uploaded       // __constant__ FIR taps uploaded for this handle's device
    size_t l2_window = 0;   // bytes of scratch covered by a persisting-L2 access policy window
    float l2_ratio = 1.0f;
    // audio phase 2
    void* d_lpf2;     // Cx<double>[N/4]
    void* d_WNfull;   // e^{-2 pi i k/N}, k<N (double) for the phase-2 transforms
    void* d_rfbase = nullptr;   // Filters['RFVideo'] as uploaded (complex128), base of ldd_set_mtf_level
    void* d_mtf = nullptr;      // Filters['MTF'] (complex128)
    double mtf_level_set = 0.0; // level d_Hv currently holds
    void* d_lnM[2] = {nullptr, nullptr};   // log(Filters['MTF']), complex128 / complex64
    double ramp_pos0 = 0.0, ramp_period = 0.0, ramp_step = 0.0;   // ldd_set_mtf_ramp (capture coordinates)
    double ramp_hold_until = -1e300, ramp_hold_level = 0.0;
    // workspace of the peak search (grown on demand)
    void* peak_ws = nullptr;
    size_t peak_ws_bytes = 0;
    void* pilot_ws = nullptr;
    size_t pilot_ws_bytes = 0;
};

inline int ldd::launch_status(ldd_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return LDD_OK;
    if (h) h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return LDD_ECUDA;
}

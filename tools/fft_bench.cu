// Micro-benchmark + self-check of the 8192-point shared-memory transforms of the float32 demodulation lane:
//   A: the Stockham plan of ldd_fft.cuh (four out-of-place passes, a CTA barrier after each)
//   B: the in-place DIF / DIT pair of ldd_fft2.cuh (one CTA-wide stage, one half-CTA barrier, warp-local stages)
// Every CTA does `reps` round trips (two forward transforms + a 1/M scaling pass) on its own array; time per
// transform = elapsed / (2 reps).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I lddecode_b200/csrc tools/fft_bench.cu -o tools/fft_bench
//   g++ -std=c++17 -O2 -DLDD_EMU -I tests/emu -I lddecode_b200/csrc -x c++ tools/fft_bench.cu tests/emu/cuda_emu.cpp -o /tmp/fft_bench_emu
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ldd_fft2.cuh"

using namespace ldd;
typedef Cx<float> C32;

template <int PADK>
__global__ void __launch_bounds__(512, 1) bench_new(const C32* W, C32* io, int reps, int check) {
    LDD_DYN_SMEM(smem);
    C32* x = (C32*)smem;
    const int tid = threadIdx.x;
    const f2::Tw tw = f2::tw_make(W, tid);
    for (int i = tid; i < 8192; i += 512) x[f2::pix<PADK>(i)] = io[(size_t)blockIdx.x * 8192 * (check ? 3 : 1) + i];
    __syncthreads();
    for (int r = 0; r < reps; ++r) {
        f2::stage1<PADK, false>(x, tw.w1, tid);
        __syncthreads();
        f2::dif_234<PADK>(x, tw, tid);
        if (check) {
            __syncthreads();
            for (int i = tid; i < 8192; i += 512) io[(size_t)blockIdx.x * 8192 * 3 + 8192 + i] = x[f2::pix<PADK>(i)];
            __syncthreads();
        }
        // (an element-wise step on the permuted spectrum would sit here; it needs the partner warp's data)
        f2::half_sync(tid >> 8);
        f2::dit_432<PADK>(x, tw, tid);
        __syncthreads();
        f2::stage1<PADK, true>(x, tw.w1, tid);
        __syncthreads();
        for (int i = 0; i < 16; ++i) {
            C32* e = &x[f2::pix<PADK>(tid) + i * f2::pst<PADK>(512)];
            *e = scale(*e, 1.0f / 8192.0f);
        }
        __syncthreads();
    }
    for (int i = tid; i < 8192; i += 512) io[(size_t)blockIdx.x * 8192 * (check ? 3 : 1) + (check ? 2 * 8192 : 0) + i] = x[f2::pix<PADK>(i)];
}

__global__ void __launch_bounds__(512, 1) bench_old(const C32* W, C32* io, int reps, int check) {
    LDD_DYN_SMEM(smem);
    C32* a = (C32*)smem;
    C32* b = a + pspan<true>(8192);
    __shared__ C32 stw[3 * 512];
    const int tid = threadIdx.x;
    fft_tw_fill<float, 8192, 512>(stw, W, tid);
    for (int i = tid; i < 8192; i += 512) a[pidx<true>(i)] = io[(size_t)blockIdx.x * 8192 * (check ? 3 : 1) + i];
    __syncthreads();
    for (int r = 0; r < reps; ++r) {
        fft8k_run<float, true, true>(a, b, stw, tid);
        if (check) {
            for (int i = tid; i < 8192; i += 512) io[(size_t)blockIdx.x * 8192 * 3 + 8192 + i] = a[pidx<true>(i)];
            __syncthreads();
        }
        fft8k_run<float, true, true>(a, b, stw, tid);
        for (int i = 0; i < 16; ++i) {
            C32* e = &a[pidx<true>(tid) + i * pstride<true>(512)];
            *e = scale(*e, 1.0f / 8192.0f);
        }
        __syncthreads();
    }
    for (int i = tid; i < 8192; i += 512) io[(size_t)blockIdx.x * 8192 * (check ? 3 : 1) + (check ? 2 * 8192 : 0) + i] = a[pidx<true>(i)];
}

static void host_fft(std::vector<std::complex<double>>& v) {
    const size_t n = v.size();
    if (n == 1) return;
    std::vector<std::complex<double>> e(n / 2), o(n / 2);
    for (size_t i = 0; i < n / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    host_fft(e);
    host_fft(o);
    for (size_t k = 0; k < n / 2; ++k) {
        std::complex<double> t = std::polar(1.0, -2.0 * M_PI * (double)k / (double)n) * o[k];
        v[k] = e[k] + t;
        v[k + n / 2] = e[k] - t;
    }
}

template <class K>
static double run(K kern, const char* name, size_t smem, const C32* dW, bool permuted, int grid, int reps) {
    const int M = 8192;
    // ---- check (one CTA, one round trip)
    std::vector<C32> h(3 * M);
    std::vector<std::complex<double>> ref(M);
    srand(1);
    for (int i = 0; i < M; ++i) {
        h[i].x = (float)(rand() % 2001 - 1000) / 1000.f;
        h[i].y = (float)(rand() % 2001 - 1000) / 1000.f;
        ref[i] = {h[i].x, h[i].y};
    }
    host_fft(ref);
    C32* dio;
    cudaMalloc((void**)&dio, sizeof(C32) * (size_t)M * (size_t)(grid > 3 ? grid : 3));
    cudaMemcpy(dio, h.data(), sizeof(C32) * 3 * M, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LDD_LAUNCH(kern, dim3(1), dim3(512), smem, 0, dW, dio, 1, 1);
    cudaDeviceSynchronize();
    std::vector<C32> out(3 * M);
    cudaMemcpy(out.data(), dio, sizeof(C32) * 3 * M, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0, nrm = 0;
    for (int p = 0; p < M; ++p) {
        const int k = permuted ? f2::idx_of_pos(p) : p;
        std::complex<double> g(out[M + p].x, out[M + p].y);
        e1 = std::max(e1, std::abs(g - ref[k]));
        nrm = std::max(nrm, std::abs(ref[k]));
        // two forward transforms / M = x[-n]
        std::complex<double> g2(out[2 * M + p].x, out[2 * M + p].y), x2(h[(M - p) % M].x, h[(M - p) % M].y);
        e2 = std::max(e2, std::abs(g2 - x2));
    }
    printf("%-10s check: spectrum max err %.3e (max |X| %.1f), round trip max err %.3e   [%s]\n", name, e1, nrm, e2, cudaGetErrorString(cudaGetLastError()));
    // ---- time
    double per = 0;
#ifndef LDD_EMU
    for (int g = 0; g < grid; ++g) cudaMemcpy(dio + (size_t)g * M, h.data(), sizeof(C32) * M, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1e;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1e);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        LDD_LAUNCH(kern, dim3(grid), dim3(512), smem, 0, dW, dio, reps, 0);
        cudaEventRecord(e1e);
        cudaEventSynchronize(e1e);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1e);
        per = ms * 1e3 / (2.0 * reps);
        printf("%-10s grid %d reps %d: %.3f ms  -> %.3f us per transform (incl. 1/2 scaling pass)  [%s]\n", name, grid, reps, ms, per, cudaGetErrorString(cudaGetLastError()));
    }
#endif
    cudaFree(dio);
    return per;
}

int main(int argc, char** argv) {
    const int M = 8192;
    const int grid = argc > 1 ? atoi(argv[1]) : 148, reps = argc > 2 ? atoi(argv[2]) : 200;
    std::vector<C32> W(M);
    for (int k = 0; k < M; ++k) { W[k].x = (float)cos(-2.0 * M_PI * k / M); W[k].y = (float)sin(-2.0 * M_PI * k / M); }
    C32* dW;
    cudaMalloc((void**)&dW, sizeof(C32) * M);
    cudaMemcpy(dW, W.data(), sizeof(C32) * M, cudaMemcpyHostToDevice);
    run(bench_old, "stockham", 2 * sizeof(C32) * pspan<true>(M), dW, false, grid, reps);
    run(bench_new<1>, "inplace/1", sizeof(C32) * f2::span<1>(), dW, true, grid, reps);
    run(bench_new<2>, "inplace/2", sizeof(C32) * f2::span<2>(), dW, true, grid, reps);
    return 0;
}

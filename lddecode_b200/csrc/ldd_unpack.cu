// Kernel (1): 10-bit unpackers.  Bit-exact restatements of ddunpack.c:11-36 (sign-extended,
// <<6) and of the Python loaders load_packed_data_3_32 / load_packed_data_4_40
// (lddutils.py:150-229, raw 0..1023), plus int->float32 for feeding a transform.  HBM-bound:
// every thread moves whole 16-byte vectors where the layout allows it.
#include "ldd_internal.h"

namespace ldd {

__device__ inline int fetch_any(const void* rf, int fmt, long long s) {
    switch (fmt) {
        case LDD_FMT_U8: return (int)((const unsigned char*)rf)[s];
        case LDD_FMT_S16: return (int)((const short*)rf)[s];
        case LDD_FMT_U16: return (int)((const unsigned short*)rf)[s];
        case LDD_FMT_R30: {
            long long w = s / 3;
            int f = (int)(s - w * 3);
            return (int)((((const unsigned*)rf)[w] >> (10 * f)) & 0x3ffu);
        }
        default: {
            long long g = s >> 2;
            int f = (int)(s & 3);
            const unsigned char* b = (const unsigned char*)rf + g * 5;
            unsigned hi = b[f], lo = b[f + 1];
            return (int)(((hi << (2 + 2 * f)) | (lo >> (6 - 2 * f))) & 0x3ffu);
        }
    }
}

__device__ inline unsigned dd_extend(unsigned field) {
    // ddunpack.c:11-21: int16 out = (uint16)(sample & 0x3ff) - 512; out <<= 6
    int v = (int)(field & 0x3ffu) - 512;
    return (unsigned)(v << 6) & 0xffffu;
}

// 8 words -> 24 int16 per thread: two 16-byte loads, three 16-byte stores.
__global__ void __launch_bounds__(256) unpack_r30_dd_kernel(const uint32_t* __restrict__ w, size_t nwords,
                                                            int16_t* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t base = t * 8;
    if (base >= nwords) return;
    if (base + 8 <= nwords && ((((uintptr_t)w) | ((uintptr_t)out)) & 15) == 0) {
        uint4 a = ((const uint4*)w)[t * 2], b = ((const uint4*)w)[t * 2 + 1];
        unsigned ww[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        unsigned h[24];
        LDD_UNROLL
        for (int i = 0; i < 8; ++i) {
            h[3 * i] = dd_extend(ww[i]);
            h[3 * i + 1] = dd_extend(ww[i] >> 10);
            h[3 * i + 2] = dd_extend(ww[i] >> 20);
        }
        uint4* o = (uint4*)(out + base * 3);
        LDD_UNROLL
        for (int v = 0; v < 3; ++v)
            o[v] = make_uint4(h[8 * v] | (h[8 * v + 1] << 16), h[8 * v + 2] | (h[8 * v + 3] << 16),
                              h[8 * v + 4] | (h[8 * v + 5] << 16), h[8 * v + 6] | (h[8 * v + 7] << 16));
    } else {
        for (size_t i = base; i < nwords && i < base + 8; ++i) {
            unsigned x = w[i];
            out[3 * i] = (int16_t)dd_extend(x);
            out[3 * i + 1] = (int16_t)dd_extend(x >> 10);
            out[3 * i + 2] = (int16_t)dd_extend(x >> 20);
        }
    }
}

// 8 consecutive samples per thread -> one 16-byte store of uint16 (raw values).
__global__ void __launch_bounds__(256) unpack_raw_kernel(const void* __restrict__ src, int fmt, size_t first, size_t n,
                                                         uint16_t* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t base = t * 8;
    if (base >= n) return;
    unsigned v[8];
    LDD_UNROLL
    for (int i = 0; i < 8; ++i) v[i] = (base + i < n) ? (unsigned)fetch_any(src, fmt, (long long)(first + base + i)) & 0xffffu : 0u;
    if (base + 8 <= n && (((uintptr_t)out) & 15) == 0) {
        ((uint4*)out)[t] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
    } else {
        for (int i = 0; i < 8 && base + i < n; ++i) out[base + i] = (uint16_t)v[i];
    }
}

// 4 consecutive samples per thread -> one float4 store.
__global__ void __launch_bounds__(256) unpack_f32_kernel(const void* __restrict__ src, int fmt, size_t first, size_t n,
                                                         float* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t base = t * 4;
    if (base >= n) return;
    float v[4];
    LDD_UNROLL
    for (int i = 0; i < 4; ++i) v[i] = (base + i < n) ? (float)fetch_any(src, fmt, (long long)(first + base + i)) : 0.f;
    if (base + 4 <= n && (((uintptr_t)out) & 15) == 0) {
        ((float4*)out)[t] = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        for (int i = 0; i < 4 && base + i < n; ++i) out[base + i] = v[i];
    }
}

}  // namespace ldd

using namespace ldd;

extern "C" {

int ldd_unpack_r30_ddunpack(const uint32_t* words_dev, size_t nwords, int16_t* out_dev, void* stream) {
    if (!words_dev || !out_dev) return LDD_EINVAL;
    if (nwords == 0) return LDD_OK;
    size_t threads = (nwords + 7) / 8;
    unsigned grid = (unsigned)((threads + 255) / 256);
    LDD_LAUNCH(unpack_r30_dd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, words_dev, nwords, out_dev);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

int ldd_unpack_raw(const void* src_dev, int fmt, size_t first_sample, size_t n, uint16_t* out_dev, void* stream) {
    if (!src_dev || !out_dev || fmt < LDD_FMT_U8 || fmt > LDD_FMT_LDS40) return LDD_EINVAL;
    if (n == 0) return LDD_OK;
    size_t threads = (n + 7) / 8;
    unsigned grid = (unsigned)((threads + 255) / 256);
    LDD_LAUNCH(unpack_raw_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src_dev, fmt, first_sample, n, out_dev);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

int ldd_unpack_f32(const void* src_dev, int fmt, size_t first_sample, size_t n, float* out_dev, void* stream) {
    if (!src_dev || !out_dev || fmt < LDD_FMT_U8 || fmt > LDD_FMT_LDS40) return LDD_EINVAL;
    if (n == 0) return LDD_OK;
    size_t threads = (n + 3) / 4;
    unsigned grid = (unsigned)((threads + 255) / 256);
    LDD_LAUNCH(unpack_f32_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src_dev, fmt, first_sample, n, out_dev);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

}  // extern "C"

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

// ---- aligned fast paths: whole packing groups per thread, 16-byte loads and stores ---------------
// .r30: 4 words (16 B) -> 12 samples.  first % 12 == 0 at the caller.
template <class OUT>
__global__ void __launch_bounds__(256) unpack_r30_vec_kernel(const uint4* __restrict__ w, size_t nquads, OUT* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nquads) return;
    uint4 q = w[t];
    unsigned ww[4] = {q.x, q.y, q.z, q.w};
    unsigned v[12];
    LDD_UNROLL
    for (int i = 0; i < 4; ++i) { v[3 * i] = ww[i] & 0x3ffu; v[3 * i + 1] = (ww[i] >> 10) & 0x3ffu; v[3 * i + 2] = (ww[i] >> 20) & 0x3ffu; }
    if (sizeof(OUT) == 4) {
        float4* o = (float4*)out + t * 3;
        LDD_UNROLL
        for (int k = 0; k < 3; ++k) o[k] = make_float4((float)v[4 * k], (float)v[4 * k + 1], (float)v[4 * k + 2], (float)v[4 * k + 3]);
    } else {
        uint2* o = (uint2*)out + t * 3;            // 12 uint16 = 24 B = three 8-byte stores
        LDD_UNROLL
        for (int k = 0; k < 3; ++k) { uint2 u; u.x = v[4 * k] | (v[4 * k + 1] << 16); u.y = v[4 * k + 2] | (v[4 * k + 3] << 16); o[k] = u; }
    }
}

// .lds: 20 bytes (five 4-byte loads) -> 16 samples -> two 16-byte (uint16) or four 16-byte (float32)
// stores.  first % 16 == 0 at the caller, so every thread's 20 bytes start on a 4-byte boundary.
template <class OUT>
__global__ void __launch_bounds__(256) unpack_lds_vec_kernel(const unsigned* __restrict__ src, size_t ngroups, OUT* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ngroups) return;
    unsigned wd[5];
    LDD_UNROLL
    for (int i = 0; i < 5; ++i) wd[i] = src[t * 5 + i];
    auto byte_at = [&](int b) -> unsigned { return (wd[b >> 2] >> (8 * (b & 3))) & 0xffu; };
    unsigned v[16];
    LDD_UNROLL
    for (int g = 0; g < 4; ++g) {
        unsigned b0 = byte_at(5 * g), b1 = byte_at(5 * g + 1), b2 = byte_at(5 * g + 2), b3 = byte_at(5 * g + 3), b4 = byte_at(5 * g + 4);
        v[4 * g] = (b0 << 2) | (b1 >> 6);
        v[4 * g + 1] = ((b1 & 0x3fu) << 4) | (b2 >> 4);
        v[4 * g + 2] = ((b2 & 0x0fu) << 6) | (b3 >> 2);
        v[4 * g + 3] = ((b3 & 0x03u) << 8) | b4;
    }
    if (sizeof(OUT) == 4) {
        float4* o = (float4*)out + t * 4;
        LDD_UNROLL
        for (int k = 0; k < 4; ++k) o[k] = make_float4((float)v[4 * k], (float)v[4 * k + 1], (float)v[4 * k + 2], (float)v[4 * k + 3]);
    } else {
        uint4* o = (uint4*)out + t * 2;
        LDD_UNROLL
        for (int k = 0; k < 2; ++k)
            o[k] = make_uint4(v[8 * k] | (v[8 * k + 1] << 16), v[8 * k + 2] | (v[8 * k + 3] << 16),
                              v[8 * k + 4] | (v[8 * k + 5] << 16), v[8 * k + 6] | (v[8 * k + 7] << 16));
    }
}

// Runs the aligned bulk of [first, first+n) through a vector kernel; returns how many samples it covered.
template <class OUT>
static size_t unpack_bulk(const void* src, int fmt, size_t first, size_t n, OUT* out, cudaStream_t st) {
    if ((((uintptr_t)src) | ((uintptr_t)out)) & 15) return 0;
    if (fmt == LDD_FMT_R30 && first % 12 == 0) {
        size_t nq = n / 12;
        if (!nq) return 0;
        const uint4* w = (const uint4*)((const char*)src + first / 3 * 4);
        LDD_LAUNCH(unpack_r30_vec_kernel<OUT>, dim3((unsigned)((nq + 255) / 256)), dim3(256), 0, st, w, nq, out);
        return nq * 12;
    }
    if (fmt == LDD_FMT_LDS40 && first % 16 == 0) {
        size_t ng = n / 16;
        if (!ng) return 0;
        const unsigned* w = (const unsigned*)((const char*)src + first / 4 * 5);
        LDD_LAUNCH(unpack_lds_vec_kernel<OUT>, dim3((unsigned)((ng + 255) / 256)), dim3(256), 0, st, w, ng, out);
        return ng * 16;
    }
    return 0;
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
    size_t done = unpack_bulk<uint16_t>(src_dev, fmt, first_sample, n, out_dev, (cudaStream_t)stream);
    first_sample += done; n -= done; out_dev += done;
    if (n == 0) return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
    size_t threads = (n + 7) / 8;
    unsigned grid = (unsigned)((threads + 255) / 256);
    LDD_LAUNCH(unpack_raw_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src_dev, fmt, first_sample, n, out_dev);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

int ldd_unpack_f32(const void* src_dev, int fmt, size_t first_sample, size_t n, float* out_dev, void* stream) {
    if (!src_dev || !out_dev || fmt < LDD_FMT_U8 || fmt > LDD_FMT_LDS40) return LDD_EINVAL;
    if (n == 0) return LDD_OK;
    size_t done = unpack_bulk<float>(src_dev, fmt, first_sample, n, out_dev, (cudaStream_t)stream);
    first_sample += done; n -= done; out_dev += done;
    if (n == 0) return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
    size_t threads = (n + 3) / 4;
    unsigned grid = (unsigned)((threads + 255) / 256);
    LDD_LAUNCH(unpack_f32_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src_dev, fmt, first_sample, n, out_dev);
    return cudaGetLastError() == cudaSuccess ? LDD_OK : LDD_ECUDA;
}

}  // extern "C"

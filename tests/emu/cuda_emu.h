// CPU emulation of the small CUDA subset used by lddecode_b200/csrc -- TEST SCAFFOLDING ONLY.
//
// The build container has nvcc but no GPU, and a GPU run costs minutes.  To unit-test the real
// kernel sources on the CPU, tests/emu/build_emu.py compiles the very same .cu files with g++
// and -DLDD_EMU; this header then supplies threadIdx/blockIdx, __syncthreads, warp shuffles,
// atomics and a few runtime calls.  Every CUDA thread of a block is a ucontext fiber; a
// barrier yields to a round-robin scheduler, so divergent barriers deadlock loudly instead of
// passing silently.  Blocks run one after another.
//
// The emulated library (tests/emu/_build/libldd_emu.so) is loaded only by tests that pass it
// explicitly; the product package loads lddecode_b200/libldd_b200.so and nothing else.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(4) ushort2 { unsigned short x, y; };
struct alignas(8) ushort4 { unsigned short x, y, z, w; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
static inline ushort4 make_ushort4(unsigned short a, unsigned short b, unsigned short c, unsigned short d) { return {a, b, c, d}; }
static inline ushort2 make_ushort2(unsigned short a, unsigned short b) { return {a, b}; }

namespace emu {
extern uint3 threadIdx_, blockIdx_;
extern dim3 blockDim_, gridDim_;
extern char* dyn_smem;
void barrier();
void warp_barrier();
uint64_t* warp_slot(int lane);            // exchange slot of `lane` in the calling fiber's warp
int lane_id();
int warp_live_lanes();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emu

#define threadIdx (emu::threadIdx_)
#define blockIdx (emu::blockIdx_)
#define blockDim (emu::blockDim_)
#define gridDim (emu::gridDim_)
#define warpSize 32

static inline void __syncthreads() { emu::barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline int __syncthreads_or(int pred) {
    static int acc;
    if (emu::threadIdx_.x == 0 && emu::threadIdx_.y == 0 && emu::threadIdx_.z == 0) acc = 0;
    emu::barrier();
    if (pred) acc = 1;
    emu::barrier();
    int r = acc;
    emu::barrier();
    return r;
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

template <class T>
static inline T emu_xchg(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of >8 bytes");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    *emu::warp_slot(emu::lane_id()) = raw;
    emu::warp_barrier();
    uint64_t got = *emu::warp_slot(src & 31);
    emu::warp_barrier();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = emu::lane_id();
    int base = lane & ~(width - 1);
    return emu_xchg(v, base + (src & (width - 1)));
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu::lane_id();
    int base = lane & ~(width - 1);
    int src = lane - (int)d;
    return emu_xchg(v, src < base ? lane : src);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu::lane_id();
    int base = lane & ~(width - 1);
    int src = lane + (int)d;
    return emu_xchg(v, src >= base + width ? lane : src);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    return emu_xchg(v, emu::lane_id() ^ m);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned mine = pred ? 1u : 0u;
    *emu::warp_slot(emu::lane_id()) = mine;
    emu::warp_barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (*emu::warp_slot(l) & 1u) r |= (1u << l);
    emu::warp_barrier();
    // lanes that do not exist (partial warp) left zeros in their slots
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    unsigned b = __ballot_sync(m, p);
    int live = emu::warp_live_lanes();
    unsigned full = live >= 32 ? 0xffffffffu : ((1u << live) - 1u);
    return (b & full) == full;
}
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

template <class T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
template <class T>
static inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <class T>
static inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <class T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T>
static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T>
static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T>
static inline T __ldg(const T* p) { return *p; }

static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(M_PI * (double)x); *c = (float)std::cos(M_PI * (double)x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
using std::max;
using std::min;
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline double fdivide(double a, double b) { return a / b; }

// ---- runtime subset --------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; int l2CacheSize; char name[64]; };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memmove(d, s, n); return cudaSuccess; }
template <class S>
static inline cudaError_t cudaMemcpyToSymbol(S& sym, const void* src, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice) {
    std::memcpy((char*)&sym + off, src, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; p->l2CacheSize = 126 << 20;
    std::snprintf(p->name, sizeof p->name, "cpu-emulation");
    return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }

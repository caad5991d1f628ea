// Bulk asynchronous copies global -> shared memory (cp.async.bulk, the 1-D form of TMA) and the mbarrier calls that track
// them.  Used by the TBC kernel (line staging) and by the float32 demodulation block (filter tables into dead arrays).
#pragma once
#include "ldd_platform.h"

namespace ldd {
#ifndef LDD_EMU
__device__ inline unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ inline void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ inline void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ inline void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ inline void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses of shared memory issued before this point are ordered before later async-proxy ones
__device__ inline void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ inline void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif
}  // namespace ldd

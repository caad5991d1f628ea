// Build-target glue.  The product is compiled by nvcc for sm_100a.  With -DLDD_EMU the same
// sources are compiled by g++ against tests/emu/cuda_emu.h so that CPU-only unit tests can run
// the kernel code (test scaffolding; the package never loads that build).
#pragma once

#ifdef LDD_EMU
#include "cuda_emu.h"
#define LDD_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define LDD_DYN_SMEM(name) char* name = emu::dyn_smem
#define LDD_HD
#define LDD_UNROLL
#else
#include <cuda_runtime.h>
#define LDD_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LDD_DYN_SMEM(name) extern __shared__ __align__(16) char name[]
#define LDD_HD __host__ __device__
#define LDD_UNROLL _Pragma("unroll")
#endif

#include <cstdint>

// Error codes of the C-ABI (include/ldd_b200.h).
#define LDD_OK 0
#define LDD_EINVAL (-1)
#define LDD_ESHORT (-2)      // capture too short for the request: the reference returns None
#define LDD_ECUDA (-3)
#define LDD_ENOMEM (-4)
#define LDD_ECAP (-5)        // caller-provided output buffer too small

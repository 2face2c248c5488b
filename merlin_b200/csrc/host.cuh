// merlin_b200 — host-side helpers shared by all launchers: status codes, last-error text,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/merlin_b200.h"

namespace mb {

void set_error(const char* fmt, ...);

#define MB_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::mb::set_error(__VA_ARGS__);      \
      return MERLIN_E_INVALID;           \
    }                                    \
  } while (0)

#define MB_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::mb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                      __LINE__);                                                             \
      return MERLIN_E_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

// Encodes a bf16 tensor map with a 128B-swizzled box.  dims/strides are innermost-first;
// strides (bytes) are given for dims 1..rank-1.  Returns MERLIN_OK or an error code.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);
// Same for fp32 tensors (128B swizzle => inner box of at most 32 floats).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box);

int sm_count();
// SMs the persistent kernels may occupy right now (merlin_set_sm_budget; defaults to sm_count()).
int sm_budget();

}  // namespace mb

// Shared helpers for libsparch_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sparch_b200.h"

namespace sparch {

void set_error(const char* fmt, ...);
int cuda_fail(const char* what, cudaError_t e);

#define SPARCH_REQUIRE(cond, msg)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      ::sparch::set_error("%s: %s", __func__, msg);       \
      return SPARCH_ERR_ARG;                              \
    }                                                     \
  } while (0)

#define SPARCH_CUDA(call)                                               \
  do {                                                                  \
    cudaError_t _e = (call);                                            \
    if (_e != cudaSuccess) return ::sparch::cuda_fail(#call, _e);       \
  } while (0)

#define SPARCH_LAUNCH_OK()                                              \
  do {                                                                  \
    cudaError_t _e = cudaGetLastError();                                \
    if (_e != cudaSuccess) return ::sparch::cuda_fail(__func__, _e);    \
  } while (0)

inline cudaStream_t as_stream(sparch_stream_t st) { return reinterpret_cast<cudaStream_t>(st); }

int sm_count();
long long* recur_debug_buffer();  // profiling aid set by sparch_recur_debug_clocks (recur.cu), normally NULL
int recur_debug_flags();

// Surrogate window and threshold tests on v = u - theta, exactly as the reference evaluates them
// in fp32 (snns.py:29, 33-35): spike iff v > 0; gradient passes iff -0.5 < v <= 0.5.
__device__ __forceinline__ float spike_of(float v) { return v > 0.0f ? 1.0f : 0.0f; }
__device__ __forceinline__ bool window_of(float v) { return v > -0.5f && v <= 0.5f; }

}  // namespace sparch

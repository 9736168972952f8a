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

// Device-dependent one-time work (cudaFuncSetAttribute, SM count) is cached PER DEVICE: a process may drive several
// GPUs (a model on cuda:1 while cuda:0 is current), and a per-process flag would leave the second device without
// its > 48 KB dynamic shared memory attribute.
constexpr int SPARCH_MAX_DEVICES = 64;
int current_device();  // cudaGetDevice(), 0 when it fails
int sm_count();        // of the current device
struct PerDeviceOnce {
  bool done[SPARCH_MAX_DEVICES] = {};
  bool first() {       // true exactly once per device (always true for an out-of-range ordinal)
    const int d = current_device();
    if (d < 0 || d >= SPARCH_MAX_DEVICES) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
long long* recur_debug_buffer();  // profiling aid set by sparch_recur_debug_clocks (recur.cu), normally NULL
int recur_debug_flags();

// Surrogate window and threshold tests on v = u - theta, exactly as the reference evaluates them
// in fp32 (snns.py:29, 33-35): spike iff v > 0; gradient passes iff -0.5 < v <= 0.5.
__device__ __forceinline__ float spike_of(float v) { return v > 0.0f ? 1.0f : 0.0f; }
__device__ __forceinline__ bool window_of(float v) { return v > -0.5f && v <= 0.5f; }


// Power-of-two scale of an fp16 hi/lo split: brings max|x| into [2^12, 2^13) (fp16 keeps 11 bits from
// there down to 2^-14, the lo term another 11 below the hi term's last bit).  The same function gives the
// split kernel its scale and the GEMM epilogue the inverse, from the bit pattern of max|x|.
__device__ __forceinline__ int f16_scale_exp(uint32_t amax_bits) {
  const int E = (int)((amax_bits >> 23) & 0xffu);          // biased exponent of max|x|
  if (amax_bits == 0u || E == 0xff) return 0;               // all zero / inf / nan: leave unscaled
  int e = E - 126;                                          // max|x| = f * 2^e, f in [0.5, 1)
  if (e < -60) e = -60;
  return 13 - e;
}

}  // namespace sparch

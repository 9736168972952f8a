// ReadoutLayer cell (snns.py:807-825): non-spiking leaky integrator whose output is the sum over
// time of softmax(u_t) across the class dimension, and its reverse pass.
// One block per batch row, one thread per class; the softmax reductions are block-wide.
#include "common.cuh"

namespace sparch {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide reduction for blockDim.x <= 1024.  `sh` holds 32 floats; every thread gets the result.
template <bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  if (nw == 1) return v;
  __syncthreads();  // previous use of sh is finished
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float r = lane < nw ? sh[lane] : (IS_MAX ? -INFINITY : 0.0f);
  return IS_MAX ? warp_max(r) : warp_sum(r);
}

__global__ void readout_fwd_kernel(const float* __restrict__ Z, const float* __restrict__ scale,
                                   const float* __restrict__ shift, const float* __restrict__ alpha,
                                   const float* __restrict__ u0, float* __restrict__ out,
                                   float* __restrict__ U, int T, int C) {
  __shared__ float sh[32];
  const int c = threadIdx.x;
  const bool live = c < C;
  const int64_t b = blockIdx.x;
  const float al = live ? alpha[c] : 0.f, oma = __fsub_rn(1.0f, al);
  const float sc = (live && scale) ? scale[c] : 1.0f, sf = (live && shift) ? shift[c] : 0.0f;
  const bool affine = scale != nullptr;
  float u = live ? u0[b * C + c] : 0.f, acc = 0.f;
  const int64_t base = b * (int64_t)T * C + c;
  float znext = (live && T > 0) ? Z[base] : 0.f;
  for (int t = 0; t < T; ++t) {
    float z = znext;
    if (live && t + 1 < T) znext = Z[base + (int64_t)(t + 1) * C];
    float cur = affine ? __fmaf_rn(z, sc, sf) : z;
    u = __fadd_rn(__fmul_rn(al, u), __fmul_rn(oma, cur));  // snns.py:822
    float m = block_reduce<true>(live ? u : -INFINITY, sh);
    float e = live ? expf(u - m) : 0.f;
    float den = block_reduce<false>(e, sh);
    if (live) {
      acc += e / den;  // snns.py:823
      U[base + (int64_t)t * C] = u;
    }
  }
  if (live) out[b * C + c] = acc;
}

__global__ void readout_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ U,
                                   const float* __restrict__ alpha, const float* __restrict__ u0,
                                   float* __restrict__ dI, float* __restrict__ p_alpha, int T, int C) {
  __shared__ float sh[32];
  const int c = threadIdx.x;
  const bool live = c < C;
  const int64_t b = blockIdx.x;
  const float al = live ? alpha[c] : 0.f, oma = __fsub_rn(1.0f, al), inv_oma = 1.0f / oma;
  const float go = live ? gout[b * C + c] : 0.f;
  const int64_t base = b * (int64_t)T * C + c;
  float du = 0.f, pa = 0.f;
  float u_t = (live && T > 0) ? U[base + (int64_t)(T - 1) * C] : 0.f;
  for (int t = T - 1; t >= 0; --t) {
    float u_prev = 0.f;
    if (live) u_prev = t > 0 ? U[base + (int64_t)(t - 1) * C] : u0[b * C + c];
    float m = block_reduce<true>(live ? u_t : -INFINITY, sh);
    float e = live ? expf(u_t - m) : 0.f;
    float den = block_reduce<false>(e, sh);
    float pr = e / den;
    float dot = block_reduce<false>(pr * go, sh);
    float du_t = pr * (go - dot) + al * du;
    if (live) {
      dI[base + (int64_t)t * C] = oma * du_t;
      pa += du_t * ((u_prev - u_t) * inv_oma);  // d u_t / d alpha = u_{t-1} - I_t
    }
    du = du_t;
    u_t = u_prev;
  }
  if (live) p_alpha[b * C + c] = pa;
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_readout_fwd(const float* Z, const float* scale, const float* shift, const float* alpha,
                       const float* u0, float* out, float* U, int B, int T, int C, sparch_stream_t st) {
  SPARCH_REQUIRE(B >= 0 && T >= 0 && C > 0 && C <= 1024, "bad shape (classes must be 1..1024)");
  SPARCH_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  if (B == 0) return SPARCH_OK;
  SPARCH_REQUIRE(alpha && u0 && out && (T == 0 || (Z && U)), "null pointer");
  int threads = ((C + 31) / 32) * 32;
  readout_fwd_kernel<<<B, threads, 0, as_stream(st)>>>(Z, scale, shift, alpha, u0, out, U, T, C);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_readout_bwd(const float* gout, const float* U, const float* alpha, const float* u0,
                       float* dI, float* p_alpha, int B, int T, int C, sparch_stream_t st) {
  SPARCH_REQUIRE(B >= 0 && T >= 0 && C > 0 && C <= 1024, "bad shape (classes must be 1..1024)");
  if (B == 0) return SPARCH_OK;
  SPARCH_REQUIRE(gout && alpha && u0 && p_alpha && (T == 0 || (U && dI)), "null pointer");
  int threads = ((C + 31) / 32) * 32;
  readout_bwd_kernel<<<B, threads, 0, as_stream(st)>>>(gout, U, alpha, u0, dI, p_alpha, T, C);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

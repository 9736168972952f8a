// Train-step glue: the Adam update of exp.py:89 (torch.optim.Adam defaults: no weight decay, no
// amsgrad) for ALL parameter tensors in one launch.  torch's fused multi-tensor Adam walks 64 K-element chunks with
// one block each (~50 blocks for the 3.2 M parameters of cfg 4); here the tensors form one virtual index space cut
// into 1024-element blocks, so the update runs at memory speed.  The step count lives on the device (graph replay).
#include "common.cuh"

namespace sparch {

constexpr int ADAM_MAX_TENSORS = 48;
constexpr int ADAM_BLOCK_ELEMS = 1024;

struct AdamTable {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long n[ADAM_MAX_TENSORS];
  int first_block[ADAM_MAX_TENSORS + 1];   // block index where tensor i starts
  int count;
};

__global__ void __launch_bounds__(256)
adam_kernel(const AdamTable tb, const long long* __restrict__ step, const float* __restrict__ hyper) {
  // hyper-parameters live on the device (lr, beta1, beta2, eps, gradient scale): a learning-rate scheduler
  // (exp.py:92-96, ReduceLROnPlateau) changes them between replays of a captured step without re-capturing
  const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], gscale = hyper[4];
  // tensor of this block: binary search over first_block
  int lo = 0, hi = tb.count;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tb.first_block[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const long long base = (long long)(blockIdx.x - tb.first_block[lo]) * ADAM_BLOCK_ELEMS;
  const long long n = tb.n[lo];
  float* __restrict__ p = tb.p[lo];
  const float* __restrict__ g = tb.g[lo];
  float* __restrict__ m = tb.m[lo];
  float* __restrict__ v = tb.v[lo];
  const float t = (float)(*step);
  const float bc1 = 1.0f - powf(beta1, t), bc2s = sqrtf(1.0f - powf(beta2, t));
  const float step_size = lr / bc1;
  // whole 1024-element blocks of 16-byte aligned tensors: one float4 per thread and array (7 x 16 bytes per thread
  // in flight instead of 7 x 4 four times over)
  const bool v4 = base + ADAM_BLOCK_ELEMS <= n &&
                  ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (v4) {
    const long long i = base + 4 * threadIdx.x;
    const float4 g4 = *reinterpret_cast<const float4*>(g + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i), v4v = *reinterpret_cast<const float4*>(v + i);
    float4 p4 = *reinterpret_cast<const float4*>(p + i);
    const float gg[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4v.x, v4v.y, v4v.z, v4v.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {                 // the scalar path's expressions, operation for operation
      mm[e] = mm[e] + (1.0f - beta1) * (gg[e] - mm[e]);
      vv[e] = beta2 * vv[e] + (1.0f - beta2) * gg[e] * gg[e];
      pp[e] -= step_size * mm[e] / (sqrtf(vv[e]) / bc2s + eps);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    return;
  }
#pragma unroll
  for (int k = 0; k < ADAM_BLOCK_ELEMS / 256; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i < n) {
      const float gi = g[i] * gscale;    // gscale = 1 / world size when the gradients hold an all-reduced SUM
      const float mi = m[i] + (1.0f - beta1) * (gi - m[i]);           // lerp, as ATen's fused kernel
      const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step_size * mi / (sqrtf(vi) / bc2s + eps);
    }
  }
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_adam_step(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                     float* const* exp_avg_sq, const int64_t* numel, const int64_t* step, const float* hyper,
                     sparch_stream_t st) {
  SPARCH_REQUIRE(count >= 0 && count <= ADAM_MAX_TENSORS, "at most 48 tensors per call");
  if (count == 0) return SPARCH_OK;
  SPARCH_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && step && hyper, "null pointer");
  AdamTable tb;
  int blocks = 0, used = 0;
  for (int i = 0; i < count; ++i) {
    if (numel[i] <= 0) continue;
    SPARCH_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "null tensor pointer");
    tb.p[used] = params[i]; tb.g[used] = grads[i]; tb.m[used] = exp_avg[i]; tb.v[used] = exp_avg_sq[i];
    tb.n[used] = numel[i];
    tb.first_block[used] = blocks;
    blocks += (int)((numel[i] + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS);
    ++used;
  }
  tb.first_block[used] = blocks;
  tb.count = used;
  if (blocks == 0) return SPARCH_OK;
  adam_kernel<<<blocks, 256, 0, as_stream(st)>>>(tb, reinterpret_cast<const long long*>(step), hyper);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

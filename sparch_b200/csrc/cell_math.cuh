// Per-neuron update and its adjoint, shared by the streaming and the tensor-core recurrence kernels.
#pragma once
#include "common.cuh"

namespace sparch {

struct NeuronParams {
  float alpha, oma, beta, a, b;
};

template <bool ADAPT>
__device__ __forceinline__ NeuronParams load_params(const float* __restrict__ alpha,
                                                    const float* __restrict__ beta,
                                                    const float* __restrict__ a,
                                                    const float* __restrict__ b, int h) {
  NeuronParams p;
  p.alpha = alpha[h];
  p.oma = __fsub_rn(1.0f, p.alpha);
  if (ADAPT) {
    p.beta = beta[h];
    p.a = a[h];
    p.b = b[h];
  } else {
    p.beta = p.a = p.b = 0.f;
  }
  return p;
}

template <bool ADAPT>
__device__ __forceinline__ void step_fwd(const NeuronParams& p, float cur, float theta, float& u,
                                         float& w, float& s) {
  float x = cur;
  if (ADAPT) {
    w = __fadd_rn(__fadd_rn(__fmul_rn(p.beta, w), __fmul_rn(p.a, u)), __fmul_rn(p.b, s));
    x = __fsub_rn(x, w);
  }
  u = __fadd_rn(__fmul_rn(p.alpha, __fsub_rn(u, s)), __fmul_rn(p.oma, x));
  s = spike_of(__fsub_rn(u, theta));
}

// one reverse step; returns dI_t.  du/dw are the adjoints carried from t+1 and are updated to t.
template <bool ADAPT>
__device__ __forceinline__ float step_bwd(const NeuronParams& p, float inv_oma, float theta, float g,
                                          float recb, float u_t, float u_prev, float s_prev,
                                          float w_prev, float& du, float& dw, float& pa, float& pb,
                                          float& pc, float& pd) {
  float ds = g - p.alpha * du + recb;
  if (ADAPT) ds += p.b * dw;
  float du_t = (window_of(__fsub_rn(u_t, theta)) ? ds : 0.0f) + p.alpha * du;
  if (ADAPT) du_t += p.a * dw;
  float dI = p.oma * du_t;
  float d = u_prev - s_prev;
  pa += du_t * ((d - u_t) * inv_oma);
  if (ADAPT) {
    float dw_t = p.beta * dw - dI;
    pb += dw_t * w_prev;
    pc += dw_t * u_prev;
    pd += dw_t * s_prev;
    dw = dw_t;
  }
  du = du_t;
  return dI;
}

}  // namespace sparch

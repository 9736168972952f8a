// Membrane recurrence of the four sparch neuron kinds and its reverse-time adjoint.
//
// Forward update per (b, h), in the reference's exact fp32 operation order
// (snns.py:297, 438-439, 572, 718-721) -- every product and sum is rounded separately
// (__fmul_rn/__fadd_rn forbid FMA contraction) so that, given the same input current,
// the membrane trajectory and therefore the spike train equals ATen's bit for bit:
//     w_t = (beta*w + a*u) + b*s                         (adaptive kinds)
//     x_t = (I_t [+ rec_t]) [- w_t]
//     u_t = alpha*(u - s) + (1-alpha)*x_t
//     s_t = (u_t - theta) > 0
// Reverse pass: the equations of SURVEY.md 8a (verified against autograd by the CPU
// tests).  The input-current slope term uses d - x_t = (d - u_t)/(1-alpha), d = u_{t-1}-s_{t-1},
// so the tape only has to hold U (and W): spikes and surrogate windows are re-derived from
// U with the same fp32 comparisons the forward made.
#include "cell_math.cuh"
#include "common.cuh"

namespace sparch {

// ------------------------------------------------------------------ streaming forward
// One thread per (b, h), h fastest so a warp reads 128 contiguous bytes per timestep.  The input
// current does not depend on the state, so TU timesteps of it are loaded ahead of the dependent
// chain; state (u, w, s) never leaves registers.
constexpr int TU = 8;

template <bool ADAPT>
__global__ void __launch_bounds__(128)
cell_fwd_stream_kernel(const float* __restrict__ Z, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ alpha,
                       const float* __restrict__ beta, const float* __restrict__ a,
                       const float* __restrict__ b, const float* __restrict__ u0,
                       const float* __restrict__ w0, const float* __restrict__ s0, float theta,
                       float* __restrict__ S, float* __restrict__ U, float* __restrict__ W, int Be,
                       int T, int H) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Be * H) return;
  int h = (int)(idx % H);
  int64_t bi = idx / H;
  const NeuronParams p = load_params<ADAPT>(alpha, beta, a, b, h);
  const float sc = scale ? scale[h] : 1.0f, sf = shift ? shift[h] : 0.0f;
  const bool affine = scale != nullptr;
  float u = u0[idx], s = s0[idx], w = ADAPT ? w0[idx] : 0.0f;
  int64_t base = bi * (int64_t)T * H + h;
  // Software pipeline: the loads of chunk c+1 are in flight while chunk c's dependent chain runs.
  float zn[TU];
#pragma unroll
  for (int k = 0; k < TU; ++k) zn[k] = (k < T) ? __ldcs(&Z[base + (int64_t)k * H]) : 0.0f;
  for (int t0 = 0; t0 < T; t0 += TU) {
    float z[TU];
#pragma unroll
    for (int k = 0; k < TU; ++k) z[k] = zn[k];
#pragma unroll
    for (int k = 0; k < TU; ++k)
      zn[k] = (t0 + TU + k < T) ? __ldcs(&Z[base + (int64_t)(t0 + TU + k) * H]) : 0.0f;
#pragma unroll
    for (int k = 0; k < TU; ++k) {
      if (t0 + k < T) {
        float cur = affine ? __fmaf_rn(z[k], sc, sf) : z[k];
        step_fwd<ADAPT>(p, cur, theta, u, w, s);
        int64_t o = base + (int64_t)(t0 + k) * H;
        __stcs(&S[o], s);
        __stcs(&U[o], u);
        if (ADAPT) __stcs(&W[o], w);
      }
    }
  }
}

// ------------------------------------------------------------------ single step (any kind)
template <bool ADAPT>
__global__ void cell_step_fwd_kernel(int t, const float* __restrict__ Z,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ alpha, const float* __restrict__ beta,
                                     const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ rec, const float* __restrict__ u0,
                                     const float* __restrict__ w0, const float* __restrict__ s0,
                                     float theta, float* __restrict__ S, float* __restrict__ U,
                                     float* __restrict__ W, int Be, int T, int H) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Be * H) return;
  int h = (int)(idx % H);
  int64_t bi = idx / H;
  const NeuronParams p = load_params<ADAPT>(alpha, beta, a, b, h);
  int64_t o = (bi * T + t) * (int64_t)H + h;
  float u, w = 0.f, s;
  if (t == 0) {
    u = u0[idx];
    s = s0[idx];
    if (ADAPT) w = w0[idx];
  } else {
    u = U[o - H];
    s = S[o - H];
    if (ADAPT) w = W[o - H];
  }
  float cur = Z[o];
  if (scale) cur = __fmaf_rn(cur, scale[h], shift[h]);
  if (rec) cur = __fadd_rn(cur, rec[idx]);
  step_fwd<ADAPT>(p, cur, theta, u, w, s);
  S[o] = s;
  U[o] = u;
  if (ADAPT) W[o] = w;
}

// ------------------------------------------------------------------ streaming backward
template <bool ADAPT>
__global__ void __launch_bounds__(128)
cell_bwd_stream_kernel(const float* __restrict__ G, const float* __restrict__ U,
                       const float* __restrict__ W, const float* __restrict__ alpha,
                       const float* __restrict__ beta, const float* __restrict__ a,
                       const float* __restrict__ b, const float* __restrict__ u0,
                       const float* __restrict__ w0, const float* __restrict__ s0, float theta,
                       float* __restrict__ dI, float* __restrict__ p_alpha,
                       float* __restrict__ p_beta, float* __restrict__ p_a, float* __restrict__ p_b,
                       int Be, int T, int H) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Be * H) return;
  int h = (int)(idx % H);
  int64_t bi = idx / H;
  const NeuronParams p = load_params<ADAPT>(alpha, beta, a, b, h);
  const float inv_oma = 1.0f / p.oma;
  int64_t base = bi * (int64_t)T * H + h;
  float du = 0.f, dw = 0.f, pa = 0.f, pb = 0.f, pc = 0.f, pd = 0.f;
  float u_t = T > 0 ? U[base + (int64_t)(T - 1) * H] : 0.f;
  // Software pipeline: the tape reads of the next (earlier) chunk are in flight while this chunk's
  // dependent chain runs.
  auto fetch = [&](int t1, float (&g)[TU], float (&up)[TU], float (&wp)[TU]) {
#pragma unroll
    for (int k = 0; k < TU; ++k) {
      int t = t1 - k;
      g[k] = up[k] = wp[k] = 0.f;
      if (t >= 0) {
        g[k] = __ldcs(&G[base + (int64_t)t * H]);
        if (t > 0) {
          up[k] = __ldcs(&U[base + (int64_t)(t - 1) * H]);
          if (ADAPT) wp[k] = __ldcs(&W[base + (int64_t)(t - 1) * H]);
        } else {
          up[k] = u0[idx];
          if (ADAPT) wp[k] = w0[idx];
        }
      }
    }
  };
  float gn[TU], upn[TU], wpn[TU];
  fetch(T - 1, gn, upn, wpn);
  for (int t1 = T - 1; t1 >= 0; t1 -= TU) {
    float g[TU], up[TU], wp[TU];
#pragma unroll
    for (int k = 0; k < TU; ++k) {
      g[k] = gn[k];
      up[k] = upn[k];
      wp[k] = wpn[k];
    }
    if (t1 - TU >= 0) fetch(t1 - TU, gn, upn, wpn);
#pragma unroll
    for (int k = 0; k < TU; ++k) {
      int t = t1 - k;
      if (t >= 0) {
        float s_prev = t > 0 ? spike_of(__fsub_rn(up[k], theta)) : s0[idx];
        float d = step_bwd<ADAPT>(p, inv_oma, theta, g[k], 0.0f, u_t, up[k], s_prev, wp[k], du, dw,
                                  pa, pb, pc, pd);
        __stcs(&dI[base + (int64_t)t * H], d);
        u_t = up[k];
      }
    }
  }
  p_alpha[idx] = pa;
  if (ADAPT) {
    p_beta[idx] = pb;
    p_a[idx] = pc;
    p_b[idx] = pd;
  }
}

template <bool ADAPT>
__global__ void cell_step_bwd_kernel(int t, const float* __restrict__ G, const float* __restrict__ U,
                                     const float* __restrict__ W, const float* __restrict__ alpha,
                                     const float* __restrict__ beta, const float* __restrict__ a,
                                     const float* __restrict__ b, const float* __restrict__ recb,
                                     const float* __restrict__ u0, const float* __restrict__ w0,
                                     const float* __restrict__ s0, float theta,
                                     float* __restrict__ dI, float* __restrict__ du_next,
                                     float* __restrict__ dw_next, float* __restrict__ p_alpha,
                                     float* __restrict__ p_beta, float* __restrict__ p_a,
                                     float* __restrict__ p_b, int Be, int T, int H) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Be * H) return;
  int h = (int)(idx % H);
  int64_t bi = idx / H;
  const NeuronParams p = load_params<ADAPT>(alpha, beta, a, b, h);
  const float inv_oma = 1.0f / p.oma;
  int64_t o = (bi * T + t) * (int64_t)H + h;
  float u_t = U[o], u_prev, s_prev, w_prev = 0.f;
  if (t > 0) {
    u_prev = U[o - H];
    s_prev = spike_of(__fsub_rn(u_prev, theta));
    if (ADAPT) w_prev = W[o - H];
  } else {
    u_prev = u0[idx];
    s_prev = s0[idx];
    if (ADAPT) w_prev = w0[idx];
  }
  float du = du_next[idx], dw = ADAPT ? dw_next[idx] : 0.f;
  float pa = 0.f, pb = 0.f, pc = 0.f, pd = 0.f;
  float d = step_bwd<ADAPT>(p, inv_oma, theta, G[o], recb ? recb[idx] : 0.0f, u_t, u_prev, s_prev,
                            w_prev, du, dw, pa, pb, pc, pd);
  dI[o] = d;
  du_next[idx] = du;
  p_alpha[idx] += pa;
  if (ADAPT) {
    dw_next[idx] = dw;
    p_beta[idx] += pb;
    p_a[idx] += pc;
    p_b[idx] += pd;
  }
}

static int check_cell_args(int kind, int Be, int T, int H) {
  if (kind < 0 || kind > 3) {
    set_error("kind must be 0..3 (LIF, adLIF, RLIF, RadLIF), got %d", kind);
    return SPARCH_ERR_ARG;
  }
  if (Be < 0 || T < 0 || H <= 0) {
    set_error("bad shape Be=%d T=%d H=%d", Be, T, H);
    return SPARCH_ERR_ARG;
  }
  return SPARCH_OK;
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_cell_fwd(int kind, const float* Z, const float* scale, const float* shift,
                    const float* alpha, const float* beta, const float* a, const float* b,
                    const float* u0, const float* w0, const float* s0, float theta, float* S,
                    float* U, float* W, int Be, int T, int H, sparch_stream_t st) {
  if (int e = check_cell_args(kind, Be, T, H)) return e;
  SPARCH_REQUIRE(!(kind & 2), "recurrent kinds go through sparch_cell_step_fwd");
  SPARCH_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  int64_t n = (int64_t)Be * H;
  if (n == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(Z && alpha && u0 && s0 && S && U, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (beta && a && b && w0 && W), "adaptive kind needs beta, a, b, w0, W");
  unsigned grid = (unsigned)((n + 127) / 128);
  if (adapt)
    cell_fwd_stream_kernel<true><<<grid, 128, 0, as_stream(st)>>>(Z, scale, shift, alpha, beta, a, b, u0,
                                                                 w0, s0, theta, S, U, W, Be, T, H);
  else
    cell_fwd_stream_kernel<false><<<grid, 128, 0, as_stream(st)>>>(Z, scale, shift, alpha, beta, a, b, u0,
                                                                  w0, s0, theta, S, U, W, Be, T, H);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_cell_step_fwd(int kind, int t, const float* Z, const float* scale, const float* shift,
                         const float* alpha, const float* beta, const float* a, const float* b,
                         const float* rec, const float* u0, const float* w0, const float* s0,
                         float theta, float* S, float* U, float* W, int Be, int T, int H,
                         sparch_stream_t st) {
  if (int e = check_cell_args(kind, Be, T, H)) return e;
  SPARCH_REQUIRE(t >= 0 && t < T, "t out of range");
  SPARCH_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  int64_t n = (int64_t)Be * H;
  if (n == 0) return SPARCH_OK;
  SPARCH_REQUIRE(Z && alpha && u0 && s0 && S && U, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (beta && a && b && w0 && W), "adaptive kind needs beta, a, b, w0, W");
  SPARCH_REQUIRE(((kind & 2) != 0) == (rec != nullptr), "rec is required for (only for) recurrent kinds");
  unsigned grid = (unsigned)((n + 255) / 256);
  if (adapt)
    cell_step_fwd_kernel<true><<<grid, 256, 0, as_stream(st)>>>(t, Z, scale, shift, alpha, beta, a, b, rec,
                                                               u0, w0, s0, theta, S, U, W, Be, T, H);
  else
    cell_step_fwd_kernel<false><<<grid, 256, 0, as_stream(st)>>>(t, Z, scale, shift, alpha, beta, a, b, rec,
                                                                u0, w0, s0, theta, S, U, W, Be, T, H);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_cell_bwd(int kind, const float* G, const float* U, const float* W, const float* alpha,
                    const float* beta, const float* a, const float* b, const float* u0,
                    const float* w0, const float* s0, float theta, float* dI, float* p_alpha,
                    float* p_beta, float* p_a, float* p_b, int Be, int T, int H, sparch_stream_t st) {
  if (int e = check_cell_args(kind, Be, T, H)) return e;
  SPARCH_REQUIRE(!(kind & 2), "recurrent kinds go through sparch_cell_step_bwd");
  int64_t n = (int64_t)Be * H;
  if (n == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && U && alpha && u0 && s0 && dI && p_alpha, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (W && beta && a && b && w0 && p_beta && p_a && p_b),
                 "adaptive kind needs W, beta, a, b, w0 and the partial buffers");
  unsigned grid = (unsigned)((n + 127) / 128);
  if (adapt)
    cell_bwd_stream_kernel<true><<<grid, 128, 0, as_stream(st)>>>(G, U, W, alpha, beta, a, b, u0, w0, s0,
                                                                 theta, dI, p_alpha, p_beta, p_a, p_b,
                                                                 Be, T, H);
  else
    cell_bwd_stream_kernel<false><<<grid, 128, 0, as_stream(st)>>>(G, U, W, alpha, beta, a, b, u0, w0, s0,
                                                                  theta, dI, p_alpha, p_beta, p_a, p_b,
                                                                  Be, T, H);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_cell_step_bwd(int kind, int t, const float* G, const float* U, const float* W,
                         const float* alpha, const float* beta, const float* a, const float* b,
                         const float* recb, const float* u0, const float* w0, const float* s0,
                         float theta, float* dI, float* du_next, float* dw_next, float* p_alpha,
                         float* p_beta, float* p_a, float* p_b, int Be, int T, int H,
                         sparch_stream_t st) {
  if (int e = check_cell_args(kind, Be, T, H)) return e;
  SPARCH_REQUIRE(t >= 0 && t < T, "t out of range");
  int64_t n = (int64_t)Be * H;
  if (n == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && U && alpha && u0 && s0 && dI && du_next && p_alpha, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (W && beta && a && b && w0 && dw_next && p_beta && p_a && p_b),
                 "adaptive kind needs W, beta, a, b, w0, dw_next and the partial buffers");
  unsigned grid = (unsigned)((n + 255) / 256);
  if (adapt)
    cell_step_bwd_kernel<true><<<grid, 256, 0, as_stream(st)>>>(t, G, U, W, alpha, beta, a, b, recb, u0, w0,
                                                               s0, theta, dI, du_next, dw_next, p_alpha,
                                                               p_beta, p_a, p_b, Be, T, H);
  else
    cell_step_bwd_kernel<false><<<grid, 256, 0, as_stream(st)>>>(t, G, U, W, alpha, beta, a, b, recb, u0, w0,
                                                                s0, theta, dI, du_next, dw_next, p_alpha,
                                                                p_beta, p_a, p_b, Be, T, H);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

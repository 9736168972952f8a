// ReadoutLayer cell (snns.py:807-825): non-spiking leaky integrator whose output is the sum over
// time of softmax(u_t) across the class dimension, and its reverse pass.
// One block per batch row; see the kernels for the decomposition.
#include <stdlib.h>

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

// Block-wide copy of n floats with several independent 16-byte accesses in flight per thread (a one-load-per-
// iteration loop leaves a 128-thread block waiting a full memory latency per 512 bytes).
__device__ __forceinline__ void block_copy(float* __restrict__ dst, const float* __restrict__ src, int n) {
  const bool vec = ((n & 3) == 0) && (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0);
  if (vec) {
    const int n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * blockDim.x;
        if (i < n4) v[k] = s4[i];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * blockDim.x;
        if (i < n4) d4[i] = v[k];
      }
    }
  } else {
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * blockDim.x;
        if (i < n) v[k] = src[i];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * blockDim.x;
        if (i < n) dst[i] = v[k];
      }
    }
  }
}

// The membrane u_t is a per-(b, c) linear scan (no coupling across classes), only the softmax couples
// the classes of one step.  So a block (one batch row) works on chunks of TCH timesteps held in shared
// memory: (1) coalesced load of the chunk's inputs, (2) the dependent chain u_t per class thread --
// T fused multiply-adds, no reduction inside the chain, (3) one warp per timestep for the softmax
// (shuffle reductions), (4) per class the sum over the chunk's steps in time order (snns.py:823 adds
// them in that order).  The first version did two block-wide reductions inside every dependent step.
__global__ void __launch_bounds__(1024) readout_fwd_kernel(const float* __restrict__ Z, const float* __restrict__ scale,
                                   const float* __restrict__ shift, const float* __restrict__ alpha,
                                   const float* __restrict__ u0, float* __restrict__ out,
                                   float* __restrict__ U, int T, int C, int TCH) {
  extern __shared__ float su[];  // [TCH][C]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const bool live = tid < C;
  const int64_t b = blockIdx.x;
  const float al = live ? alpha[tid] : 0.f, oma = __fsub_rn(1.0f, al);
  const float sc = (live && scale) ? scale[tid] : 1.0f, sf = (live && shift) ? shift[tid] : 0.0f;
  const bool affine = scale != nullptr;
  float u = live ? u0[b * C + tid] : 0.f, acc = 0.f;
  for (int t0 = 0; t0 < T; t0 += TCH) {
    const int nt = min(TCH, T - t0), n = nt * C;
    const int64_t base = (b * (int64_t)T + t0) * C;
    block_copy(su, Z + base, n);
    __syncthreads();
    if (live)
      for (int t = 0; t < nt; ++t) {
        const float z = su[t * C + tid];
        const float cur = affine ? __fmaf_rn(z, sc, sf) : z;
        u = __fadd_rn(__fmul_rn(al, u), __fmul_rn(oma, cur));  // snns.py:822
        su[t * C + tid] = u;
      }
    __syncthreads();
    block_copy(U + base, su, n);
    for (int t = warp; t < nt; t += nw) {
      float* row = su + t * C;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
      m = warp_max(m);
      float den = 0.f;
      for (int c = lane; c < C; c += 32) {      // each lane revisits only its own entries: exp once, kept in place
        const float e = expf(row[c] - m);
        row[c] = e;
        den += e;
      }
      den = warp_sum(den);
      for (int c = lane; c < C; c += 32) row[c] = row[c] / den;
    }
    __syncthreads();
    if (live)
      for (int t = 0; t < nt; ++t) acc += su[t * C + tid];  // snns.py:823
    __syncthreads();
  }
  if (live) out[b * C + tid] = acc;
}

// Reverse pass, same chunking (chunks walked backwards): the softmax Jacobian term x_t = p_t (g - <p_t, g>)
// depends only on the taped u_t, so one warp per timestep computes it for the whole chunk; the dependent
// chain du_t = x_t + alpha du_{t+1} then runs per class thread without reductions.
__global__ void __launch_bounds__(1024) readout_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ U,
                                   const float* __restrict__ alpha, const float* __restrict__ u0,
                                   float* __restrict__ dI, float* __restrict__ p_alpha, int T, int C, int TCH) {
  extern __shared__ float sm[];
  const int pad = (4 - (C & 3)) & 3;                            // so that su + C (the chunk itself) is 16-byte aligned
  float* su = sm + pad;                                          // [TCH + 1][C]: row 0 = u of the step before the chunk
  float* sx = sm + (pad + (TCH + 1) * C + 3) / 4 * 4;            // [TCH][C], 16-byte aligned
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const bool live = tid < C;
  const int64_t b = blockIdx.x;
  const float al = live ? alpha[tid] : 0.f, oma = __fsub_rn(1.0f, al), inv_oma = 1.0f / oma;
  const float* go = gout + b * C;
  float du = 0.f, pa = 0.f;
  const int nchunks = (T + TCH - 1) / TCH;
  for (int ch = nchunks - 1; ch >= 0; --ch) {
    const int t0 = ch * TCH, nt = min(TCH, T - t0), n = nt * C;
    const int64_t base = (b * (int64_t)T + t0) * C;
    block_copy(su + C, U + base, n);
    for (int i = tid; i < C; i += blockDim.x) su[i] = t0 > 0 ? U[base - C + i] : u0[b * C + i];
    __syncthreads();
    for (int t = warp; t < nt; t += nw) {
      const float* row = su + (t + 1) * C;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
      m = warp_max(m);
      float den = 0.f;
      for (int c = lane; c < C; c += 32) {      // exp once per entry, parked in the output row (own entries only)
        const float e = expf(row[c] - m);
        sx[t * C + c] = e;
        den += e;
      }
      den = warp_sum(den);
      float dot = 0.f;
      for (int c = lane; c < C; c += 32) {
        const float pr = sx[t * C + c] / den;
        sx[t * C + c] = pr;
        dot += pr * go[c];
      }
      dot = warp_sum(dot);
      for (int c = lane; c < C; c += 32) sx[t * C + c] = sx[t * C + c] * (go[c] - dot);
    }
    __syncthreads();
    if (live)
      for (int t = nt - 1; t >= 0; --t) {
        const float du_t = sx[t * C + tid] + al * du;
        sx[t * C + tid] = oma * du_t;
        pa += du_t * ((su[t * C + tid] - su[(t + 1) * C + tid]) * inv_oma);  // d u_t / d alpha = u_{t-1} - I_t
        du = du_t;
      }
    __syncthreads();
    block_copy(dI + base, sx, n);
    __syncthreads();
  }
  if (live) p_alpha[b * C + tid] = pa;
}

static int readout_chunk(int T, int C, int arrays) {
  int tch = (40 * 1024) / (arrays * C * (int)sizeof(float)) - 1;
  if (tch > T) tch = T;
  return tch < 1 ? 1 : tch;
}

// ---- cross-entropy of the readout's output (exp.py:83 nn.CrossEntropyLoss(), exp.py:362): mean over the batch of
// logsumexp(x_b) - x_b[y_b], and its gradient (softmax(x_b) - onehot(y_b)) * gloss / B.  One block; a warp per row
// (lane-strided classes, shuffle reductions), the B row losses are summed by one warp in a fixed order: deterministic.
__global__ void ce_fwd_kernel(const float* __restrict__ X, const long long* __restrict__ y, int B, int C,
                              float* __restrict__ loss, float* __restrict__ lse) {
  extern __shared__ float row_loss[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int b = warp; b < B; b += nw) {
    const float* x = X + (int64_t)b * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int c = lane; c < C; c += 32) z += expf(x[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (lane == 0) {
      const float l = m + logf(z);
      const long long t = y[b];
      lse[b] = l;
      row_loss[b] = (t >= 0 && t < C) ? l - x[t] : 0.f;
    }
  }
  __syncthreads();
  if (warp == 0) {
    // lane l sums rows l, l + 32, ... in row order, then a fixed shuffle tree: deterministic, without the 256-step serial
    // chain of a single thread
    float sum = 0.f;
    for (int b = lane; b < B; b += 32) sum += row_loss[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) *loss = sum / (float)B;
  }
}

__global__ void ce_bwd_kernel(const float* __restrict__ X, const long long* __restrict__ y, const float* __restrict__ lse,
                              const float* __restrict__ gloss, int B, int C, float* __restrict__ dX) {
  const int64_t n = (int64_t)B * C;
  const float g = *gloss / (float)B;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / C;
    const int c = (int)(i - b * C);
    const long long t = y[b];
    const bool valid = t >= 0 && t < C;
    dX[i] = valid ? (expf(X[i] - lse[b]) - (c == t ? 1.f : 0.f)) * g : 0.f;
  }
}

}  // namespace sparch

using namespace sparch;

// The chain phase uses one thread per class, the softmax phase one WARP per timestep of the chunk: with few classes the
// block is sized for the softmax phase (16 warps share a chunk's ~100 timesteps instead of 4).
static int readout_min_threads(int B) {
  static const char* e = getenv("SPARCH_B200_READOUT_THREADS");
  if (e) return atoi(e);
  return B <= sm_count() ? 1024 : 512;      // one block per batch row: few rows = few blocks, make each one wide
}

extern "C" {

int sparch_readout_fwd(const float* Z, const float* scale, const float* shift, const float* alpha,
                       const float* u0, float* out, float* U, int B, int T, int C, sparch_stream_t st) {
  SPARCH_REQUIRE(B >= 0 && T >= 0 && C > 0 && C <= 1024, "bad shape (classes must be 1..1024)");
  SPARCH_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  if (B == 0) return SPARCH_OK;
  SPARCH_REQUIRE(alpha && u0 && out && (T == 0 || (Z && U)), "null pointer");
  int threads = ((C + 31) / 32) * 32;
  if (threads < readout_min_threads(B)) threads = readout_min_threads(B);
  const int tch = readout_chunk(T, C, 1);
  readout_fwd_kernel<<<B, threads, (size_t)tch * C * sizeof(float), as_stream(st)>>>(Z, scale, shift, alpha, u0, out, U,
                                                                                      T, C, tch);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_readout_bwd(const float* gout, const float* U, const float* alpha, const float* u0,
                       float* dI, float* p_alpha, int B, int T, int C, sparch_stream_t st) {
  SPARCH_REQUIRE(B >= 0 && T >= 0 && C > 0 && C <= 1024, "bad shape (classes must be 1..1024)");
  if (B == 0) return SPARCH_OK;
  SPARCH_REQUIRE(gout && alpha && u0 && p_alpha && (T == 0 || (U && dI)), "null pointer");
  int threads = ((C + 31) / 32) * 32;
  if (threads < readout_min_threads(B)) threads = readout_min_threads(B);
  const int tch = readout_chunk(T, C, 2);
  readout_bwd_kernel<<<B, threads, (size_t)(2 * tch + 1) * C * sizeof(float) + 32, as_stream(st)>>>(gout, U, alpha, u0, dI,
                                                                                                p_alpha, T, C, tch);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_ce_fwd(const float* logits, const int64_t* target, int B, int C, float* loss, float* lse, sparch_stream_t st) {
  SPARCH_REQUIRE(logits && target && loss && lse && B > 0 && C > 0, "bad argument");
  SPARCH_REQUIRE((size_t)B * sizeof(float) <= 200 * 1024, "batch too large for the one-block loss kernel");
  static PerDeviceOnce attr_once;
  if (attr_once.first()) SPARCH_CUDA(cudaFuncSetAttribute(ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ce_fwd_kernel<<<1, 1024, (size_t)B * sizeof(float), as_stream(st)>>>(logits, reinterpret_cast<const long long*>(target), B, C,
                                                                      loss, lse);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_ce_bwd(const float* logits, const int64_t* target, const float* lse, const float* gloss, int B, int C,
                  float* dlogits, sparch_stream_t st) {
  SPARCH_REQUIRE(logits && target && lse && gloss && dlogits && B > 0 && C > 0, "bad argument");
  const int64_t n = (int64_t)B * C;
  int nb = (int)((n + 255) / 256);
  if (nb > sm_count() * 8) nb = sm_count() * 8;
  ce_bwd_kernel<<<nb, 256, 0, as_stream(st)>>>(logits, reinterpret_cast<const long long*>(target), lse, gloss, B, C, dlogits);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

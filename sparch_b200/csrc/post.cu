// What surrounds the membrane recurrence of a spiking layer, fused into single passes:
//   * spike_post_fwd: dropout of the spike tensor (snns.py:692) + the per-neuron spike counts behind
//     SNN.forward's firing rates (snns.py:174) + the 16-bit {0,1} operand terms the next layer's projection
//     GEMM and this layer's dV GEMM read -- one read of S instead of four passes (dropout, mean, two splits);
//   * spike_post_bwd: dropout backward with the mask REGENERATED from the counter-based generator (no mask
//     tensor is stored) + the row maxima of the gradient that the tcgen05 reverse recurrence scales with;
//   * neuron_params / param_grads: the clamp of alpha, beta, a, b (snns.py:706-709) and its backward (batch
//     reduction of the per-(b,h) partial gradients + clamp mask) as one launch each instead of ~25 ATen ops.
#include <stdlib.h>

#include "common.cuh"

namespace sparch {

// Philox4x32-10 (Salmon et al. 2011): counter-based, so forward and backward draw the same mask from
// (seed, item) without storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// keep[j] for the 8 elements of work item `item` (8 consecutive columns of one row): element kept iff its
// 32-bit draw >= thresh = p * 2^32.
__device__ __forceinline__ void keep8(const unsigned long long* __restrict__ seed, long long item, uint32_t thresh,
                                      bool (&keep)[8]) {
  const unsigned long long sd = *seed;
  const uint2 key = make_uint2((uint32_t)sd, (uint32_t)(sd >> 32));
  const uint4 a = philox4x32_10(make_uint4((uint32_t)item, (uint32_t)((unsigned long long)item >> 32), 0u, 0u), key);
  const uint4 b = philox4x32_10(make_uint4((uint32_t)item, (uint32_t)((unsigned long long)item >> 32), 1u, 0u), key);
  keep[0] = a.x >= thresh; keep[1] = a.y >= thresh; keep[2] = a.z >= thresh; keep[3] = a.w >= thresh;
  keep[4] = b.x >= thresh; keep[5] = b.y >= thresh; keep[6] = b.z >= thresh; keep[7] = b.w >= thresh;
}

// Where the spikes come from when the forward recurrence did not write them as fp32: the words it published
// (csrc/recur_fwd_tc.cu: [T][group][slice][128 rows], 16 spikes per word as 2-bit fields, spike i at bit 2 i).
struct SpikeBits {
  const uint32_t* words;   // NULL: read S
  int T, groups, nsl;
  float* s_last;           // optional (Be, H): the spikes of the last step as fp32 (the dV boundary operand needs them)
  int rev_from;            // bidirectional merge (snns.py:686-689) fused: rows b >= rev_from are the time-reversed pass of
                           // row b - rev_from; out / term / counts / the dropout position then live in the merged
                           // (rev_from, T, 2 H) layout, out[b - rev_from][T - 1 - t][H + h].  0 = off.  (H % 8 == 0.)
};

__global__ void __launch_bounds__(256)
spike_post_fwd_kernel(const float* __restrict__ S, const SpikeBits sb_, long long M, int H, long long ld, float scale, uint32_t thresh,
                      const unsigned long long* __restrict__ seed, float* __restrict__ out,
                      uint16_t* __restrict__ term, uint16_t* __restrict__ sterm, uint16_t one_bits,
                      int* __restrict__ counts, int use_hist) {
  extern __shared__ int hist[];
  const int Hc = (sb_.words && sb_.rev_from) ? 2 * H : H;   // columns of the counts
  if (counts && use_hist) {
    for (int h = threadIdx.x; h < Hc; h += blockDim.x) hist[h] = 0;
    __syncthreads();
  }
  const long long segs = ld / 8, n = M * segs;
  const bool vec = ((H & 3) == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0) &&
                   (!out || (reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / segs;
    const int c = (int)(i - r * segs) * 8;
    const float* src = S + r * H + c;
    float x[8];
    if (sb_.words) {
      // row r = (b, t); the 8 columns c .. c + 7 are one byte of the word of slice c / 16
      const long long b = r / sb_.T;
      const int t = (int)(r - b * sb_.T);
      const uint32_t w = c < H ? sb_.words[(((size_t)t * sb_.groups + (size_t)(b >> 7)) * sb_.nsl + (c >> 4)) * 128 + (b & 127)] : 0u;
      const uint32_t f = w >> (2 * (c & 15));
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c + j < H && ((f >> (2 * j)) & 1u)) ? 1.f : 0.f;
      if (sb_.s_last && t == sb_.T - 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < H) sb_.s_last[b * H + c + j] = x[j];
      }
    } else if (vec && c + 8 <= H) {
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = c + j < H ? src[j] : 0.f;
    }
    // where the item goes: its own place, or its place in the merged bidirectional tensor
    long long ro = r, io = i, Ho = H, ldo = ld;
    int co = c;
    if (sb_.words && sb_.rev_from) {
      const long long b = r / sb_.T;
      const int t = (int)(r - b * sb_.T);
      Ho = 2 * (long long)H;
      ldo = 2 * ld;
      if (b >= sb_.rev_from) {
        ro = (b - sb_.rev_from) * sb_.T + (sb_.T - 1 - t);
        co = c + H;
      }
      io = ro * (ldo / 8) + co / 8;
    }
    bool keep[8];
    if (thresh) {
      keep8(seed, io, thresh, keep);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) keep[j] = true;
    }
    float y[8];
    __align__(16) uint16_t tb[8], sb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      y[j] = keep[j] ? x[j] * scale : 0.f;
      tb[j] = y[j] != 0.f ? one_bits : (uint16_t)0;
      sb[j] = x[j] != 0.f ? one_bits : (uint16_t)0;
    }
    if (out) {
      float* dst = out + ro * Ho + co;
      if (vec && c + 8 <= H) {
        *reinterpret_cast<float4*>(dst) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(y[4], y[5], y[6], y[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < H) dst[j] = y[j];
      }
    }
    if (term) *reinterpret_cast<uint4*>(term + ro * ldo + co) = *reinterpret_cast<const uint4*>(tb);
    if (sterm) *reinterpret_cast<uint4*>(sterm + r * ld + c) = *reinterpret_cast<const uint4*>(sb);
    if (counts) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (y[j] != 0.f) atomicAdd(use_hist ? &hist[co + j] : &counts[co + j], 1);
    }
  }
  if (counts && use_hist) {
    __syncthreads();
    for (int h = threadIdx.x; h < Hc; h += blockDim.x)
      if (hist[h]) atomicAdd(&counts[h], hist[h]);
  }
}

// The same pass for the packed-plane input as MASK arithmetic.  The generic kernel above spends ~930 warp instructions
// per 8-element item (ncu: issue slots 67 % busy, DRAM 29 %), most of them per-element selects, compares and
// reconverging branches around the count atomics; here an item is an 8-bit spike mask AND an 8-bit keep mask, the
// outputs are selected from the mask bits, and the per-neuron counts live in registers: every launch makes the grid
// stride a multiple of the segments per row, so a thread keeps its 8 columns for all its rows and adds them to the
// block's histogram once, at the end.  Same item numbering, mask stream and results as the generic kernel.
__global__ void __launch_bounds__(256)
spike_post_bits_kernel(const SpikeBits sb_, long long M, int H, long long ld, float scale, uint32_t thresh,
                       const unsigned long long* __restrict__ seed, float* __restrict__ out, uint16_t* __restrict__ term,
                       uint16_t* __restrict__ sterm, uint16_t one_bits, int* __restrict__ counts) {
  extern __shared__ int hist[];
  const bool bidir = sb_.rev_from > 0;
  const int Hc = bidir ? 2 * H : H;
  if (counts) {
    for (int h = threadIdx.x; h < Hc; h += blockDim.x) hist[h] = 0;
    __syncthreads();
  }
  const long long segs = ld / 8, n = M * segs;
  const long long stride = (long long)gridDim.x * blockDim.x;          // a multiple of segs (host)
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)(i0 % segs) * 8;                                   // this thread's columns, for all its items
  const int nvalid = max(0, min(8, H - c));
  const uint32_t vmask = (1u << nvalid) - 1u;
  const bool vec = ((H & 3) == 0) && (!out || (reinterpret_cast<uintptr_t>(out) & 15) == 0) && nvalid == 8;
  const uint32_t one2 = (uint32_t)one_bits * 0x00010001u;
  const unsigned long long sd = thresh ? *seed : 0ull;
  const uint2 key = make_uint2((uint32_t)sd, (uint32_t)(sd >> 32));
  const long long ldo = bidir ? 2 * ld : ld, Ho = bidir ? 2 * (long long)H : H;
  int cnt0[8], cnt1[8];                                                 // counts of the two directions' columns
#pragma unroll
  for (int j = 0; j < 8; ++j) cnt0[j] = cnt1[j] = 0;
  // row, batch row and timestep of an item advance by constants (the stride is a whole number of rows): no division in
  // the loop
  const long long rstep = stride / segs;
  const long long bstep = rstep / sb_.T;
  const int tstep = (int)(rstep - bstep * sb_.T);
  long long r = i0 / segs, b = r / sb_.T;
  int t = (int)(r - b * sb_.T);
  for (long long i = i0; i < n; i += stride, r += rstep, b += bstep, t += tstep) {
    if (t >= sb_.T) {
      t -= sb_.T;
      ++b;
    }
    const uint32_t w = c < H ? sb_.words[(((size_t)t * sb_.groups + (size_t)(b >> 7)) * sb_.nsl + (c >> 4)) * 128 + (b & 127)] : 0u;
    uint32_t m = (w >> (2 * (c & 15))) & 0x5555u;                       // 8 spikes at the even bits -> 8 adjacent bits
    m = (m | (m >> 1)) & 0x3333u;
    m = (m | (m >> 2)) & 0x0f0fu;
    m = (m | (m >> 4)) & vmask;
    if (sb_.s_last && t == sb_.T - 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) sb_.s_last[b * H + c + j] = (float)((m >> j) & 1u);
    }
    long long ro = r, io = i;
    int co = c;
    const bool second = bidir && b >= sb_.rev_from;
    if (bidir) {
      if (second) {
        ro = (b - sb_.rev_from) * sb_.T + (sb_.T - 1 - t);
        co = c + H;
      }
      io = ro * (ldo / 8) + co / 8;
    }
    uint32_t y = m;
    if (thresh) {
      const uint4 a = philox4x32_10(make_uint4((uint32_t)io, (uint32_t)((unsigned long long)io >> 32), 0u, 0u), key);
      const uint4 q = philox4x32_10(make_uint4((uint32_t)io, (uint32_t)((unsigned long long)io >> 32), 1u, 0u), key);
      const uint32_t k = (a.x >= thresh ? 1u : 0u) | (a.y >= thresh ? 2u : 0u) | (a.z >= thresh ? 4u : 0u) |
                         (a.w >= thresh ? 8u : 0u) | (q.x >= thresh ? 16u : 0u) | (q.y >= thresh ? 32u : 0u) |
                         (q.z >= thresh ? 64u : 0u) | (q.w >= thresh ? 128u : 0u);
      y &= k;
    }
    if (out) {
      float* dst = out + ro * Ho + co;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((y >> j) & 1u) ? scale : 0.f;
      if (vec) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nvalid) dst[j] = v[j];
      }
    }
    // bit pair (2 p, 2 p + 1) -> two 16-bit patterns: (bit ? one : 0) | (bit' ? one : 0) << 16
    auto halves = [&](uint32_t bits) {
      uint4 h;
      uint32_t e[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const uint32_t lo = 0u - ((bits >> (2 * p)) & 1u), hi = 0u - ((bits >> (2 * p + 1)) & 1u);
        e[p] = one2 & ((lo & 0x0000ffffu) | (hi & 0xffff0000u));
      }
      h.x = e[0]; h.y = e[1]; h.z = e[2]; h.w = e[3];
      return h;
    };
    if (term) *reinterpret_cast<uint4*>(term + ro * ldo + co) = halves(y);
    if (sterm) *reinterpret_cast<uint4*>(sterm + r * ld + c) = halves(m);
    if (second) {
#pragma unroll
      for (int j = 0; j < 8; ++j) cnt1[j] += (int)((y >> j) & 1u);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) cnt0[j] += (int)((y >> j) & 1u);
    }
  }
  if (counts) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < nvalid && cnt0[j]) atomicAdd(&hist[c + j], cnt0[j]);
      if (bidir && j < nvalid && cnt1[j]) atomicAdd(&hist[H + c + j], cnt1[j]);
    }
    __syncthreads();
    for (int h = threadIdx.x; h < Hc; h += blockDim.x)
      if (hist[h]) atomicAdd(&counts[h], hist[h]);
  }
}

__global__ void __launch_bounds__(256)
spike_post_bwd_kernel(const float* __restrict__ G, long long M, int H, float scale, uint32_t thresh,
                      const unsigned long long* __restrict__ seed, float* __restrict__ GS, uint32_t* __restrict__ gmax,
                      const int rev_from, const int T) {
  // rev_from > 0: G is the gradient of the MERGED bidirectional output (rev_from, T, H = 2 Hc); GS / gmax are written
  // in the recurrence's own order (2 rev_from, T, Hc): columns >= Hc go to row b + rev_from at time T - 1 - t
  const int Hc = rev_from ? H / 2 : H;
  const long long segs = (H + 7) / 8, n = M * segs;
  const bool vec = ((H & 3) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(GS) & 15) == 0);
  const long long total = (n + 31) / 32 * 32;  // whole warps walk the loop (shuffles below)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const bool on = i < n;
    long long r = on ? i / segs : -1;
    const int c = on ? (int)(i - r * segs) * 8 : 0;
    float m = 0.f;
    if (on) {
      const float* src = G + r * H + c;
      float x[8];
      if (vec && c + 8 <= H) {
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = c + j < H ? src[j] : 0.f;
      }
      bool keep[8];
      if (thresh) {
        keep8(seed, i, thresh, keep);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) keep[j] = true;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = keep[j] ? x[j] * scale : 0.f;
        m = fmaxf(m, fabsf(x[j]));
      }
      float* dst = GS + r * H + c;
      if (rev_from) {
        const long long b = r / T;
        const int t = (int)(r - b * T);
        r = c >= Hc ? (b + rev_from) * T + (T - 1 - t) : r;    // from here on: the row of GS / gmax
        dst = GS + r * Hc + (c >= Hc ? c - Hc : c);
      }
      if (vec && c + 8 <= H) {
        *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < H) dst[j] = x[j];
      }
    }
    if (gmax) {
      // one atomic per warp when all its items sit in the same row (H a multiple of 256), else per thread
      const long long r0 = __shfl_sync(0xffffffffu, r, 0);
      if (__all_sync(0xffffffffu, r == r0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0 && on && m > 0.f) atomicMax(&gmax[r], __float_as_uint(m));
      } else if (on && m > 0.f) {
        atomicMax(&gmax[r], __float_as_uint(m));
      }
    }
  }
}

// out[k][h] = clamp(p_k[h], lo_k, hi_k) for the given parameter vectors (NULL = skipped), torch.clamp semantics.
struct ParamSet {
  const float* p[4];
  float lo[4], hi[4];
};

__global__ void neuron_params_kernel(const ParamSet ps, int nk, int H, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nk * H) return;
  const int k = i / H, h = i - k * H;
  const float x = ps.p[k][h];
  out[i] = x != x ? x : fminf(fmaxf(x, ps.lo[k]), ps.hi[k]);
}

// grads[k][h] = (sum_b part[k][b][h]) * [lo_k <= p_k[h] <= hi_k]   (clamp backward passes on the closed interval)
__global__ void __launch_bounds__(256) param_grads_kernel(const float* __restrict__ part, const ParamSet ps, int nk, int Be,
                                                          int H, float* __restrict__ grads) {
  __shared__ float red[8][33];
  const int k = blockIdx.y, h = blockIdx.x * 32 + threadIdx.x, ry = threadIdx.y;
  float s = 0.f;
  if (h < H)
    for (int b = ry; b < Be; b += 8) s += part[((size_t)k * Be + b) * H + h];
  red[ry][threadIdx.x] = s;
  __syncthreads();
  if (ry == 0 && h < H) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    const float x = ps.p[k][h];
    grads[(size_t)k * H + h] = (x >= ps.lo[k] && x <= ps.hi[k]) ? t : 0.f;
  }
}


// ---- small helpers of the recurrent layers (snns.py:712, 702): no library kernel on the product path
__global__ void v0_copy_kernel(const float* __restrict__ V, int H, float* __restrict__ V0) {
  const int64_t n = (int64_t)H * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    V0[i] = (i / H == i % H) ? 0.f : V[i];
}

__global__ void zero_diag_kernel(float* __restrict__ A, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H) A[(int64_t)i * H + i] = 0.f;
}

// first[b] = s0[b] - S[b-1][T-1] (b > 0): what frame (b, 0) of dI must be paired with INSTEAD of the frame before it
// in the (b, t)-flattened spike matrix (the dV GEMM pairs frame m of dI with frame m - 1 of S)
__global__ void dv_boundary_kernel(const float* __restrict__ s0, const float* __restrict__ S, int Be, int T, int H,
                                   float* __restrict__ first) {
  const int64_t n = (int64_t)Be * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / H, h = i - b * H;
    float x = s0[i];
    if (b > 0) x -= S[((b - 1) * T + (T - 1)) * H + h];
    first[i] = x;
  }
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_spike_post_fwd(const float* S, int64_t M, int H, float p_drop, const void* seed, float* out, void* term,
                          void* sterm, int fp16_terms, int* counts, sparch_stream_t st_) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && p_drop >= 0.f && p_drop < 1.f, "bad argument");
  SPARCH_REQUIRE(p_drop == 0.f || (seed && out), "dropout needs the seed word and an output tensor");
  cudaStream_t st = as_stream(st_);
  if (counts) SPARCH_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * H, st));
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(S, "null pointer");
  const int64_t ld = ((int64_t)H + 7) / 8 * 8;
  const int64_t n = M * (ld / 8);
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  const int use_hist = counts && H <= 8192;
  const uint32_t thresh = p_drop > 0.f ? (uint32_t)fmin((double)p_drop * 4294967296.0, 4294967295.0) : 0u;
  spike_post_fwd_kernel<<<(unsigned)(g < cap ? g : cap), 256, use_hist ? sizeof(int) * H : 0, st>>>(
      S, SpikeBits{nullptr, 0, 0, 0, nullptr, 0}, M, H, ld, 1.0f / (1.0f - p_drop), thresh,
      reinterpret_cast<const unsigned long long*>(seed), out,
      reinterpret_cast<uint16_t*>(term), reinterpret_cast<uint16_t*>(sterm), fp16_terms ? (uint16_t)0x3C00 : (uint16_t)0x3F80,
      counts, use_hist);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_spike_post_fwd_bits(const uint32_t* bits, int Be, int T, int H, float p_drop, const void* seed, float* out,
                               void* term, void* sterm, int fp16_terms, int* counts, float* s_last, sparch_stream_t st_) {
  return sparch_spike_post_fwd_bits_bidir(bits, Be, T, H, p_drop, seed, out, term, sterm, fp16_terms, counts, s_last, 0, st_);
}

int sparch_spike_post_fwd_bits_bidir(const uint32_t* bits, int Be, int T, int H, float p_drop, const void* seed,
                                     float* out, void* term, void* sterm, int fp16_terms, int* counts, float* s_last,
                                     int rev_from, sparch_stream_t st_) {
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0 && p_drop >= 0.f && p_drop < 1.f, "bad argument");
  SPARCH_REQUIRE(p_drop == 0.f || seed, "dropout needs the seed word");
  SPARCH_REQUIRE(rev_from == 0 || (rev_from > 0 && Be == 2 * rev_from && H % 8 == 0),
                 "merged bidirectional output: Be = 2 * rev_from and H a multiple of 8");
  cudaStream_t st = as_stream(st_);
  if (counts) SPARCH_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (rev_from ? 2 * H : H), st));
  const int64_t M = (int64_t)Be * T;
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(bits && out, "null pointer");
  const int64_t ld = ((int64_t)H + 7) / 8 * 8;
  const int64_t n = M * (ld / 8);
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  const int Hc = rev_from ? 2 * H : H;
  const int use_hist = counts && Hc <= 8192;
  const uint32_t thresh = p_drop > 0.f ? (uint32_t)fmin((double)p_drop * 4294967296.0, 4294967295.0) : 0u;
  const SpikeBits sb{bits, T, (Be + 127) / 128, (H + 15) / 16, s_last, rev_from};
  static const bool generic = getenv("SPARCH_B200_POST_GENERIC") != nullptr;
  if ((use_hist || !counts) && !generic) {
    // mask-arithmetic kernel: the grid stride must be a multiple of the segments per row (a thread keeps its columns)
    const int64_t segs = ld / 8;
    int64_t unit = segs;                           // smallest block count whose 256 * blocks is a multiple of segs
    for (int64_t x = 256, y = segs; y;) { const int64_t tq = x % y; x = y; y = tq; unit = segs / x; }
    int64_t nb = g < cap ? g : cap;
    nb = (nb / unit) * unit;
    if (nb >= unit && nb >= 1) {
      spike_post_bits_kernel<<<(unsigned)nb, 256, counts ? sizeof(int) * Hc : 0, st>>>(
          sb, M, H, ld, 1.0f / (1.0f - p_drop), thresh, reinterpret_cast<const unsigned long long*>(seed), out,
          reinterpret_cast<uint16_t*>(term), reinterpret_cast<uint16_t*>(sterm), fp16_terms ? (uint16_t)0x3C00 : (uint16_t)0x3F80,
          counts);
      SPARCH_LAUNCH_OK();
      return SPARCH_OK;
    }
  }
  spike_post_fwd_kernel<<<(unsigned)(g < cap ? g : cap), 256, use_hist ? sizeof(int) * Hc : 0, st>>>(
      nullptr, sb, M, H, ld, 1.0f / (1.0f - p_drop), thresh, reinterpret_cast<const unsigned long long*>(seed), out,
      reinterpret_cast<uint16_t*>(term), reinterpret_cast<uint16_t*>(sterm), fp16_terms ? (uint16_t)0x3C00 : (uint16_t)0x3F80,
      counts, use_hist);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_spike_post_bwd(const float* G, int64_t M, int H, float p_drop, const void* seed, float* GS, float* gmax,
                          sparch_stream_t st_) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && p_drop > 0.f && p_drop < 1.f && seed, "bad argument");
  cudaStream_t st = as_stream(st_);
  if (gmax && M > 0) SPARCH_CUDA(cudaMemsetAsync(gmax, 0, sizeof(float) * M, st));
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && GS, "null pointer");
  const int64_t n = M * (((int64_t)H + 7) / 8);
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  const uint32_t thresh = (uint32_t)fmin((double)p_drop * 4294967296.0, 4294967295.0);
  spike_post_bwd_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(G, M, H, 1.0f / (1.0f - p_drop), thresh,
                                                                        reinterpret_cast<const unsigned long long*>(seed),
                                                                        GS, reinterpret_cast<uint32_t*>(gmax), 0, 1);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_spike_post_bwd_bidir(const float* G, int B, int T, int H, float p_drop, const void* seed, float* GS,
                                float* gmax, sparch_stream_t st_) {
  SPARCH_REQUIRE(B >= 0 && T >= 0 && H > 0 && H % 8 == 0 && p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || seed),
                 "bad argument");
  cudaStream_t st = as_stream(st_);
  const int64_t M = (int64_t)B * T;
  if (gmax && M > 0) SPARCH_CUDA(cudaMemsetAsync(gmax, 0, sizeof(float) * 2 * M, st));
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && GS, "null pointer");
  const int64_t n = M * (2 * (int64_t)H / 8);
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  const uint32_t thresh = p_drop > 0.f ? (uint32_t)fmin((double)p_drop * 4294967296.0, 4294967295.0) : 0u;
  spike_post_bwd_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(G, M, 2 * H, 1.0f / (1.0f - p_drop), thresh,
                                                                        reinterpret_cast<const unsigned long long*>(seed),
                                                                        GS, reinterpret_cast<uint32_t*>(gmax), B, T);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

static ParamSet make_param_set(const float* alpha, const float* beta, const float* a, const float* b, const float* lims,
                               int nk) {
  ParamSet ps;
  const float* p[4] = {alpha, beta, a, b};
  for (int k = 0; k < 4; ++k) {
    ps.p[k] = k < nk ? p[k] : nullptr;
    ps.lo[k] = lims[2 * k];
    ps.hi[k] = lims[2 * k + 1];
  }
  return ps;
}

int sparch_neuron_params(const float* alpha, const float* beta, const float* a, const float* b, const float* lims,
                         int nk, int H, float* out, sparch_stream_t st) {
  SPARCH_REQUIRE((nk == 1 || nk == 4) && H > 0 && alpha && lims && out && (nk == 1 || (beta && a && b)), "bad argument");
  const ParamSet ps = make_param_set(alpha, beta, a, b, lims, nk);
  neuron_params_kernel<<<(nk * H + 255) / 256, 256, 0, as_stream(st)>>>(ps, nk, H, out);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_param_grads(const float* part, const float* alpha, const float* beta, const float* a, const float* b,
                       const float* lims, int nk, int Be, int H, float* grads, sparch_stream_t st) {
  SPARCH_REQUIRE((nk == 1 || nk == 4) && H > 0 && Be >= 0 && alpha && lims && grads && (Be == 0 || part) &&
                     (nk == 1 || (beta && a && b)),
                 "bad argument");
  const ParamSet ps = make_param_set(alpha, beta, a, b, lims, nk);
  param_grads_kernel<<<dim3((H + 31) / 32, nk), dim3(32, 8), 0, as_stream(st)>>>(part, ps, nk, Be, H, grads);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_recur_v0(const float* V, int H, float* V0, sparch_stream_t st) {
  SPARCH_REQUIRE(V && V0 && H > 0, "bad argument");
  const int64_t n = (int64_t)H * H;
  int nb = (int)((n + 255) / 256);
  if (nb > sm_count() * 8) nb = sm_count() * 8;
  v0_copy_kernel<<<nb, 256, 0, as_stream(st)>>>(V, H, V0);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_zero_diag(float* A, int H, sparch_stream_t st) {
  SPARCH_REQUIRE(A && H > 0, "bad argument");
  zero_diag_kernel<<<(H + 255) / 256, 256, 0, as_stream(st)>>>(A, H);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_dv_boundary(const float* s0, const float* S, int Be, int T, int H, float* first, sparch_stream_t st) {
  SPARCH_REQUIRE(s0 && S && first && Be > 0 && T > 0 && H > 0, "bad argument");
  const int64_t n = (int64_t)Be * H;
  int nb = (int)((n + 255) / 256);
  if (nb > sm_count() * 8) nb = sm_count() * 8;
  dv_boundary_kernel<<<nb, 256, 0, as_stream(st)>>>(s0, S, Be, T, H, first);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

// Boxcar surrogate (elementwise) and the BatchNorm1d pieces around the projection:
// column statistics, train-mode fold to (scale, shift), backward reductions and apply.
// All are HBM-bound streaming kernels: 16-byte loads where the row length allows, grids
// sized from the SM count, fp64 accumulation for the column sums.
#include <stdarg.h>

#include <string>

#include <cuda_fp16.h>

#include "common.cuh"

namespace sparch {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

int cuda_fail(const char* what, cudaError_t e) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return SPARCH_ERR_CUDA;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return dev;
}

int sm_count() {
  static int n[SPARCH_MAX_DEVICES] = {};
  const int dev = current_device();
  const int slot = (dev >= 0 && dev < SPARCH_MAX_DEVICES) ? dev : 0;
  if (n[slot] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[slot] = v;
  }
  return n[slot];
}

// ------------------------------------------------------------------ boxcar
__global__ void boxcar_fwd_kernel(const float* __restrict__ x, float* __restrict__ s, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) s[i] = spike_of(x[i]);
}

__global__ void boxcar_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                  float* __restrict__ gx, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) gx[i] = window_of(x[i]) ? g[i] : 0.0f;
}

// ------------------------------------------------------------------ column sums
// Block = 32 columns x 8 row-lanes; each block walks a contiguous slab of rows, so a warp reads
// 128 contiguous bytes per row.  Partials are combined in fp64 with atomics (order effects are
// ~1e-16, far below the fp32 results derived from them).
constexpr int CS_ROWS = 8;

template <bool DOT>
__global__ void col_reduce_kernel(const float* __restrict__ A, const float* __restrict__ Zn,
                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                  int64_t M, int H, int64_t rows_per_block, double* __restrict__ o1,
                                  double* __restrict__ o2) {
  __shared__ double sh1[CS_ROWS][33];
  __shared__ double sh2[CS_ROWS][33];
  const int h = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  double a1 = 0.0, a2 = 0.0;
  if (h < H) {
    float mu = 0.f, rs = 1.f;
    if (DOT && mean) { mu = mean[h]; rs = rstd[h]; }
    for (int64_t r = r0 + threadIdx.y; r < r1; r += CS_ROWS) {
      float v = A[r * H + h];
      if (DOT) {
        float xh = (Zn[r * H + h] - mu) * rs;
        a1 += (double)v;
        a2 += (double)v * (double)xh;
      } else {
        a1 += (double)v;
        a2 += (double)v * (double)v;
      }
    }
  }
  sh1[threadIdx.y][threadIdx.x] = a1;
  sh2[threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && h < H) {
#pragma unroll
    for (int k = 1; k < CS_ROWS; ++k) {
      a1 += sh1[k][threadIdx.x];
      a2 += sh2[k][threadIdx.x];
    }
    atomicAdd(&o1[h], a1);
    atomicAdd(&o2[h], a2);
  }
}

// Same reductions with 16-byte loads: a thread owns 4 consecutive columns, a block 128 columns x 8 row lanes,
// four rows in flight per thread.  A row's partial products are summed in fp32 over 4 rows before they enter the
// fp64 accumulators (the fp64 pipe is narrow; 4-term fp32 sums of same-sign-agnostic values cost <1 ulp each).
// amax (may be NULL): bit pattern of max|A|, for the fp16 split of A that follows (dV's dI operand).
constexpr int CR_FLY = 8;   // rows in flight per thread (4 left the kernel latency-bound at 62 % of the DRAM throughput)

template <bool DOT>
__global__ void __launch_bounds__(256)
col_reduce4_kernel(const float* __restrict__ A, const float* __restrict__ Zn, const float* __restrict__ mean,
                   const float* __restrict__ rstd, int64_t M, int H, int64_t rows_per_block, double* __restrict__ o1,
                   double* __restrict__ o2, uint32_t* __restrict__ amax, const int64_t rev_from = 0, const int T = 1) {
  // rev_from > 0 (bidirectional layer without a flipped copy of Z): row (b, t) of A with b >= rev_from pairs with row
  // (b - rev_from, T - 1 - t) of Zn
  __shared__ double sh[2][CS_ROWS][32][4];
  const int h = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  double a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  float mx = 0.f;
  if (h < H) {
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = make_float4(1.f, 1.f, 1.f, 1.f);
    if (DOT && mean) {
      mu = *reinterpret_cast<const float4*>(mean + h);
      rs = *reinterpret_cast<const float4*>(rstd + h);
    }
    for (int64_t r = r0 + threadIdx.y; r < r1; r += CR_FLY * CS_ROWS) {
      float4 v[CR_FLY], z[CR_FLY];
#pragma unroll
      for (int k = 0; k < CR_FLY; ++k) {
        const int64_t rr = r + k * CS_ROWS;
        v[k] = rr < r1 ? *reinterpret_cast<const float4*>(A + rr * H + h) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (DOT) {
          int64_t zr = rr;
          if (rev_from && rr >= rev_from * T && rr < r1) {
            const int64_t b = rr / T;
            zr = (b - rev_from) * T + (T - 1 - (rr - b * T));
          }
          z[k] = rr < r1 ? *reinterpret_cast<const float4*>(Zn + zr * H + h) : mu;
        }
      }
      float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < CR_FLY; ++k) {
        const float vv[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        const float zz[4] = {DOT ? (z[k].x - mu.x) * rs.x : v[k].x, DOT ? (z[k].y - mu.y) * rs.y : v[k].y,
                             DOT ? (z[k].z - mu.z) * rs.z : v[k].z, DOT ? (z[k].w - mu.w) * rs.w : v[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s1[j] += vv[j];
          s2[j] += vv[j] * zz[j];
          mx = fmaxf(mx, fabsf(vv[j]));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a1[j] += (double)s1[j];
        a2[j] += (double)s2[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sh[0][threadIdx.y][threadIdx.x][j] = a1[j];
    sh[1][threadIdx.y][threadIdx.x][j] = a2[j];
  }
  __syncthreads();
  if (threadIdx.y < 2 && h < H) {      // row lane 0 finishes the sums, lane 1 the dot products
    double* o = threadIdx.y ? o2 : o1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < CS_ROWS; ++k) t += sh[threadIdx.y][k][threadIdx.x][j];
      atomicAdd(&o[h + j], t);
    }
  }
  if (amax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0 && mx > 0.f) atomicMax(amax, __float_as_uint(mx));
  }
}

__global__ void absmax_scalar_kernel(const float* __restrict__ A, int64_t n, uint32_t* __restrict__ amax) {
  float mx = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mx = fmaxf(mx, fabsf(A[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(amax, __float_as_uint(mx));
}

static int col_reduce_launch(bool dot, const float* A, const float* Zn, const float* mean,
                             const float* rstd, int64_t M, int H, double* o1, double* o2,
                             cudaStream_t st, uint32_t* amax = nullptr, int64_t rev_from = 0, int T = 1) {
  SPARCH_CUDA(cudaMemsetAsync(o1, 0, sizeof(double) * H, st));
  SPARCH_CUDA(cudaMemsetAsync(o2, 0, sizeof(double) * H, st));
  if (amax) SPARCH_CUDA(cudaMemsetAsync(amax, 0, sizeof(uint32_t), st));
  if (M == 0) return SPARCH_OK;
  const bool al16 = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Zn) | reinterpret_cast<uintptr_t>(mean) |
                      reinterpret_cast<uintptr_t>(rstd)) & 15) == 0;
  if ((H & 3) == 0 && al16) {
    const int cb4 = (H / 4 + 31) / 32;
    int64_t want4 = (int64_t)sm_count() * 8 / cb4;
    if (want4 < 1) want4 = 1;
    int64_t rpb4 = (M + want4 - 1) / want4;
    if (rpb4 < 64) rpb4 = 64;
    dim3 grid4(cb4, (unsigned)((M + rpb4 - 1) / rpb4)), block4(32, CS_ROWS);
    if (dot)
      col_reduce4_kernel<true><<<grid4, block4, 0, st>>>(A, Zn, mean, rstd, M, H, rpb4, o1, o2, amax, rev_from, T);
    else
      col_reduce4_kernel<false><<<grid4, block4, 0, st>>>(A, nullptr, nullptr, nullptr, M, H, rpb4, o1, o2, amax);
    SPARCH_LAUNCH_OK();
    return SPARCH_OK;
  }
  SPARCH_REQUIRE(rev_from == 0, "the paired-row reduction needs H % 4 == 0 and 16-byte aligned tensors");
  if (amax) {
    absmax_scalar_kernel<<<sm_count() * 4, 256, 0, st>>>(A, M * (int64_t)H, amax);
    SPARCH_LAUNCH_OK();
  }
  int cb = (H + 31) / 32;
  // enough row-slabs to fill the machine a few times over, at least 64 rows each
  int64_t want = (int64_t)sm_count() * 8 / cb;
  if (want < 1) want = 1;
  int64_t rpb = (M + want - 1) / want;
  if (rpb < 64) rpb = 64;
  int64_t nb = (M + rpb - 1) / rpb;
  dim3 grid(cb, (unsigned)nb), block(32, CS_ROWS);
  if (dot)
    col_reduce_kernel<true><<<grid, block, 0, st>>>(A, Zn, mean, rstd, M, H, rpb, o1, o2);
  else
    col_reduce_kernel<false><<<grid, block, 0, st>>>(A, nullptr, nullptr, nullptr, M, H, rpb, o1, o2);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

// ------------------------------------------------------------------ BN fold / backward apply
__global__ void bn_fold_train_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq,
                                     int64_t M, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float momentum,
                                     float* running_mean, float* running_var, float* mean,
                                     float* rstd, float* scale, float* shift, int H) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  double mu = sum[h] / (double)M;
  double var = sumsq[h] / (double)M - mu * mu;  // biased; fp64 so the subtraction is safe
  if (var < 0.0) var = 0.0;
  float muf = (float)mu;
  float rs = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma ? gamma[h] : 1.0f;
  float b = beta ? beta[h] : 0.0f;
  float sc = g * rs;
  mean[h] = muf;
  rstd[h] = rs;
  scale[h] = sc;
  shift[h] = b - muf * sc;
  if (running_mean) {
    double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    running_mean[h] = (1.0f - momentum) * running_mean[h] + momentum * muf;
    running_var[h] = (1.0f - momentum) * running_var[h] + momentum * (float)unb;
  }
}

__global__ void bn_bwd_apply_kernel(float* __restrict__ dI, const float* __restrict__ Z,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ scale, const double* __restrict__ s1,
                                    const double* __restrict__ s2, int64_t M, int H, uint32_t* __restrict__ amax) {
  const double invM = 1.0 / (double)M;
  float mx = 0.f;
  if ((H & 3) == 0) {
    // a thread keeps its 4 columns for the whole grid-stride walk when the stride is a multiple of H/4
    const int64_t h4 = H / 4, n4 = M * h4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int h = (int)(i % h4) * 4;
      const float4 mu = *reinterpret_cast<const float4*>(mean + h), rs = *reinterpret_cast<const float4*>(rstd + h);
      const float4 sc = *reinterpret_cast<const float4*>(scale + h);
      const float4 z = reinterpret_cast<const float4*>(Z)[i];
      float4 d = reinterpret_cast<float4*>(dI)[i];
      d.x = sc.x * (d.x - (float)(s1[h] * invM) - (z.x - mu.x) * rs.x * (float)(s2[h] * invM));
      d.y = sc.y * (d.y - (float)(s1[h + 1] * invM) - (z.y - mu.y) * rs.y * (float)(s2[h + 1] * invM));
      d.z = sc.z * (d.z - (float)(s1[h + 2] * invM) - (z.z - mu.z) * rs.z * (float)(s2[h + 2] * invM));
      d.w = sc.w * (d.w - (float)(s1[h + 3] * invM) - (z.w - mu.w) * rs.w * (float)(s2[h + 3] * invM));
      reinterpret_cast<float4*>(dI)[i] = d;
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(d.x), fabsf(d.y))), fmaxf(fabsf(d.z), fabsf(d.w)));
    }
  } else {
    int64_t n = M * (int64_t)H;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
      int h = (int)(i % H);
      float c1 = (float)(s1[h] * invM), c2 = (float)(s2[h] * invM);
      float xh = (Z[i] - mean[h]) * rstd[h];
      const float d = scale[h] * (dI[i] - c1 - xh * c2);
      dI[i] = d;
      mx = fmaxf(mx, fabsf(d));
    }
  }
  if (amax) {  // max|dZ| for the fp16 split of the weight- and data-gradient GEMMs (sparch_split_f16)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(amax, __float_as_uint(mx));
  }
}

// Bidirectional layer without a flipped copy (snns.py:666-668 feeds W x twice, once reversed): the gradient of W x is the
// sum of the two passes' dZ, dZ[b][t] = scale ((dI[b][t] - c1 - xhat c2) + (dI[b + B][T-1-t] - c1 - xhat c2)) with ONE
// xhat (the same row of Z) and the column sums taken over all 2 B T rows.  In place on the first half of dI.
__global__ void bn_bwd_apply_bidir_kernel(float* __restrict__ dI, const float* __restrict__ Z,
                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                          const float* __restrict__ scale, const double* __restrict__ s1,
                                          const double* __restrict__ s2, int64_t M, int H, int T) {
  const double inv = 1.0 / (double)(2 * M);
  const int64_t n = M * (int64_t)H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / H;
    const int h = (int)(i - r * H);
    const int64_t b = r / T;
    const int64_t rp = M + b * T + (T - 1 - (r - b * T));
    const float c1 = (float)(s1[h] * inv), c2 = (float)(s2[h] * inv);
    const float xh = (Z[i] - mean[h]) * rstd[h];
    dI[i] = scale[h] * ((dI[i] - c1 - xh * c2) + (dI[rp * H + h] - c1 - xh * c2));
  }
}

// bound[0] = bit pattern of an upper bound of max|dZ|: max_h |scale_h| * (max|dI| + max_h |c1_h| + sqrt(M) max_h |c2_h|)
// (|xhat| <= sqrt(M - 1) for batch statistics).  The bound only positions the fp16 scale; an fp16 hi/lo pair keeps
// its 22 bits over 2^-20 of the scaled range, so a bound that is loose by two orders of magnitude costs nothing.
__global__ void __launch_bounds__(1024)
bn_bwd_bound_kernel(const float* __restrict__ scale, const double* __restrict__ s1, const double* __restrict__ s2,
                    int64_t M, int H, const uint32_t* __restrict__ amax_dI, uint32_t* __restrict__ bound,
                    float* __restrict__ coef, const float pair = 1.0f) {   // pair = 2: dZ sums two rows (bidirectional)
  __shared__ float sh[3][32];
  float ms = 0.f, m1 = 0.f, m2 = 0.f;
  const double invM = 1.0 / (double)M;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float c1 = (float)(s1[h] * invM), c2 = (float)(s2[h] * invM);
    coef[h] = c1;          // the per-column constants of the apply pass, converted once
    coef[H + h] = c2;
    coef[2 * H + h] = (float)s1[h];   // = d(beta) and d(gamma) of the BatchNorm affine as fp32 (no conversion launches)
    coef[3 * H + h] = (float)s2[h];
    ms = fmaxf(ms, fabsf(scale[h]));
    m1 = fmaxf(m1, fabsf(c1));
    m2 = fmaxf(m2, fabsf(c2));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = ms; sh[1][threadIdx.x >> 5] = m1; sh[2][threadIdx.x >> 5] = m2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      ms = fmaxf(ms, sh[0][w]); m1 = fmaxf(m1, sh[1][w]); m2 = fmaxf(m2, sh[2][w]);
    }
    const float b = pair * ms * (__uint_as_float(*amax_dI) + m1 + sqrtf((float)M) * m2);
    *bound = __float_as_uint(b);
  }
}

// dZ = scale (dI - c1 - xhat c2) written straight as the two scaled fp16 terms the gradient GEMMs read (and as
// fp32 only if dZ32 is given): the fp32 dZ tensor and the split pass over it disappear.
__global__ void __launch_bounds__(256)
bn_bwd_apply_f16_kernel(const float* __restrict__ dI, const float* __restrict__ Z, const float* __restrict__ mean,
                        const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ coef,
                        int64_t M, int H, const uint32_t* __restrict__ bound,
                        __half* __restrict__ P0, __half* __restrict__ P1, int64_t ldp, float* __restrict__ dZ32,
                        const int pairT = 0) {
  // pairT = T > 0: bidirectional layer, M = B T output rows; row (b, t) also takes dI of row (B + b, T - 1 - t) (see
  // bn_bwd_apply_bidir_kernel; coef holds the column means over all 2 M rows)
  const float sc2 = ldexpf(1.0f, f16_scale_exp(*bound));
  const int64_t segs = ldp / 8, n = M * segs;
  const bool vec = (H & 3) == 0;
  const float* c1 = coef;
  const float* c2 = coef + H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / segs;
    const int c = (int)(i - r * segs) * 8;
    float d[8], z[8];
    const float* dp = dI + r * H + c;
    const float* zp = Z + r * H + c;
    float d2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d2[j] = 0.f;
    if (vec && c + 8 <= H) {
      const float4 a = *reinterpret_cast<const float4*>(dp), b = *reinterpret_cast<const float4*>(dp + 4);
      const float4 e = *reinterpret_cast<const float4*>(zp), f = *reinterpret_cast<const float4*>(zp + 4);
      d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
      z[0] = e.x; z[1] = e.y; z[2] = e.z; z[3] = e.w; z[4] = f.x; z[5] = f.y; z[6] = f.z; z[7] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        d[j] = c + j < H ? dp[j] : 0.f;
        z[j] = c + j < H ? zp[j] : 0.f;
      }
    }
    if (pairT) {
      const int64_t b = r / pairT;
      const float* qp = dI + (M + b * pairT + (pairT - 1 - (r - b * pairT))) * H + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) d2[j] = c + j < H ? qp[j] : 0.f;
    }
    float pm[8], pr[8], ps[8], p1[8], p2[8];
    if (vec && c + 8 <= H) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(mean + c + 4 * q), b = *reinterpret_cast<const float4*>(rstd + c + 4 * q);
        const float4 e = *reinterpret_cast<const float4*>(scale + c + 4 * q), f = *reinterpret_cast<const float4*>(c1 + c + 4 * q);
        const float4 g = *reinterpret_cast<const float4*>(c2 + c + 4 * q);
        pm[4 * q] = a.x; pm[4 * q + 1] = a.y; pm[4 * q + 2] = a.z; pm[4 * q + 3] = a.w;
        pr[4 * q] = b.x; pr[4 * q + 1] = b.y; pr[4 * q + 2] = b.z; pr[4 * q + 3] = b.w;
        ps[4 * q] = e.x; ps[4 * q + 1] = e.y; ps[4 * q + 2] = e.z; ps[4 * q + 3] = e.w;
        p1[4 * q] = f.x; p1[4 * q + 1] = f.y; p1[4 * q + 2] = f.z; p1[4 * q + 3] = f.w;
        p2[4 * q] = g.x; p2[4 * q + 1] = g.y; p2[4 * q + 2] = g.z; p2[4 * q + 3] = g.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int h = min(c + j, H - 1);
        pm[j] = mean[h]; pr[j] = rstd[h]; ps[j] = scale[h]; p1[j] = c1[h]; p2[j] = c2[h];
      }
    }
    __align__(16) __half h0[8], h1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (z[j] - pm[j]) * pr[j];
      float x = c + j < H ? ps[j] * (d[j] - p1[j] - xh * p2[j]) : 0.f;
      if (pairT) x = c + j < H ? ps[j] * ((d[j] - p1[j] - xh * p2[j]) + (d2[j] - p1[j] - xh * p2[j])) : 0.f;
      d[j] = x;
      const float v = x * sc2;
      h0[j] = __float2half_rn(v);
      h1[j] = __float2half_rn(v - __half2float(h0[j]));
    }
    *reinterpret_cast<uint4*>(P0 + r * ldp + c) = *reinterpret_cast<const uint4*>(h0);
    *reinterpret_cast<uint4*>(P1 + r * ldp + c) = *reinterpret_cast<const uint4*>(h1);
    if (dZ32) {
      float* op = dZ32 + r * H + c;
      if (vec && c + 8 <= H) {
        *reinterpret_cast<float4*>(op) = make_float4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(d[4], d[5], d[6], d[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + j < H) op[j] = d[j];
      }
    }
  }
}

// The same pass for the common layout (H % 8 == 0, 16-byte aligned rows, 256 % (H / 8) == 0): a thread KEEPS its 8 columns
// and walks rows, so the five per-column constants are loaded once instead of with every item (10 of an item's 16 load
// instructions in the generic kernel, which ncu showed latency-bound at 48 % of the DRAM throughput), and two rows are in
// flight per thread.
template <bool PAIR>
__global__ void __launch_bounds__(256)
bn_bwd_apply_f16_rows_kernel(const float* __restrict__ dI, const float* __restrict__ Z, const float* __restrict__ mean,
                             const float* __restrict__ rstd, const float* __restrict__ scale,
                             const float* __restrict__ coef, int64_t M, int H, const uint32_t* __restrict__ bound,
                             __half* __restrict__ P0, __half* __restrict__ P1, float* __restrict__ dZ32, int T) {
  const float sc2 = ldexpf(1.0f, f16_scale_exp(*bound));
  const int segs = H / 8, ty = 256 / segs;
  const int c = (threadIdx.x % segs) * 8, ry = threadIdx.x / segs;
  float pm[8], pr[8], ps[8], p1[8], p2[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(mean + c + 4 * q), b = *reinterpret_cast<const float4*>(rstd + c + 4 * q);
    const float4 e = *reinterpret_cast<const float4*>(scale + c + 4 * q), f = *reinterpret_cast<const float4*>(coef + c + 4 * q);
    const float4 g = *reinterpret_cast<const float4*>(coef + H + c + 4 * q);
    pm[4 * q] = a.x; pm[4 * q + 1] = a.y; pm[4 * q + 2] = a.z; pm[4 * q + 3] = a.w;
    pr[4 * q] = b.x; pr[4 * q + 1] = b.y; pr[4 * q + 2] = b.z; pr[4 * q + 3] = b.w;
    ps[4 * q] = e.x; ps[4 * q + 1] = e.y; ps[4 * q + 2] = e.z; ps[4 * q + 3] = e.w;
    p1[4 * q] = f.x; p1[4 * q + 1] = f.y; p1[4 * q + 2] = f.z; p1[4 * q + 3] = f.w;
    p2[4 * q] = g.x; p2[4 * q + 1] = g.y; p2[4 * q + 2] = g.z; p2[4 * q + 3] = g.w;
  }
  const int64_t stride = (int64_t)gridDim.x * ty;
  for (int64_t r0 = (int64_t)blockIdx.x * ty + ry; r0 < M; r0 += 2 * stride) {
    float4 dv[2][2], zv[2][2], qv[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < M) {
        const float* dp = dI + r * H + c;
        const float* zp = Z + r * H + c;
        dv[u][0] = *reinterpret_cast<const float4*>(dp); dv[u][1] = *reinterpret_cast<const float4*>(dp + 4);
        zv[u][0] = *reinterpret_cast<const float4*>(zp); zv[u][1] = *reinterpret_cast<const float4*>(zp + 4);
        if (PAIR) {
          const int64_t b = r / T;
          const float* qp = dI + (M + b * T + (T - 1 - (r - b * T))) * H + c;
          qv[u][0] = *reinterpret_cast<const float4*>(qp); qv[u][1] = *reinterpret_cast<const float4*>(qp + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= M) continue;
      const float d[8] = {dv[u][0].x, dv[u][0].y, dv[u][0].z, dv[u][0].w, dv[u][1].x, dv[u][1].y, dv[u][1].z, dv[u][1].w};
      const float z[8] = {zv[u][0].x, zv[u][0].y, zv[u][0].z, zv[u][0].w, zv[u][1].x, zv[u][1].y, zv[u][1].z, zv[u][1].w};
      float d2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (PAIR) {
        d2[0] = qv[u][0].x; d2[1] = qv[u][0].y; d2[2] = qv[u][0].z; d2[3] = qv[u][0].w;
        d2[4] = qv[u][1].x; d2[5] = qv[u][1].y; d2[6] = qv[u][1].z; d2[7] = qv[u][1].w;
      }
      float x[8];
      __align__(16) __half h0[8], h1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // the generic kernel's expressions, operation for operation
        const float xh = (z[j] - pm[j]) * pr[j];
        x[j] = PAIR ? ps[j] * ((d[j] - p1[j] - xh * p2[j]) + (d2[j] - p1[j] - xh * p2[j])) : ps[j] * (d[j] - p1[j] - xh * p2[j]);
        const float v = x[j] * sc2;
        h0[j] = __float2half_rn(v);
        h1[j] = __float2half_rn(v - __half2float(h0[j]));
      }
      *reinterpret_cast<uint4*>(P0 + r * H + c) = *reinterpret_cast<const uint4*>(h0);
      *reinterpret_cast<uint4*>(P1 + r * H + c) = *reinterpret_cast<const uint4*>(h1);
      if (dZ32) {
        float* op = dZ32 + r * H + c;
        *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
      }
    }
  }
}

static bool bn_apply_rows_ok(const void* dI, const void* Z, const void* P0, const void* P1, const void* dZ32, int H,
                             int64_t ldp) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(dI) | reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(P0) |
                       reinterpret_cast<uintptr_t>(P1) | reinterpret_cast<uintptr_t>(dZ32);
  return (H % 8) == 0 && ldp == H && H / 8 <= 256 && 256 % (H / 8) == 0 && (al & 15) == 0;
}

static int ew_grid(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ------------------------------------------------------------------ LayerNorm (normalization="layernorm", snns.py:98-99)
// nn.LayerNorm(H) over the last dimension of W x (snns.py:678-680): y = (x - mean) * rstd * gamma + beta with the
// biased row variance.  One warp per row: the row is read once into registers (H <= 32 * LN_MAX values per lane),
// two-pass mean / variance like ATen's RowwiseMoments for short rows, mean and rstd kept for the backward.
constexpr int LN_MAX = 64;   // row length <= 2048

__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int64_t M, int H,
                                                     float* __restrict__ Y, float* __restrict__ mean, float* __restrict__ rstd) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* x = X + row * H;
  float v[LN_MAX];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < H ? x[c] : 0.f;
    s += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX; ++i) {
    const float d = (lane + 32 * i < H) ? v[i] - mu : 0.f;
    q += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = rsqrtf(q / (float)H + eps);
  float* y = Y + row * H;
#pragma unroll
  for (int i = 0; i < LN_MAX; ++i) {
    const int c = lane + 32 * i;
    if (c < H) {
      const float xh = (v[i] - mu) * rs;
      y[c] = (gamma ? xh * gamma[c] : xh) + (beta ? beta[c] : 0.f);
    }
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)), g = dy * gamma; one warp per row
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, int64_t M, int H,
                                                        float* __restrict__ dX) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float mu = mean[row], rs = rstd[row];
  const float* x = X + row * H;
  const float* dy = dY + row * H;
  float g[LN_MAX], xh[LN_MAX];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX; ++i) {
    const int c = lane + 32 * i;
    const bool in = c < H;
    xh[i] = in ? (x[c] - mu) * rs : 0.f;
    g[i] = in ? dy[c] * (gamma ? gamma[c] : 1.f) : 0.f;
    s1 += g[i];
    s2 += g[i] * xh[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 / (float)H, m2 = s2 / (float)H;
  float* dx = dX + row * H;
#pragma unroll
  for (int i = 0; i < LN_MAX; ++i) {
    const int c = lane + 32 * i;
    if (c < H) dx[c] = rs * (g[i] - m1 - xh[i] * m2);
  }
}

// dgamma[c] = sum_rows dy * xhat, dbeta[c] = sum_rows dy: a thread per column, a block per slab of rows, per-block partials
// in fp64 written to part[2][nblocks][H] and summed in block order by ln_bwd_finish_kernel (deterministic)
__global__ void __launch_bounds__(256) ln_bwd_param_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           int64_t M, int H, int rows_per_block, double* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  double sg = 0.0, sb = 0.0;
  for (int64_t r = r0; r < r1; ++r) {
    const float dy = dY[r * H + c];
    sg += (double)(dy * ((X[r * H + c] - mean[r]) * rstd[r]));
    sb += (double)dy;
  }
  part[((size_t)0 * gridDim.y + blockIdx.y) * H + c] = sg;
  part[((size_t)1 * gridDim.y + blockIdx.y) * H + c] = sb;
}

__global__ void ln_bwd_finish_kernel(const double* __restrict__ part, int nb, int H, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  double sg = 0.0, sb = 0.0;
  for (int b = 0; b < nb; ++b) {
    sg += part[((size_t)0 * nb + b) * H + c];
    sb += part[((size_t)1 * nb + b) * H + c];
  }
  if (dgamma) dgamma[c] = (float)sg;
  if (dbeta) dbeta[c] = (float)sb;
}

}  // namespace sparch

using namespace sparch;

extern "C" {

const char* sparch_last_error(void) { return g_err.c_str(); }
int sparch_abi_version(void) { return 1; }

int sparch_boxcar_fwd(const float* x, float* s, int64_t n, sparch_stream_t st) {
  SPARCH_REQUIRE(n >= 0 && (n == 0 || (x && s)), "null pointer");
  if (n == 0) return SPARCH_OK;
  boxcar_fwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(st)>>>(x, s, n);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_boxcar_bwd(const float* x, const float* gs, float* gx, int64_t n, sparch_stream_t st) {
  SPARCH_REQUIRE(n >= 0 && (n == 0 || (x && gs && gx)), "null pointer");
  if (n == 0) return SPARCH_OK;
  boxcar_bwd_kernel<<<ew_grid(n, 256), 256, 0, as_stream(st)>>>(x, gs, gx, n);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_col_stats(const float* Z, int64_t M, int H, double* sum, double* sumsq,
                     sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && sum && sumsq && (M == 0 || Z), "bad shape or null pointer");
  return col_reduce_launch(false, Z, nullptr, nullptr, nullptr, M, H, sum, sumsq, as_stream(st));
}

int sparch_col_dot(const float* A, const float* Zn, const float* mean, const float* rstd, int64_t M,
                   int H, double* sum1, double* sum2, uint32_t* amax_a, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && sum1 && sum2 && (M == 0 || (A && Zn)), "bad shape or null pointer");
  SPARCH_REQUIRE((mean == nullptr) == (rstd == nullptr), "mean and rstd go together");
  return col_reduce_launch(true, A, Zn, mean, rstd, M, H, sum1, sum2, as_stream(st), amax_a);
}

int sparch_col_dot_bidir(const float* A, const float* Zn, const float* mean, const float* rstd, int64_t M, int H, int T,
                         int64_t rev_from, double* sum1, double* sum2, uint32_t* amax_a, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && T > 0 && rev_from > 0 && M == 2 * rev_from * T && sum1 && sum2 && (M == 0 || (A && Zn)),
                 "bad shape or null pointer");
  SPARCH_REQUIRE((mean == nullptr) == (rstd == nullptr), "mean and rstd go together");
  return col_reduce_launch(true, A, Zn, mean, rstd, M, H, sum1, sum2, as_stream(st), amax_a, rev_from, T);
}

int sparch_bn_fold_train(const double* sum, const double* sumsq, int64_t M, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, float* mean, float* rstd, float* scale, float* shift,
                         int H, sparch_stream_t st) {
  SPARCH_REQUIRE(M > 0 && H > 0 && sum && sumsq && mean && rstd && scale && shift, "bad argument");
  SPARCH_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "running stats go together");
  bn_fold_train_kernel<<<(H + 127) / 128, 128, 0, as_stream(st)>>>(
      sum, sumsq, M, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift, H);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_bn_bwd_apply(float* dI, const float* Z, const float* mean, const float* rstd,
                        const float* scale, const double* sum1, const double* sum2, int64_t M, int H,
                        uint32_t* amax, sparch_stream_t st) {
  SPARCH_REQUIRE(M > 0 && H > 0 && dI && Z && mean && rstd && scale && sum1 && sum2, "bad argument");
  if (amax) SPARCH_CUDA(cudaMemsetAsync(amax, 0, sizeof(uint32_t), as_stream(st)));
  bn_bwd_apply_kernel<<<ew_grid(M * (int64_t)H / 4, 256), 256, 0, as_stream(st)>>>(dI, Z, mean, rstd, scale,
                                                                                  sum1, sum2, M, H, amax);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_bn_bwd_apply_f16(const float* dI, const float* Z, const float* mean, const float* rstd, const float* scale,
                            const double* sum1, const double* sum2, int64_t M, int H, const uint32_t* amax_dI,
                            uint32_t* bound, float* coef, void* P0, void* P1, int64_t ldp, float* dZ32,
                            sparch_stream_t st) {
  SPARCH_REQUIRE(M > 0 && H > 0 && dI && Z && mean && rstd && scale && sum1 && sum2 && amax_dI && bound && coef &&
                     P0 && P1,
                 "bad argument");
  SPARCH_REQUIRE((ldp % 8) == 0 && ldp >= H, "ldp must be a multiple of 8 covering a row");
  bn_bwd_bound_kernel<<<1, 1024, 0, as_stream(st)>>>(scale, sum1, sum2, M, H, amax_dI, bound, coef);
  SPARCH_LAUNCH_OK();
  if (bn_apply_rows_ok(dI, Z, P0, P1, dZ32, H, ldp)) {
    const int ty = 256 / (H / 8);
    int64_t nb = (M + 2 * ty - 1) / (2 * ty), cap = (int64_t)sm_count() * 8;
    bn_bwd_apply_f16_rows_kernel<false><<<(unsigned)(nb < cap ? nb : cap), 256, 0, as_stream(st)>>>(
        dI, Z, mean, rstd, scale, coef, M, H, bound, (__half*)P0, (__half*)P1, dZ32, 1);
  } else {
    bn_bwd_apply_f16_kernel<<<ew_grid(M * (ldp / 8), 256), 256, 0, as_stream(st)>>>(
        dI, Z, mean, rstd, scale, coef, M, H, bound, (__half*)P0, (__half*)P1, ldp, dZ32);
  }
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_bn_bwd_apply_bidir(float* dI, const float* Z, const float* mean, const float* rstd, const float* scale,
                              const double* sum1, const double* sum2, int64_t M, int H, int T, sparch_stream_t st) {
  SPARCH_REQUIRE(M > 0 && H > 0 && T > 0 && M % T == 0 && dI && Z && mean && rstd && scale && sum1 && sum2, "bad argument");
  bn_bwd_apply_bidir_kernel<<<ew_grid(M * (int64_t)H, 256), 256, 0, as_stream(st)>>>(dI, Z, mean, rstd, scale, sum1, sum2,
                                                                                    M, H, T);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_bn_bwd_apply_f16_bidir(const float* dI, const float* Z, const float* mean, const float* rstd,
                                  const float* scale, const double* sum1, const double* sum2, int64_t M, int H, int T,
                                  const uint32_t* amax_dI, uint32_t* bound, float* coef, void* P0, void* P1,
                                  int64_t ldp, float* dZ32, sparch_stream_t st) {
  SPARCH_REQUIRE(M > 0 && H > 0 && T > 0 && M % T == 0 && dI && Z && mean && rstd && scale && sum1 && sum2 && amax_dI &&
                     bound && coef && P0 && P1,
                 "bad argument");
  SPARCH_REQUIRE((ldp % 8) == 0 && ldp >= H, "ldp must be a multiple of 8 covering a row");
  // column means over all 2 M rows of the two passes
  bn_bwd_bound_kernel<<<1, 1024, 0, as_stream(st)>>>(scale, sum1, sum2, 2 * M, H, amax_dI, bound, coef, 2.0f);
  SPARCH_LAUNCH_OK();
  if (bn_apply_rows_ok(dI, Z, P0, P1, dZ32, H, ldp)) {
    const int ty = 256 / (H / 8);
    int64_t nb = (M + 2 * ty - 1) / (2 * ty), cap = (int64_t)sm_count() * 8;
    bn_bwd_apply_f16_rows_kernel<true><<<(unsigned)(nb < cap ? nb : cap), 256, 0, as_stream(st)>>>(
        dI, Z, mean, rstd, scale, coef, M, H, bound, (__half*)P0, (__half*)P1, dZ32, T);
  } else {
    bn_bwd_apply_f16_kernel<<<ew_grid(M * (ldp / 8), 256), 256, 0, as_stream(st)>>>(
        dI, Z, mean, rstd, scale, coef, M, H, bound, (__half*)P0, (__half*)P1, ldp, dZ32, T);
  }
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_layernorm_fwd(const float* X, const float* gamma, const float* beta, float eps, int64_t M, int H, float* Y,
                         float* mean, float* rstd, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && H <= 32 * LN_MAX, "row length must be 1..2048");
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(X && Y && mean && rstd, "null pointer");
  ln_fwd_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, as_stream(st)>>>(X, gamma, beta, eps, M, H, Y, mean, rstd);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

size_t sparch_layernorm_bwd_workspace(int64_t M, int H) {
  const int64_t nb = M < 148 * 4 ? (M > 0 ? M : 1) : 148 * 4;
  return (size_t)2 * nb * H * sizeof(double);
}

int sparch_layernorm_bwd(const float* dY, const float* X, const float* gamma, const float* mean, const float* rstd,
                         int64_t M, int H, float* dX, float* dgamma, float* dbeta, void* workspace, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && H <= 32 * LN_MAX, "row length must be 1..2048");
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(dY && X && mean && rstd && dX, "null pointer");
  ln_bwd_dx_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, as_stream(st)>>>(dY, X, gamma, mean, rstd, M, H, dX);
  SPARCH_LAUNCH_OK();
  if (dgamma || dbeta) {
    SPARCH_REQUIRE(workspace, "the parameter gradients need the workspace");
    const int nb = (int)(M < 148 * 4 ? M : 148 * 4);
    const int rpb = (int)((M + nb - 1) / nb);
    const int nbu = (int)((M + rpb - 1) / rpb);
    double* part = reinterpret_cast<double*>(workspace);
    ln_bwd_param_kernel<<<dim3((H + 255) / 256, nbu), 256, 0, as_stream(st)>>>(dY, X, mean, rstd, M, H, rpb, part);
    SPARCH_LAUNCH_OK();
    ln_bwd_finish_kernel<<<(H + 255) / 256, 256, 0, as_stream(st)>>>(part, nbu, H, dgamma, dbeta);
    SPARCH_LAUNCH_OK();
  }
  return SPARCH_OK;
}

}  // extern "C"

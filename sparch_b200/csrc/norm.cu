// Boxcar surrogate (elementwise) and the BatchNorm1d pieces around the projection:
// column statistics, train-mode fold to (scale, shift), backward reductions and apply.
// All are HBM-bound streaming kernels: 16-byte loads where the row length allows, grids
// sized from the SM count, fp64 accumulation for the column sums.
#include <stdarg.h>

#include <string>

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

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
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

static int col_reduce_launch(bool dot, const float* A, const float* Zn, const float* mean,
                             const float* rstd, int64_t M, int H, double* o1, double* o2,
                             cudaStream_t st) {
  SPARCH_CUDA(cudaMemsetAsync(o1, 0, sizeof(double) * H, st));
  SPARCH_CUDA(cudaMemsetAsync(o2, 0, sizeof(double) * H, st));
  if (M == 0) return SPARCH_OK;
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

static int ew_grid(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
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
                   int H, double* sum1, double* sum2, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && H > 0 && sum1 && sum2 && (M == 0 || (A && Zn)), "bad shape or null pointer");
  SPARCH_REQUIRE((mean == nullptr) == (rstd == nullptr), "mean and rstd go together");
  return col_reduce_launch(true, A, Zn, mean, rstd, M, H, sum1, sum2, as_stream(st));
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

}  // extern "C"

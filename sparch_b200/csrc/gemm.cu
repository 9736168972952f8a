// Time-parallel GEMMs of the SNN stack on the 5th-generation tensor cores (tcgen05 + TMEM + TMA):
//   projection  Z  = X  W^T          (snns.py:675)        M = Be*T, N = H,   K = Fin
//   data grad   dX = dZ W            (autograd of :675)   M = Be*T, N = Fin, K = H
//   weight grad dW = dZ^T X          (autograd of :675)   M = H,    N = Fin, K = Be*T
//   recur. grad dV = S_prev^T dI     (autograd of :720)   M = H,    N = H,   K = Be*T
// All are computed as C[M,N] = alpha * sum_pairs A_i[M,K] . B_j[N,K]^T with both operands K-major
// bf16.  fp32 accuracy comes from splitting an fp32 operand into up to three bf16 parts
// (x = x0 + x1 + x2, 24 mantissa bits); a spike operand is exact in one part.  The products of
// all requested (i, j) part pairs accumulate into ONE fp32 TMEM accumulator, so the split costs
// tensor-pipe passes but no extra traffic on C.
//
// Kernel: one 128 x 256 output tile per CTA (cta_group::1, UMMA 128x256x16, full-rate shape),
// 4-stage TMA -> shared (SWIZZLE_128B) ring, warp-specialised: warp 0 TMA producer, warp 1 MMA
// issuer (one elected lane) + TMEM allocator, warps 2..5 epilogue (tcgen05.ld -> registers ->
// global).  Split-K over blockIdx.z writes fp32 partial tiles that a small kernel reduces in a
// fixed order (deterministic).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05_utils.cuh"

namespace sparch {

constexpr int GM = 128, GN = 256, GK = 64, GSTAGES = 4;
constexpr int G_A_BYTES = GM * GK * 2, G_B_BYTES = GN * GK * 2, G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_THREADS = 192;
constexpr uint32_t G_TMEM_COLS = 512;   // two fp32 accumulators of 256 columns
constexpr int G_EP_PITCH = 36;           // floats per staged row: 16-byte accesses stay conflict-free
constexpr int G_EP_BYTES = 4 * 32 * G_EP_PITCH * 4;  // one 32 x 32 staging tile per epilogue warp
constexpr size_t G_SMEM = (size_t)GSTAGES * G_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + G_EP_BYTES;

struct TmapSet {
  CUtensorMap a[3];
  CUtensorMap b[3];
};

struct GemmParams {
  int M, N, K;
  int npairs;
  int pair_a[8], pair_b[8];
  int kblocks;         // ceil(K / 64)
  int kb_per_split;    // k-blocks handled by one split
  int splits;          // split-K factor (work items = splits x tiles)
  int tile_n;          // output tile width = UMMA N: 256, or N rounded up to 16 when the whole output is narrower
                       // (readout projection N = 35, input-layer dW N = 40: a 256-wide tile would be 80 % padding)
  float* C;            // final output (splits == 1) or partial buffer [splits][M][ldc]
  long long ldc;
  long long split_stride;
  float alpha;
  const float* bias;   // per column n, may be NULL (applied only when splits == 1)
  double* stat_sum;    // optional per-column sum / sum of squares of the OUTPUT (BatchNorm statistics fused
  double* stat_sumsq;  // into the epilogue, snns.py:678-680); only with splits == 1
  int a_mn, b_mn;      // operand is MN-major in memory: a (K, MN) row-major matrix (weight-gradient GEMMs)
  int a_koff;          // added to A's K coordinate (TMA zero-fills out-of-range rows): S_prev = S delayed by one frame
  int fp16;            // operand terms are fp16 (scaled hi/lo pairs, sparch_split_f16) instead of bf16
  const uint32_t* amax_a;  // fp16 terms: bit pattern of max|x| of the split tensor (its power-of-two scale is
  const uint32_t* amax_b;  // undone in the epilogue); NULL = unscaled
};

// MN-major, SWIZZLE_128B operand tile as TMA delivers it from a (K, MN) row-major matrix: per block of
// 64 MN elements, 64 K rows of 128 bytes (8 KB); blocks of 64 MN elements follow each other.
//   leading byte offset = distance between 64-element MN blocks, stride byte offset = distance between
//   groups of 8 K rows (cute::UMMA canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;                   // leading byte offset: next 64-wide MN block
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: next 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 -> fp32 accumulator, M = 128, N = 256; bits 7 / 10 = A / B format (0 fp16, 1 bf16), bit 15 / 16 = A / B
// is MN-major
constexpr uint32_t G_IDESC = (1u << 4) | ((uint32_t)(GM >> 4) << 24);   // N field (bits 17..22) = tile_n >> 3, set per launch
constexpr uint32_t G_IDESC_BF16 = (1u << 7) | (1u << 10);

// Column sums across the 32 lanes of a warp for 32 per-lane values in 31 shuffles: at each step a lane
// keeps one half of its values and receives the partner's copy of that half; lane l ends up with the
// total of column (bit-reversed bookkeeping folded into the index arithmetic): returns column `l`.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int half = 16 >> step;            // values kept per lane after this step
    const int bit = 16 >> step;             // lane bit deciding which half is kept
    const bool upper = lane & bit;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? v[i + half] : v[i];
      const float give = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, give, bit);
    }
  }
  return v[0];  // lane l holds the sum of column l: bit (16>>s) of l selected the upper half at step s
}

// Persistent kernel: gridDim.x = min(#work items, #SMs) CTAs walk the (split, m-tile, n-tile) items with a stride
// of gridDim.x.  The fp32 accumulator is double-buffered in TMEM (2 x 256 columns): while the epilogue warps
// drain tile j (tcgen05.ld -> alpha / bias / statistics -> global), the MMA warp already accumulates tile j + 1,
// and the TMA ring keeps its phase across tiles, so only the first tile's prologue and the last tile's epilogue
// are exposed.
template <bool STATS>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tn_bf16_kernel(const __grid_constant__ TmapSet maps, const GemmParams p) {
  extern __shared__ unsigned char gsm_raw[];
  __shared__ float2 sstat[STATS ? 4 : 1][STATS ? GN : 1];  // per row quarter: column (sum, sum of squares) of the tile
  const uint32_t raw = smem_u32(gsm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* gsm = gsm_raw + (base - raw);
  constexpr int NST = GSTAGES;
  const uint32_t bars = base + NST * G_STAGE_BYTES;             // full[4], empty[4], tmem_full[2], tmem_empty[2]
  constexpr int B_TFULL = 2 * GSTAGES, B_TEMPTY = 2 * GSTAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gsm + NST * G_STAGE_BYTES + 128);
  float* ep_stage = reinterpret_cast<float*>(gsm + NST * G_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TN = p.tile_n;
  const int tiles_n = (p.N + TN - 1) / TN, tiles = tiles_n * ((p.M + GM - 1) / GM);
  const uint32_t b_bytes = p.b_mn ? (uint32_t)((TN + 63) / 64) * 8192u : (uint32_t)TN * 128u;
  const int items = tiles * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (GSTAGES + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bars + 8 * (B_TFULL + b), 1);    // one tcgen05.commit
      mbar_init(bars + 8 * (B_TEMPTY + b), 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(G_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int z = item / tiles, tile = item - z * tiles;
        const int m0 = (tile / tiles_n) * GM, n0 = (tile % tiles_n) * TN;
        const int kb0 = z * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        for (int pr = 0; pr < p.npairs; ++pr) {
          const CUtensorMap* ma = &maps.a[p.pair_a[pr]];
          const CUtensorMap* mb = &maps.b[p.pair_b[pr]];
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % NST;
            const uint32_t ph = (it / NST) & 1;
            mbar_wait(bars + 8 * (GSTAGES + s), ph ^ 1);
            mbar_expect_tx(bars + 8 * s, G_A_BYTES + b_bytes);
            const uint32_t sa = base + s * G_STAGE_BYTES;
            if (!p.a_mn) {
              tma_load_2d(sa, ma, kb * GK, m0, bars + 8 * s);
            } else {
#pragma unroll
              for (int blk = 0; blk < GM / 64; ++blk)
                tma_load_2d(sa + blk * 8192, ma, m0 + blk * 64, kb * GK + p.a_koff, bars + 8 * s);
            }
            if (!p.b_mn) {
              tma_load_2d(sa + G_A_BYTES, mb, kb * GK, n0, bars + 8 * s);
            } else {
#pragma unroll
              for (int blk = 0; blk < (TN + 63) / 64; ++blk)
                tma_load_2d(sa + G_A_BYTES + blk * 8192, mb, n0 + blk * 64, kb * GK, bars + 8 * s);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // K advance per UMMA (16 elements): K-major 32 bytes inside the swizzle atom (2 units of 16 B),
      // MN-major 16 rows of 128 bytes (128 units)
      const uint32_t ka = p.a_mn ? 128u : 2u, kb_ = p.b_mn ? 128u : 2u;
      const uint32_t idesc = G_IDESC | ((uint32_t)(TN >> 3) << 17) | (p.fp16 ? 0u : G_IDESC_BF16) | (p.a_mn ? (1u << 15) : 0u) |
                             (p.b_mn ? (1u << 16) : 0u);
      int it = 0, j = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++j) {
        const int z = item / tiles;
        const int kb0 = z * p.kb_per_split, kb1 = min(p.kblocks, kb0 + p.kb_per_split);
        const int iters = p.npairs * (kb1 - kb0);
        const int buf = j & 1;
        mbar_wait(bars + 8 * (B_TEMPTY + buf), ((j >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dacc = tmem + (uint32_t)(buf * GN);
        for (int li = 0; li < iters; ++li, ++it) {
          const int s = it % NST;
          const uint32_t ph = (it / NST) & 1;
          mbar_wait(bars + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * G_STAGE_BYTES;
          const uint64_t da = p.a_mn ? make_desc_mn_sw128(sa) : make_desc_k_sw128(sa);
          const uint64_t db = p.b_mn ? make_desc_mn_sw128(sa + G_A_BYTES) : make_desc_k_sw128(sa + G_A_BYTES);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k)
            umma_f16(dacc, da + ka * k, db + kb_ * k, idesc, (li > 0 || k > 0) ? 1u : 0u);
          umma_commit(bars + 8 * (GSTAGES + s));   // frees the smem stage when these MMAs retire
        }
        umma_commit(bars + 8 * (B_TFULL + buf));   // accumulator complete
      }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int quarter = warp & 3;
    const bool final_out = p.splits == 1;
    // alpha and the inverse scales of fp16 operands (powers of two: exact), applied one after the other so that
    // no intermediate product leaves the fp32 range
    const float inv_a = p.amax_a ? ldexpf(1.0f, -f16_scale_exp(*p.amax_a)) : 1.0f;
    const float inv_b = p.amax_b ? ldexpf(1.0f, -f16_scale_exp(*p.amax_b)) : 1.0f;
    const bool vec = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    int j = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++j) {
      const int z = item / tiles, tile = item - z * tiles;
      const int m0 = (tile / tiles_n) * GM, n0 = (tile % tiles_n) * TN;
      const int buf = j & 1;
      const int row = m0 + quarter * 32 + lane;
      mbar_wait(bars + 8 * (B_TFULL + buf), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* crow = p.C + (size_t)z * p.split_stride + (size_t)row * p.ldc;
      const int nchunks = (TN + 31) / 32;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * GN + c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
              "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
              "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
              "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == nchunks - 1) {
          // the whole accumulator is in registers: hand the TMEM buffer back before the stores
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * (B_TEMPTY + buf)) : "memory");
        }
        const int nb = n0 + c * 32;
        if (row < p.M && nb < p.N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (final_out) {   // uniform branches: predicated-off bias loads would still cost ~10 issue slots per element
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = f[i] * inv_a * inv_b * p.alpha;
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (nb + i < p.N) f[i] += p.bias[nb + i];
            }
          }
          if (!(vec && nb + 32 <= p.N)) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) crow[nb + i] = f[i];
          } else {
            // stage the warp's 32 x 32 block: a thread holds one ROW, but a coalesced store wants 8 lanes per row
            float* my = ep_stage + (size_t)(quarter * 32 + lane) * G_EP_PITCH;
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(my + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          }
        }
        if (vec && nb + 32 <= p.N) {  // warp-uniform
          __syncwarp();
          const int rl = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rl, grow = m0 + quarter * 32 + rr;
            const float4 v4 = *reinterpret_cast<const float4*>(ep_stage + (size_t)(quarter * 32 + rr) * G_EP_PITCH + c4);
            if (grow < p.M)
              *reinterpret_cast<float4*>(p.C + (size_t)z * p.split_stride + (size_t)grow * p.ldc + nb + c4) = v4;
          }
          __syncwarp();
        }
        if (STATS && nb < p.N) {  // warp-uniform: fused BatchNorm statistics of the finished output
          float a[32], b[32];
          const bool row_ok = row < p.M;
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = __uint_as_float(v[i]) * inv_a * inv_b * p.alpha;
          if (p.bias) {   // uniform branch (see above)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (nb + i < p.N) a[i] += p.bias[nb + i];
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            a[i] = (row_ok && nb + i < p.N) ? a[i] : 0.f;
            b[i] = a[i] * a[i];
          }
          const float cs = warp_colsum32(a, lane), cq = warp_colsum32(b, lane);
          sstat[quarter][c * 32 + lane] = make_float2(cs, cq);
        }
      }
      if (STATS) {
        // combine the four row quarters of the tile, then one fp64 atomic per column and quantity
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int e = threadIdx.x - 64;  // 0..127 over the epilogue warps
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int col = e + 128 * cc;
          if (n0 + col < p.N) {
            const float2 q0 = sstat[0][col], q1 = sstat[1][col], q2 = sstat[2][col], q3 = sstat[3][col];
            atomicAdd(&p.stat_sum[n0 + col], (double)q0.x + (double)q1.x + (double)q2.x + (double)q3.x);
            atomicAdd(&p.stat_sumsq[n0 + col], (double)q0.y + (double)q1.y + (double)q2.y + (double)q3.y);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // sstat is rewritten by the next tile
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(G_TMEM_COLS) : "memory");
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long long split_stride, int M,
                                     int N, long long ldc_part, float* __restrict__ C, long long ldc, float alpha,
                                     const float* __restrict__ bias, const uint32_t* __restrict__ amax_a,
                                     const uint32_t* __restrict__ amax_b) {
  long long n = (long long)M * N;
  const float inv_a = amax_a ? ldexpf(1.0f, -f16_scale_exp(*amax_a)) : 1.0f;
  const float inv_b = amax_b ? ldexpf(1.0f, -f16_scale_exp(*amax_b)) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / N), c = (int)(i % N);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[z * split_stride + (long long)r * ldc_part + c];
    s = s * inv_a * inv_b * alpha;
    if (bias) s += bias[c];
    C[(long long)r * ldc + c] = s;
  }
}

// ------------------------------------------------------------------ fp32 -> bf16 parts
// parts[i][r][c] (row stride ldp, zero padded to ldp) = i-th bf16 term of X[r][c].
__global__ void split_rows_kernel(const float* __restrict__ X, long long ldx, int M, int K, int nparts, float prescale,
                                  __nv_bfloat16* __restrict__ P0, __nv_bfloat16* __restrict__ P1,
                                  __nv_bfloat16* __restrict__ P2, long long ldp) {
  // one thread per 8 consecutive output elements (ldp % 8 == 0): 16-byte stores
  const long long segs_per_row = ldp / 8;
  const long long n = (long long)M * segs_per_row;
  const bool vec_in = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / segs_per_row;
    const int c = (int)(i - r * segs_per_row) * 8;
    float x[8];
    const float* src = X + r * ldx + c;
    if (vec_in && c + 8 <= K) {
      float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = c + j < K ? src[j] : 0.f;
    }
    __align__(16) __nv_bfloat16 h0[8], h1[8], h2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = x[j] * prescale;
      h0[j] = __float2bfloat16_rn(v);
      float r1 = v - __bfloat162float(h0[j]);
      h1[j] = __float2bfloat16_rn(r1);
      h2[j] = __float2bfloat16_rn(r1 - __bfloat162float(h1[j]));
    }
    const long long o = r * ldp + c;
    *reinterpret_cast<uint4*>(P0 + o) = *reinterpret_cast<const uint4*>(h0);
    if (nparts > 1) *reinterpret_cast<uint4*>(P1 + o) = *reinterpret_cast<const uint4*>(h1);
    if (nparts > 2) *reinterpret_cast<uint4*>(P2 + o) = *reinterpret_cast<const uint4*>(h2);
  }
}

// Transposing split: X is (R rows, C cols) fp32 row-major; parts are (C, ldp >= R) bf16 row-major,
// i.e. parts[i][c][r] = term_i(X[r][c]).  With T > 0 the rows are (b, t) pairs and `shift` delays
// time: output column (b, t) takes X[b, t - shift, :], zero for t < shift (S_prev for dV).
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ X, int R, int C, int nparts, int T, int shift, float prescale,
                       __nv_bfloat16* __restrict__ P0, __nv_bfloat16* __restrict__ P1,
                       __nv_bfloat16* __restrict__ P2, long long ldp) {
  // tile: 64 input rows (r) x 32 input columns (c).  Reads are 128-byte rows; every output row c
  // receives 64 consecutive r = 128 bytes, written as eight 16-byte stores.
  __shared__ float tile[64][33];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 64; j += 8) {
    int r = r0 + j, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      long long src = r;
      bool ok = true;
      if (T > 0 && shift > 0) {
        int t = r % T;
        ok = t >= shift;
        src = r - shift;
      }
      if (ok) v = X[src * (long long)C + c] * prescale;
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  const int cl = threadIdx.x >> 3, seg = threadIdx.x & 7;  // 32 output rows x 8 segments of 8 elements
  const int c = c0 + cl, rb = r0 + seg * 8;
  if (c < C && rb < ldp) {
    __align__(16) __nv_bfloat16 h0[8], h1[8], h2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float x = tile[seg * 8 + i][cl];  // zero beyond R (loaded as 0)
      h0[i] = __float2bfloat16_rn(x);
      float r1 = x - __bfloat162float(h0[i]);
      h1[i] = __float2bfloat16_rn(r1);
      h2[i] = __float2bfloat16_rn(r1 - __bfloat162float(h1[i]));
    }
    const long long o = (long long)c * ldp + rb;  // ldp % 8 == 0 and rb % 8 == 0: 16-byte aligned
    if (rb + 8 <= ldp) {
      *reinterpret_cast<uint4*>(P0 + o) = *reinterpret_cast<const uint4*>(h0);
      if (nparts > 1) *reinterpret_cast<uint4*>(P1 + o) = *reinterpret_cast<const uint4*>(h1);
      if (nparts > 2) *reinterpret_cast<uint4*>(P2 + o) = *reinterpret_cast<const uint4*>(h2);
    } else {
      for (int i = 0; i < 8 && rb + i < ldp; ++i) {
        P0[o + i] = h0[i];
        if (nparts > 1) P1[o + i] = h1[i];
        if (nparts > 2) P2[o + i] = h2[i];
      }
    }
  }
}


// ------------------------------------------------------------------ fp32 -> scaled fp16 hi/lo terms
// max|X| as a float bit pattern (non-negative floats order like unsigned integers); *amax zeroed by the caller.
__global__ void absmax_kernel(const float* __restrict__ X, long long ldx, long long M, int K, uint32_t* __restrict__ amax) {
  const bool vec = ((ldx & 3) == 0) && ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  float m = 0.f;
  if (vec) {
    const long long k4 = K / 4, n = M * k4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / k4;
      const float4 v = *reinterpret_cast<const float4*>(X + r * ldx + (i - r * k4) * 4);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
  } else {
    const long long n = M * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const long long r = i / K;
      m = fmaxf(m, fabsf(X[r * ldx + (i - r * K)]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // one atomic per BLOCK (a thousand blocks x 8 warps on one word serialised: 9.6 us for a 4 MB tensor)
  __shared__ float wm[32];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? wm[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
  }
}

// parts[i][r][c] = i-th fp16 term of prescale * 2^k * X[r][c], k = f16_scale_exp(*amax) (0 without amax).
__global__ void split_rows_f16_kernel(const float* __restrict__ X, long long ldx, int M, int K, int nparts, float prescale,
                                      const uint32_t* __restrict__ amax, __half* __restrict__ P0, __half* __restrict__ P1,
                                      long long ldp) {
  const long long segs_per_row = ldp / 8;
  const long long n = (long long)M * segs_per_row;
  const bool vec_in = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const float sc = amax ? ldexpf(1.0f, f16_scale_exp(*amax)) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / segs_per_row;
    const int c = (int)(i - r * segs_per_row) * 8;
    float x[8];
    const float* src = X + r * ldx + c;
    if (vec_in && c + 8 <= K) {
      float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = c + j < K ? src[j] : 0.f;
    }
    __align__(16) __half h0[8], h1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = x[j] * prescale * sc;   // powers of two / exact spike prescale: no rounding
      h0[j] = __float2half_rn(v);
      h1[j] = __float2half_rn(v - __half2float(h0[j]));
    }
    const long long o = r * ldp + c;
    *reinterpret_cast<uint4*>(P0 + o) = *reinterpret_cast<const uint4*>(h0);
    if (nparts > 1) *reinterpret_cast<uint4*>(P1 + o) = *reinterpret_cast<const uint4*>(h1);
  }
}

// ---------------------------------------------------------------- small fp32 products of the recurrent layers
// C[M][N] (=, +=) sum_k A(m, k) B[k][n] in plain fp32 FFMA: the two Be-sized products of a recurrent layer
// (rec_0 = s0 @ V0 at t = 0, snns.py:702 / 720, and the t = 0 frames of dV), each ~0.5 GFLOP -- too small to repay
// splitting their operands into 16-bit terms for the tensor pipe, and exact in fp32 whatever the precision mode.
// A_KM: A is stored (K, M) row-major (A(m, k) = A[k * lda + m]), else (M, K).  64 x 64 tile, K step 16, 4 x 4 per thread.
constexpr int SG_T = 64, SG_K = 16;

// four consecutive elements of a row: one 16-byte load when aligned and inside the matrix, guarded scalars otherwise
__device__ __forceinline__ float4 sg_ld4(const float* __restrict__ P, int64_t ld, int r, int c, int R, int Cn, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= R) return v;
  const float* q = P + r * ld + c;
  if (vec && c + 3 < Cn) return *reinterpret_cast<const float4*>(q);
  if (c < Cn) v.x = q[0];
  if (c + 1 < Cn) v.y = q[1];
  if (c + 2 < Cn) v.z = q[2];
  if (c + 3 < Cn) v.w = q[3];
  return v;
}

template <bool A_KM>
__global__ void __launch_bounds__(256)
small_gemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, float* __restrict__ C,
                  int64_t ldc, int M, int N, int K, int flags, int kper) {
  __shared__ __align__(16) float As[SG_K][SG_T + 4], Bs[SG_K][SG_T + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  const bool b_nk = flags & 8;                      // B stored (N, K) row-major: the product with B^T
  const bool va = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vb = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  // each thread moves one float4 of A and one of B per K step: (row = tid / 4, 4 columns at 4 (tid % 4)) of a
  // [64][16] block for a K-contiguous operand, (row = tid / 16, 4 columns at 4 (tid % 16)) of a [16][64] block otherwise
  auto fetch_a = [&](int k0) {
    return A_KM ? sg_ld4(A, lda, k0 + (tid >> 4), m0 + 4 * (tid & 15), kend, M, va)
                : sg_ld4(A, lda, m0 + (tid >> 2), k0 + 4 * (tid & 3), M, kend, va);
  };
  auto fetch_b = [&](int k0) {
    return b_nk ? sg_ld4(B, ldb, n0 + (tid >> 2), k0 + 4 * (tid & 3), N, kend, vb)
                : sg_ld4(B, ldb, k0 + (tid >> 4), n0 + 4 * (tid & 15), kend, N, vb);
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 ra = fetch_a(kbeg), rb = fetch_b(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += SG_K) {
    if (A_KM) {
      *reinterpret_cast<float4*>(&As[tid >> 4][4 * (tid & 15)]) = ra;
    } else {
      const int mm = tid >> 2, kk = 4 * (tid & 3);
      As[kk][mm] = ra.x; As[kk + 1][mm] = ra.y; As[kk + 2][mm] = ra.z; As[kk + 3][mm] = ra.w;
    }
    if (b_nk) {
      const int nn = tid >> 2, kk = 4 * (tid & 3);
      float bv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[kk + e][nn] = ((flags & 1) && k0 + kk + e == n0 + nn) ? 0.f : bv[e];
    } else {
      const int kk = tid >> 4, nn = 4 * (tid & 15);
      if (flags & 1) {                              // B with a zero diagonal (V0 of snns.py:712)
        const int dn = k0 + kk - (n0 + nn);
        if (dn == 0) rb.x = 0.f;
        if (dn == 1) rb.y = 0.f;
        if (dn == 2) rb.z = 0.f;
        if (dn == 3) rb.w = 0.f;
      }
      *reinterpret_cast<float4*>(&Bs[kk][nn]) = rb;
    }
    __syncthreads();
    if (k0 + SG_K < kend) {                         // next block's global loads fly under this block's FFMAs
      ra = fetch_a(k0 + SG_K);
      rb = fetch_b(k0 + SG_K);
    }
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (gridDim.z > 1) {                          // two K halves into a zeroed C: x + y is the same in either order
        atomicAdd(&C[m * ldc + n], v);
        continue;
      }
      if (flags & 4) v += C[m * ldc + n];
      if ((flags & 2) && m == n) v = 0.f;
      C[m * ldc + n] = v;
    }
  }
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_split_bf16(const float* X, int64_t ldx, int M, int K, int nparts, float prescale, void* P0, void* P1,
                      void* P2, int64_t ldp, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && K > 0 && nparts >= 1 && nparts <= 3 && ldp >= K && ldx >= K, "bad shape");
  SPARCH_REQUIRE(P0 && (nparts < 2 || P1) && (nparts < 3 || P2), "null part pointer");
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(X, "null pointer");
  SPARCH_REQUIRE((ldp % 8) == 0, "ldp must be a multiple of 8");
  int64_t n = (int64_t)M * (ldp / 8);
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  split_rows_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, as_stream(st)>>>(
      X, ldx, M, K, nparts, prescale, (__nv_bfloat16*)P0, (__nv_bfloat16*)P1, (__nv_bfloat16*)P2, ldp);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_split_bf16_transpose(const float* X, int R, int C, int nparts, int T, int shift, float prescale,
                                void* P0, void* P1, void* P2, int64_t ldp, sparch_stream_t st) {
  SPARCH_REQUIRE(R >= 0 && C > 0 && nparts >= 1 && nparts <= 3 && ldp >= R && shift >= 0, "bad shape");
  SPARCH_REQUIRE(P0 && (nparts < 2 || P1) && (nparts < 3 || P2), "null part pointer");
  if (ldp == 0) return SPARCH_OK;
  SPARCH_REQUIRE(X || R == 0, "null pointer");
  SPARCH_REQUIRE((ldp % 8) == 0, "ldp must be a multiple of 8");
  dim3 grid((C + 31) / 32, (unsigned)((ldp + 63) / 64)), block(256);
  split_transpose_kernel<<<grid, block, 0, as_stream(st)>>>(X, R, C, nparts, T, shift, prescale, (__nv_bfloat16*)P0,
                                                            (__nv_bfloat16*)P1, (__nv_bfloat16*)P2, ldp);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

// Split-K factor allowed for an (M, N) output: 8 sets of fp32 partials, up to 32 while they stay under 16 MB (the weight
// gradients of narrow layers -- dW of the 40-input first layer, of the 35-class readout -- contract over all Be*T frames
// into 8 output tiles: with 8 splits they ran on 64 CTAs and took longer than the full-width GEMMs next to them).
static int gemm_max_splits(int M, int N) {
  const size_t per = (size_t)M * (((size_t)N + 3) / 4 * 4) * sizeof(float);
  int cap = 8;
  while (cap < 32 && (size_t)(2 * cap) * per <= ((size_t)16 << 20)) cap *= 2;
  return cap;
}

size_t sparch_gemm_workspace(int M, int N, int K) {
  return (size_t)gemm_max_splits(M, N) * M * (((size_t)N + 3) / 4 * 4) * sizeof(float);
}

int sparch_absmax(const float* X, int64_t ldx, int64_t M, int K, uint32_t* amax, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && K > 0 && ldx >= K && amax, "bad argument");
  SPARCH_CUDA(cudaMemsetAsync(amax, 0, sizeof(uint32_t), as_stream(st)));
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(X, "null pointer");
  int64_t n = M * (int64_t)K / 4, g = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  if (g < 1) g = 1;
  absmax_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, as_stream(st)>>>(X, ldx, M, K, amax);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_split_f16(const float* X, int64_t ldx, int M, int K, int nparts, float prescale, uint32_t* amax,
                     int compute_amax, void* P0, void* P1, int64_t ldp, sparch_stream_t st) {
  SPARCH_REQUIRE(M >= 0 && K > 0 && nparts >= 1 && nparts <= 2 && ldp >= K && ldx >= K, "bad shape");
  SPARCH_REQUIRE(P0 && (nparts < 2 || P1), "null part pointer");
  SPARCH_REQUIRE(!compute_amax || amax, "compute_amax needs the amax word");
  if (compute_amax)
    if (int e = sparch_absmax(X, ldx, M, K, amax, st)) return e;
  if (M == 0) return SPARCH_OK;
  SPARCH_REQUIRE(X, "null pointer");
  SPARCH_REQUIRE((ldp % 8) == 0, "ldp must be a multiple of 8");
  int64_t n = (int64_t)M * (ldp / 8);
  int64_t g = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
  split_rows_f16_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, as_stream(st)>>>(X, ldx, M, K, nparts, prescale, amax,
                                                                                  (__half*)P0, (__half*)P1, ldp);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_gemm_bf16(const void* const* A_parts, int na, const void* const* B_parts, int nb, int64_t lda,
                     int64_t ldb, int a_mn, int b_mn, int a_koff, const int* pair_a, const int* pair_b, int npairs,
                     int M, int N, int K, float alpha, const float* bias, float* C, int64_t ldc, double* stat_sum,
                     double* stat_sumsq, void* workspace, sparch_stream_t st_) {
  return sparch_gemm_terms(0, A_parts, na, nullptr, B_parts, nb, nullptr, lda, ldb, a_mn, b_mn, a_koff, pair_a, pair_b,
                           npairs, M, N, K, alpha, bias, C, ldc, stat_sum, stat_sumsq, workspace, st_);
}

int sparch_gemm_terms(int fp16, const void* const* A_parts, int na, const uint32_t* amax_a, const void* const* B_parts,
                      int nb, const uint32_t* amax_b, int64_t lda, int64_t ldb, int a_mn, int b_mn, int a_koff,
                      const int* pair_a, const int* pair_b, int npairs, int M, int N, int K, float alpha,
                      const float* bias, float* C, int64_t ldc, double* stat_sum, double* stat_sumsq, void* workspace,
                      sparch_stream_t st_) {
  SPARCH_REQUIRE(fp16 || (!amax_a && !amax_b), "scaled operands are fp16 terms");
  SPARCH_REQUIRE(M > 0 && N > 0 && K > 0 && na >= 1 && na <= 3 && nb >= 1 && nb <= 3, "bad shape");
  SPARCH_REQUIRE(npairs >= 1 && npairs <= 8 && A_parts && B_parts && pair_a && pair_b && C, "bad argument");
  SPARCH_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K),
                 "operand row strides must be multiples of 8 bf16 elements (16 bytes) and cover a row");
  SPARCH_REQUIRE(a_koff == 0 || a_mn, "a_koff applies to an MN-major A operand");
  SPARCH_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "stat_sum and stat_sumsq go together");
  cudaStream_t st = as_stream(st_);
  const int tile_n = N >= GN ? GN : (N + 15) / 16 * 16;
  TmapSet maps;
  memset(&maps, 0, sizeof maps);
  for (int i = 0; i < na; ++i) {
    SPARCH_REQUIRE(A_parts[i] && (reinterpret_cast<uintptr_t>(A_parts[i]) & 15) == 0, "A part null or unaligned");
    if (int e = a_mn ? make_map_mn(&maps.a[i], A_parts[i], K, M, lda) : make_map(&maps.a[i], A_parts[i], M, K, lda, GM))
      return e;
  }
  for (int i = 0; i < nb; ++i) {
    SPARCH_REQUIRE(B_parts[i] && (reinterpret_cast<uintptr_t>(B_parts[i]) & 15) == 0, "B part null or unaligned");
    if (int e = b_mn ? make_map_mn(&maps.b[i], B_parts[i], K, N, ldb) : make_map(&maps.b[i], B_parts[i], N, K, ldb, tile_n))
      return e;
  }
  GemmParams p;
  memset(&p, 0, sizeof p);
  p.M = M; p.N = N; p.K = K; p.npairs = npairs;
  p.a_mn = a_mn; p.b_mn = b_mn; p.a_koff = a_koff;
  p.fp16 = fp16 ? 1 : 0; p.amax_a = amax_a; p.amax_b = amax_b;
  p.stat_sum = stat_sum; p.stat_sumsq = stat_sumsq;
  if (stat_sum) {
    SPARCH_CUDA(cudaMemsetAsync(stat_sum, 0, sizeof(double) * N, st));
    SPARCH_CUDA(cudaMemsetAsync(stat_sumsq, 0, sizeof(double) * N, st));
  }
  for (int i = 0; i < npairs; ++i) {
    SPARCH_REQUIRE(pair_a[i] >= 0 && pair_a[i] < na && pair_b[i] >= 0 && pair_b[i] < nb, "pair index out of range");
    p.pair_a[i] = pair_a[i];
    p.pair_b[i] = pair_b[i];
  }
  p.kblocks = (K + GK - 1) / GK;
  const int tiles = ((M + GM - 1) / GM) * ((N + tile_n - 1) / tile_n);
  int splits = 1;
  if (workspace && tiles < sm_count() && !stat_sum) {  // statistics need the finished tile in one epilogue
    splits = sm_count() / tiles;
    if (splits > gemm_max_splits(M, N)) splits = gemm_max_splits(M, N);
    if (splits > p.kblocks) splits = p.kblocks;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (p.kblocks + splits - 1) / splits;
  splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
  const long long ldw = ((long long)N + 3) / 4 * 4;
  if (splits > 1) {
    p.C = reinterpret_cast<float*>(workspace);
    p.ldc = ldw;
    p.split_stride = (long long)M * ldw;
    p.alpha = 1.f;
    p.bias = nullptr;
  } else {
    p.C = C; p.ldc = ldc; p.split_stride = 0; p.alpha = alpha; p.bias = bias;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    SPARCH_CUDA(cudaFuncSetAttribute(gemm_tn_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM));
    SPARCH_CUDA(cudaFuncSetAttribute(gemm_tn_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM));
  }
  // (A 2-stage ring with two CTAs per SM for short contractions was measured: no gain.)
  p.splits = splits;
  p.tile_n = tile_n;
  const int items = tiles * splits;
  dim3 grid(items < sm_count() ? items : sm_count());
  if (stat_sum)
    gemm_tn_bf16_kernel<true><<<grid, G_THREADS, G_SMEM, st>>>(maps, p);
  else
    gemm_tn_bf16_kernel<false><<<grid, G_THREADS, G_SMEM, st>>>(maps, p);
  SPARCH_LAUNCH_OK();
  if (splits > 1) {
    long long n = (long long)M * N;
    long long g = (n + 255) / 256, cap = (long long)sm_count() * 8;
    splitk_reduce_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(reinterpret_cast<float*>(workspace), splits,
                                                                        p.split_stride, M, N, ldw, C, ldc, alpha, bias, amax_a,
                                                                        amax_b);
    SPARCH_LAUNCH_OK();
  }
  return SPARCH_OK;
}

int sparch_small_gemm(const float* A, int64_t lda, int a_km, const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                      int N, int K, int flags, sparch_stream_t st) {
  SPARCH_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "bad argument");
  SPARCH_REQUIRE(lda >= (a_km ? M : K) && ldb >= ((flags & 8) ? K : N) && ldc >= N, "leading dimension smaller than the row");
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T, 1);
  int kper = K;
  // few tiles and a long contraction (rec_0: 64 tiles, K = H): two K halves, added into a zeroed C by atomics
  if (!(flags & (2 | 4)) && (int)(grid.x * grid.y) < sm_count() && K >= 8 * SG_K && ldc == N) {
    grid.z = 2;
    kper = ((K + 2 * SG_K - 1) / (2 * SG_K)) * SG_K;
    SPARCH_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), as_stream(st)));
  }
  if (a_km) small_gemm_kernel<true><<<grid, 256, 0, as_stream(st)>>>(A, lda, B, ldb, C, ldc, M, N, K, flags, kper);
  else small_gemm_kernel<false><<<grid, 256, 0, as_stream(st)>>>(A, lda, B, ldb, C, ldc, M, N, K, flags, kper);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

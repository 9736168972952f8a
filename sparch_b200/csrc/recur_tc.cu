// Reverse-time recurrence of the recurrent kinds with dI_{t+1} @ V0^T on tcgen05 (alternative to
// rec_bwd_persist_kernel in recur.cu; same BPTT update, same hand-over idea), split-K over a cluster.
//
//   Cluster of 4 CTAs = 64 batch rows x 128 presynaptic neurons j, persistent over all T steps.  CTA
//   rank r of the cluster multiplies over its QUARTER of K (neurons c in [r Hp/4, (r+1) Hp/4)) for all
//   128 columns, then the four 64 x 128 partial products are reduce-scattered through distributed
//   shared memory: rank r sums and owns columns 32 r .. 32 r + 31 (so CTA bx updates neurons 32 bx ..).
//   B operand: V0[128 j][Hp/4 c] as fp16 hi + lo, K-major SWIZZLE_128B tiles resident in shared memory
//              (Hp/256 x 2 x 16 KB = 128 KB at H = 1024).
//   A operand: dI_{t+1} of the 64 rows, handed over through L2 as two plain row-major fp16 matrices
//              (hi, lo); each CTA fetches only its K quarter by TMA (64 x 64 boxes, SWIZZLE_128B).
//   D: 64 x 128 fp32 in 128 TMEM columns; per K16 three UMMAs (hi*hi, hi*lo, lo*hi) with N = 128, issued
//      by one thread; the update warps read D with tcgen05.ld (M = 64 layout: rows 16q..16q+15 live in
//      TMEM lanes 32q..32q+15) and scatter the column quarters to their owners (st.shared::cluster +
//      remote mbarrier arrive).
//
// Why split K instead of N (measured on B200 with the first version of this kernel, one CTA = 64 rows x
// 32 neurons over the whole K range): a tcgen05.mma at M = 64 costs ~58 cycles whether N is 32 or 128,
// so the 192 UMMAs a whole-K CTA needs per step took 11.3 k cycles, more than the mma.sync kernel.  With
// N = 128 and a quarter of K a CTA issues 48 per step, and the panel traffic out of L2 drops from 256 KB
// to 64 KB per CTA per step.
//
// Scaling.  fp16 needs a scale, and the tensor core accumulates over the whole K range without
// intervention, so the scale must be per ROW (not per 32-column chunk as in the mma.sync kernel).  The
// row maximum of |dI_t| is only known once all slices have produced their part, so the scale of step t
// is derived one step late: M_t = max(rowmax|dI_{t+1}|, 0.25 * rowmax|g_t|), s_t = 2^(4 - exp(M_t)).
// |dI_t| can exceed M_t by the growth of one step (bounded by a few hundred), far inside fp16's range
// above 16; values below it only lose ABSOLUTE precision (fp16 hi + lo keeps 2^-25 of the scaled unit),
// i.e. <= 2^-29 of the row maximum -- below fp32 rounding of the dominant terms.  Every slice computes
// the same s_t from the same published chunk maxima, so producer and consumers agree.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "cell_math.cuh"
#include "common.cuh"
#include "tcgen05_utils.cuh"

namespace sparch {

constexpr int TC_ROWS = 64;            // batch rows per CTA
constexpr int TC_COLS = 32;            // neurons updated per CTA
constexpr int TC_CL = 4;               // CTAs per cluster = K quarters
constexpr int TC_N = TC_COLS * TC_CL;  // UMMA N: neurons per cluster
constexpr int TC_STAGES = 3;           // A ring depth (a fourth stage leaves too little L1: measured slower)
constexpr int TC_RS = 65;              // receive buffer: float4 row stride (64 rows + 1: conflict-free reads)
constexpr int TC_RECV_BYTES = TC_CL * 8 * TC_RS * 16;
constexpr int TC_STAGE_BYTES = 16384;  // 64 rows x 64 K x (hi, lo) fp16
constexpr int TC_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 update
constexpr int TC_VSCALE_EXP = 13;

struct TcMaps {
  CUtensorMap hi, lo;
};

struct RecBwdTcArgs {
  const float *G, *U, *W, *alpha, *beta, *a, *b, *u0, *w0, *s0;
  const uint32_t* img;   // [CTA][kb][part][128 rows x 128 B] swizzled fp16 tiles of V0 (vprep_umma_kernel)
  const int* meta;
  const float* gmax;     // [Be][T] row maxima of |G|
  float theta;
  float *dI, *p_alpha, *p_beta, *p_a, *p_b;
  __half *panel_hi, *panel_lo;  // [2][groups*64][Hp]
  float* cmax;                  // [2][groups][Hp/32][64] chunk maxima of |dI_t|
  int Be, T, H, Hp, KB;  // Hp: H padded to 256; KB = Hp / 256 k-blocks of 64 per CTA
  int reduced;
  long long* dbg;  // optional [T][8] phase clocks of CTA (0,0) (profiling aid), normally NULL
  int dbg_flags;   // profiling experiments (results invalid): 4 no TMA traffic, 16 no UMMA
};

// V0 as UMMA B tiles.  CTA bx = (cluster bx / 4, rank bx % 4) holds rows j = 128 (bx / 4) + n, n < 128, and
// K = (bx % 4) Hp / 4 + 64 kb + k; element (n, k) of a tile: halves offset n*64 + ((k/8) ^ (n&7))*8 + k%8.
__global__ void vprep_umma_kernel(const float* __restrict__ V, int H, int Hp, int KB, const int* __restrict__ meta,
                                  __half* __restrict__ img) {
  // one thread per 16-byte swizzle chunk (8 consecutive K of one row n): reads 32 contiguous bytes of V once and
  // writes the chunk of the hi tile and of the lo tile
  const int64_t per_cta8 = (int64_t)KB * (TC_N * 8);             // chunks per CTA and part
  const int64_t total8 = per_cta8 * (Hp / TC_COLS);
  const float sc = ldexpf(1.0f, TC_VSCALE_EXP - meta[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int bx = (int)(i / per_cta8);
    const int r = (int)(i - (int64_t)bx * per_cta8);
    const int chunk_sw = r & 7, n = (r >> 3) & (TC_N - 1), kb = r >> 10;
    const int k0 = (chunk_sw ^ (n & 7)) << 3;
    const int row = (bx / TC_CL) * TC_N + n;                      // V[row = presynaptic j][col = neuron c]
    const int col0 = (bx % TC_CL) * (Hp / TC_CL) + kb * 64 + k0;
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = col0 + e;
      float x = 0.f;
      if (row < H && col < H && row != col) x = V[(int64_t)row * H + col] * sc;
      hi[e] = __float2half_rn(x);
      lo[e] = __float2half_rn(x - __half2float(hi[e]));
    }
    // halves offset inside the CTA's image: ((kb * 2 + part) * TC_N + n) * 64 + chunk_sw * 8
    __half* base = img + ((int64_t)bx * per_cta8 * 2 + ((int64_t)(kb * 2) * TC_N + n) * 8 + chunk_sw) * 8;
    *reinterpret_cast<uint4*>(base) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + (int64_t)TC_N * 64) = *reinterpret_cast<const uint4*>(lo);
  }
}

__global__ void gmax_kernel(const float* __restrict__ G, int Be, int T, int H, float* __restrict__ gmax) {
  // one warp per (b, t) row of H values
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)Be * T) return;
  const float* row = G + w * H;
  float m = 0.f;
  for (int i = lane; i < H; i += 32) m = fmaxf(m, fabsf(row[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) gmax[w] = m;
}

__device__ __forceinline__ bool tc_elect_one() {
  uint32_t e;
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(e));
  return e != 0;
}

__device__ __forceinline__ void upd_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float scale_from_max(float m) {
  int e = 0;
  if (m > 0.f && m <= 3.0e38f) frexpf(m, &e);
  e = max(e, -100);
  return ldexpf(1.0f, 4 - e);
}

template <bool ADAPT>
__global__ void __launch_bounds__(TC_THREADS, 1)
rec_bwd_tc_kernel(const __grid_constant__ TcMaps maps, const RecBwdTcArgs p, const int group0, const int ngroups_total,
                  int* __restrict__ counters) {
  extern __shared__ unsigned char tsm_raw[];
  const uint32_t raw = smem_u32(tsm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* tsm = tsm_raw + (base - raw);
  const size_t v_bytes = (size_t)p.KB * 2 * (TC_N * 128);       // KB x (hi, lo) x 16 KB
  const uint32_t vimg = base;
  const uint32_t ring = base + (uint32_t)v_bytes;               // TC_STAGES * 16 KB
  const uint32_t recv = ring + TC_STAGES * TC_STAGE_BYTES;      // [4 source ranks][8 column groups][65] float4
  const uint32_t bars = recv + TC_RECV_BYTES;                   // full[4], empty[4], acc_full, acc_empty, recv_full
  unsigned char* tail = tsm + v_bytes + TC_STAGES * TC_STAGE_BYTES + TC_RECV_BYTES;
  const float4* recv_f4 = reinterpret_cast<const float4*>(tail - TC_RECV_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 128);
  __shared__ float sprm[6][TC_COLS];
  constexpr int B_ACC_FULL = 2 * TC_STAGES, B_ACC_EMPTY = 2 * TC_STAGES + 1, B_RECV = 2 * TC_STAGES + 2;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, group = group0 + blockIdx.y, row0 = group * TC_ROWS;
  const int rank = blockIdx.x % TC_CL;  // == %cluster_ctarank for cluster dims (4, 1, 1)
  const int nslices = gridDim.x, NCH = p.Hp / 32;
  // Hand-over flags: one arrival counter per (row group, 64-neuron K block); the two slices that produce a K block
  // add 1 each per step.  A consumer waits per K block of its own K quarter (and not for all 32 slices of the group):
  // less skew to wait out, and two instead of 32 CTAs contend for an address.
  const int NKB = p.Hp / 64;                           // K blocks per row = flags per group (<= 16 for H <= 1024)
  int* flags = counters + (size_t)group * NKB;
  int* my_flag = flags + slice / 2;
  const bool dbg_cta = p.dbg && blockIdx.x == 0 && blockIdx.y == 0;

  {  // resident V0 tiles
    const uint4* src = reinterpret_cast<const uint4*>(p.img) + (size_t)slice * (v_bytes / 16);
    uint4* dst = reinterpret_cast<uint4*>(tsm);
    for (int i = tid; i < (int)(v_bytes / 16); i += TC_THREADS) {
      uint32_t sa = smem_u32(dst + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i));
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
  }
  if (tid < TC_COLS) {
    const int col = min(slice * TC_COLS + tid, p.H - 1);
    const NeuronParams q0 = load_params<ADAPT>(p.alpha, p.beta, p.a, p.b, col);
    sprm[0][tid] = q0.alpha; sprm[1][tid] = q0.oma; sprm[2][tid] = q0.beta; sprm[3][tid] = q0.a;
    sprm[4][tid] = q0.b; sprm[5][tid] = 1.0f / q0.oma;
  }
  if (tid == 0) {
    for (int st = 0; st < TC_STAGES; ++st) {
      mbar_init(bars + 8 * st, 1);
      mbar_init(bars + 8 * (TC_STAGES + st), 1);
    }
    mbar_init(bars + 8 * B_ACC_FULL, 1);    // one tcgen05.commit
    mbar_init(bars + 8 * B_ACC_EMPTY, 8);   // one arrival per update warp
    mbar_init(bars + 8 * B_RECV, 1);  // armed per step with the 32 KB the four ranks deliver (st.async complete_tx)
    mbar_expect_tx(bars + 8 * B_RECV, TC_CL * TC_ROWS * TC_COLS * 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // make the cp.async-written tiles visible to the tensor core's (async proxy) reads
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // cluster-wide: nobody may arrive on a peer's recv barrier before it is initialised
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const float rs = ldexpf(1.0f, p.meta[0] - TC_VSCALE_EXP);

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the steps; lane 0 does the work) =====
    int it = 0;
    for (int t = p.T - 2; t >= 0; --t) {
      const int rbuf = (t + 1) & 1;
      const int target = 2 * (p.T - 1 - t);            // both producer slices of a K block have published panel(t+1)
      const int y = (rbuf * ngroups_total + group) * TC_ROWS;
      // lanes 0..KB-1: wait for K block `lane` of this rank's quarter, then fetch it (each lane drives its own ring
      // slot).  Rounds of TC_STAGES lanes: two fills of the same slot must reach its empty barrier one after the other,
      // a parity wait cannot tell "one phase behind" from "three behind".
      for (int base_kb = 0; base_kb < p.KB; base_kb += TC_STAGES) {
      if (lane >= base_kb && lane < min(base_kb + TC_STAGES, p.KB)) {
        const int kb = lane;
        const int* f = flags + rank * p.KB + kb;
        const long long t0 = clock64();
        while (true) {
          int v;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
          if (v >= target) break;
          if (clock64() - t0 > 4000000000LL) __trap();
        }
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of other SMs -> TMA reads
        if (dbg_cta && kb == 0) p.dbg[t * 8 + 0] = clock64();
        const int it_kb = it + kb;
        const int s = it_kb % TC_STAGES;
        const uint32_t ph = (it_kb / TC_STAGES) & 1;
        mbar_wait(bars + 8 * (TC_STAGES + s), ph ^ 1);
        if (p.dbg_flags & 4) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * s) : "memory");
        } else {
          mbar_expect_tx(bars + 8 * s, TC_STAGE_BYTES);
          const uint32_t sa = ring + s * TC_STAGE_BYTES;
          const int x = rank * (p.Hp / TC_CL) + kb * 64;  // this rank's K quarter
          tma_load_2d(sa, &maps.hi, x, y, bars + 8 * s);
          tma_load_2d(sa + 8192, &maps.lo, x, y, bars + 8 * s);
        }
      }
      __syncwarp();
      }
      it += p.KB;
      // the chunk maxima of ALL slices feed the row scale: wait for the rest of the group's flags (off the MMA's path)
      if (lane < NKB) {
        const int* f = flags + lane;
        const long long t0 = clock64();
        while (true) {
          int v;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
          if (v >= target) break;
          if (clock64() - t0 > 4000000000LL) __trap();
        }
      }
      __syncwarp();
      asm volatile("bar.arrive 2, 288;" ::: "memory");    // the update warps may read the chunk maxima now
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp stays converged, one elected lane issues each K block's UMMAs as straight-line code
    // (under a plain `if (lane == 0)` ptxas wraps every UMMA in an ELECT / PLOP3 / BRA.U.ANY retry loop, ~45 cycles
    // per instruction whatever its shape: measured with tools/ubench/umma_i8_ts.cu) =====
    {
      // kind::f16, fp16 x fp16 -> fp32, both K-major, M = 64, N = 128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      const uint64_t ring_desc = make_desc_k_sw128(ring), v_desc = make_desc_k_sw128(vimg);
      int it = 0, step = 0;
      for (int t = p.T - 2; t >= 0; --t, ++step) {
        mbar_wait_sleep(bars + 8 * B_ACC_EMPTY, (step & 1) ^ 1);  // D drained by the update warps
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait_sleep(bars + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (dbg_cta && lane == 0 && kb == 0) p.dbg[t * 8 + 1] = clock64();
          __syncwarp();
          if (tc_elect_one()) {
            const uint64_t a_hi = ring_desc + (uint64_t)(s * (TC_STAGE_BYTES >> 4)), a_lo = a_hi + (8192 >> 4);
            const uint64_t b_hi = v_desc + (uint64_t)((kb * 2) * (TC_N * 128 >> 4)), b_lo = b_hi + (TC_N * 128 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (p.dbg_flags & 16) continue;
              umma_f16(tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              if (p.reduced) continue;  // reduced-precision mode: hi * hi only
              umma_f16(tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
              umma_f16(tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
            }
            umma_commit(bars + 8 * (TC_STAGES + s));
            if (kb == p.KB - 1) umma_commit(bars + 8 * B_ACC_FULL);
          }
          __syncwarp();
        }
        if (dbg_cta && lane == 0) p.dbg[t * 8 + 2] = clock64();
      }
    }
  } else {
    // ===== update warps: BPTT for the 64 x 32 block, 8 neurons of one row per thread =====
    const int uw = warp - 2, q = warp & 3, half = uw >> 2;  // TMEM lane quarter of this warp, column half
    // (q, half) only address TMEM; the update itself maps 4 consecutive lanes to the 4 column quarters of a row
    const int r = 8 * uw + (lane >> 2), cq = lane & 3;
    const int row = row0 + r;
    const int col0 = slice * TC_COLS + cq * 8;
    const bool live = row < p.Be && col0 < p.H;
    const bool vec = ((p.H & 3) == 0) && (col0 + 8 <= p.H);
    const int nv = live ? min(8, p.H - col0) : 0;
    const int64_t idx0 = (int64_t)row * p.H + col0;
    float du[8], dw[8], pa[8], pb[8], pc[8], pd[8], ut[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) du[i] = dw[i] = pa[i] = pb[i] = pc[i] = pd[i] = ut[i] = 0.f;
    if (live && p.T > 0) {
      const float* src = p.U + ((int64_t)row * p.T + (p.T - 1)) * p.H + col0;
#pragma unroll
      for (int i = 0; i < 8; ++i) ut[i] = i < nv ? src[i] : 0.f;
    }
    float inv_s_next = 1.0f;  // 1 / s_{t+1}: unscales D at step t
    int step = 0;
    for (int t = p.T - 1; t >= 0; --t) {
      const int64_t o0 = ((int64_t)row * p.T + t) * p.H + col0;
      const int rbuf = (t + 1) & 1, wbuf = t & 1;
      float gq[8], up[8], wp[8], sp[8], recb[8], d[8];
      const bool dbg_on = dbg_cta && tid == 64;
      if (dbg_on) p.dbg[t * 8 + 5] = clock64();
#pragma unroll
      for (int i = 0; i < 8; ++i) gq[i] = up[i] = wp[i] = sp[i] = recb[i] = d[i] = 0.f;
      if (live) {
        const float* gp = p.G + o0;
        const float* upp = t > 0 ? p.U + o0 - p.H : p.u0 + idx0;
        const float* wpp = ADAPT ? (t > 0 ? p.W + o0 - p.H : p.w0 + idx0) : nullptr;
        if (vec) {
          const float4 a0 = *reinterpret_cast<const float4*>(gp), a1 = *reinterpret_cast<const float4*>(gp + 4);
          gq[0] = a0.x; gq[1] = a0.y; gq[2] = a0.z; gq[3] = a0.w; gq[4] = a1.x; gq[5] = a1.y; gq[6] = a1.z; gq[7] = a1.w;
          const float4 b0 = *reinterpret_cast<const float4*>(upp), b1 = *reinterpret_cast<const float4*>(upp + 4);
          up[0] = b0.x; up[1] = b0.y; up[2] = b0.z; up[3] = b0.w; up[4] = b1.x; up[5] = b1.y; up[6] = b1.z; up[7] = b1.w;
          if (ADAPT) {
            const float4 c0 = *reinterpret_cast<const float4*>(wpp), c1 = *reinterpret_cast<const float4*>(wpp + 4);
            wp[0] = c0.x; wp[1] = c0.y; wp[2] = c0.z; wp[3] = c0.w; wp[4] = c1.x; wp[5] = c1.y; wp[6] = c1.z; wp[7] = c1.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nv) {
              gq[i] = gp[i];
              up[i] = upp[i];
              if (ADAPT) wp[i] = wpp[i];
            }
        }
        if (t == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nv) sp[i] = p.s0[idx0 + i];
        }
      }
      // ---- row scale of this step's hand-over: M_t = max(rowmax |dI_{t+1}|, 0.25 rowmax |g_t|)
      float m_next = 0.f;
      if (t < p.T - 1) {
        asm volatile("bar.sync 2, 288;" ::: "memory");  // producer saw the group's counter: chunk maxima are visible
        const float* cm = p.cmax + ((size_t)rbuf * ngroups_total + group) * NCH * TC_ROWS;
        for (int c = cq; c < NCH; c += 4) m_next = fmaxf(m_next, __ldcg(&cm[c * TC_ROWS + r]));
      }
      m_next = fmaxf(m_next, __shfl_xor_sync(0xffffffffu, m_next, 1));
      m_next = fmaxf(m_next, __shfl_xor_sync(0xffffffffu, m_next, 2));
      const float g_row = row < p.Be ? p.gmax[(size_t)row * p.T + t] : 0.f;
      const float s_t = scale_from_max(fmaxf(m_next, 0.25f * g_row));
      if (t < p.T - 1) {
        // ---- partial D = dI_{t+1}[:, K quarter] (scaled) @ V0^T (scaled): this warp holds rows 16q..16q+15
        // (lanes 0..15) x columns 64 half .. 64 half + 63 = the column quarters of ranks 2 half, 2 half + 1
        mbar_wait_sleep(bars + 8 * B_ACC_FULL, step & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_on) p.dbg[t * 8 + 3] = clock64();
#pragma unroll
        for (int dq = 0; dq < 2; ++dq) {
          const int dest = 2 * half + dq;
          uint32_t v[32];
          const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * dest);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (lane < 16) {
            // row 16q + lane, 32 columns -> slot [this rank] of the owner's receive buffer, laid out
            // [rank][4-column group j][row] so that one store instruction covers contiguous bytes.
            // (Spreading the stores over all 32 lanes by shuffles was measured: no faster, more registers.)
            const uint32_t la = recv + (uint32_t)((rank * 8 * TC_RS) + 16 * q + lane) * 16;
            uint32_t ra, rb;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(dest));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(bars + 8 * B_RECV), "r"(dest));
#pragma unroll
            for (int j = 0; j < 8; ++j)  // each store reports its 16 bytes to the owner's barrier: no fence, no arrival
              asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                               ra + 16 * TC_RS * j),
                           "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]), "r"(rb)
                           : "memory");
          }
        }
        if (dbg_on) p.dbg[(p.T + t) * 8 + 4] = clock64();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * B_ACC_EMPTY) : "memory");
        if (dbg_on) p.dbg[(p.T + t) * 8 + 5] = clock64();
        // ---- this CTA's column quarter from the four ranks, summed in rank order
        {
          uint32_t done = 0;
          const long long t0 = clock64();
          while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(bars + 8 * B_RECV), "r"((uint32_t)(step & 1))
                : "memory");
            if (!done && clock64() - t0 > 4000000000LL) __trap();
          }
        }
        if (tid == 64 && t > 0) mbar_expect_tx(bars + 8 * B_RECV, TC_CL * TC_ROWS * TC_COLS * 4);  // arm the next phase
        if (dbg_on) p.dbg[t * 8 + 6] = clock64();
        const float k = inv_s_next * rs;
        float acc[8];
#pragma unroll
        for (int src = 0; src < TC_CL; ++src) {
          const float4* rp = recv_f4 + (src * 8 + 2 * cq) * TC_RS + r;
          const float4 x0 = rp[0], x1 = rp[TC_RS];
          const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) acc[jj] = src == 0 ? xs[jj] : acc[jj] + xs[jj];
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) recb[jj] = acc[jj] * k;
        ++step;
      }
      if (live) {
        if (t > 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) sp[i] = spike_of(__fsub_rn(up[i], p.theta));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nv) {
            const int lc = cq * 8 + i;
            NeuronParams npi;
            npi.alpha = sprm[0][lc]; npi.oma = sprm[1][lc]; npi.beta = sprm[2][lc]; npi.a = sprm[3][lc];
            npi.b = sprm[4][lc];
            float dwi = dw[i], pbi = pb[i], pci = pc[i], pdi = pd[i];
            d[i] = step_bwd<ADAPT>(npi, sprm[5][lc], p.theta, gq[i], recb[i], ut[i], up[i], sp[i], wp[i], du[i], dwi,
                                   pa[i], pbi, pci, pdi);
            dw[i] = dwi; pb[i] = pbi; pc[i] = pci; pd[i] = pdi;
            ut[i] = up[i];
          }
        }
      }
      if (dbg_on) p.dbg[(p.T + t) * 8 + 0] = clock64();
      if (t > 0) {
        // ---- hand dI_t over: chunk maximum (for the next scale) and fp16 hi/lo rows scaled by s_t
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(d[i]));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        if (dbg_on) p.dbg[(p.T + t) * 8 + 1] = clock64();
        if (cq == 0) {
          p.cmax[(((size_t)wbuf * ngroups_total + group) * NCH + slice) * TC_ROWS + r] = m;
        }
        __align__(16) __half hh[8], hl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = d[i] * s_t;
          hh[i] = __float2half_rn(x);
          hl[i] = __float2half_rn(x - __half2float(hh[i]));
        }
        const size_t po = ((size_t)(wbuf * ngroups_total + group) * TC_ROWS + r) * p.Hp + slice * TC_COLS + cq * 8;
        *reinterpret_cast<uint4*>(p.panel_hi + po) = *reinterpret_cast<const uint4*>(hh);
        *reinterpret_cast<uint4*>(p.panel_lo + po) = *reinterpret_cast<const uint4*>(hl);
        if (dbg_on) p.dbg[(p.T + t) * 8 + 2] = clock64();
        upd_sync();                       // every update thread's panel stores are issued
        if (dbg_on) p.dbg[(p.T + t) * 8 + 3] = clock64();
        if (tid == 64) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(my_flag) : "memory");
        if (dbg_on) p.dbg[t * 8 + 4] = clock64();
      }
      if (live) {
        float* dp = p.dI + o0;
        if (vec) {
          *reinterpret_cast<float4*>(dp) = make_float4(d[0], d[1], d[2], d[3]);
          *reinterpret_cast<float4*>(dp + 4) = make_float4(d[4], d[5], d[6], d[7]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < nv) dp[i] = d[i];
        }
      }
      inv_s_next = 1.0f / s_t;
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nv) {
          p.p_alpha[idx0 + i] = pa[i];
          if (ADAPT) {
            p.p_beta[idx0 + i] = pb[i];
            p.p_a[idx0 + i] = pc[i];
            p.p_b[idx0 + i] = pd[i];
          }
        }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // cluster-wide: no CTA leaves while a peer may still store into its receive buffer
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

static size_t rec_bwd_tc_smem(int KB) {
  return (size_t)KB * 2 * (TC_N * 128) + TC_STAGES * TC_STAGE_BYTES + TC_RECV_BYTES + 256 + 1024;
}

}  // namespace sparch

using namespace sparch;

extern "C" {

// Hidden size padded so that each of the cluster's four K quarters is a whole number of 64-neuron K blocks.
int sparch_recur_tc_padded(int H) { return ((H + 255) / 256) * 256; }

size_t sparch_recur_bwd_tc_image_bytes(int H) {
  const int Hp = sparch_recur_tc_padded(H);
  return (size_t)Hp * Hp * 2 * sizeof(__half);  // hi + lo
}

size_t sparch_recur_bwd_tc_workspace(int Be, int T, int H) {
  const int Hp = sparch_recur_tc_padded(H);
  const size_t groups = (size_t)(Be + TC_ROWS - 1) / TC_ROWS;
  return 2 * (2 * groups * TC_ROWS * Hp * sizeof(__half))  // panels hi, lo
         + 2 * groups * (Hp / 32) * TC_ROWS * sizeof(float)  // chunk maxima
         + (size_t)Be * T * sizeof(float)                    // gmax
         + groups * (Hp / 64) * sizeof(int) + 256;
}

// V (H,H) raw recurrent weight -> swizzled fp16 hi/lo UMMA tiles of V0 (zero diagonal); meta from
// sparch_recur_prepare (meta[0] = E, max|V0| < 2^E).
int sparch_recur_prepare_tc(const float* V, int H, void* img, const int* meta, sparch_stream_t st_) {
  SPARCH_REQUIRE(V && H > 0 && img && meta, "null pointer");
  const int Hp = sparch_recur_tc_padded(H), KB = Hp / (64 * TC_CL);
  int64_t total = (int64_t)Hp * Hp / 8;   // one thread per 8-element chunk (hi and lo)
  int nb = (int)((total + 255) / 256);
  if (nb > sm_count() * 16) nb = sm_count() * 16;
  vprep_umma_kernel<<<nb, 256, 0, as_stream(st_)>>>(V, H, Hp, KB, meta, reinterpret_cast<__half*>(img));
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_recur_bwd_tc(int kind, const float* G, const float* U, const float* W, const float* alpha,
                        const float* beta, const float* a, const float* b, const void* img, const int* meta,
                        const float* u0, const float* w0, const float* s0, float theta, float* dI, float* p_alpha,
                        float* p_beta, float* p_a, float* p_b, void* workspace, int reduced, int Be, int T, int H,
                        const float* gmax_in, sparch_stream_t st_) {
  SPARCH_REQUIRE(kind == SPARCH_RLIF || kind == SPARCH_RADLIF, "recurrent kinds only");
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0, "bad shape");
  if (Be == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && U && alpha && img && meta && u0 && s0 && dI && p_alpha && workspace, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (W && beta && a && b && w0 && p_beta && p_a && p_b),
                 "adaptive kind needs W, beta, a, b, w0 and the partial buffers");
  const int Hp = sparch_recur_tc_padded(H), KB = Hp / (64 * TC_CL);
  const size_t smem = rec_bwd_tc_smem(KB);
  SPARCH_REQUIRE(smem + 768 <= 227 * 1024, "hidden size too large for the resident V0 tiles");
  cudaStream_t st = as_stream(st_);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    SPARCH_CUDA(cudaFuncSetAttribute(rec_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 768));
    SPARCH_CUDA(cudaFuncSetAttribute(rec_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 768));
  }
  const int groups = (Be + TC_ROWS - 1) / TC_ROWS, slices = Hp / TC_COLS;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t panel_bytes = (size_t)2 * groups * TC_ROWS * Hp * sizeof(__half);
  __half* panel_hi = reinterpret_cast<__half*>(ws);
  __half* panel_lo = reinterpret_cast<__half*>(ws + panel_bytes);
  float* cmax = reinterpret_cast<float*>(ws + 2 * panel_bytes);
  float* gmax = cmax + (size_t)2 * groups * (Hp / 32) * TC_ROWS;
  int* counters = reinterpret_cast<int*>(gmax + (size_t)Be * T);
  SPARCH_CUDA(cudaMemsetAsync(ws, 0, 2 * panel_bytes, st));  // rows beyond Be / columns beyond H read as zero
  SPARCH_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * groups * (Hp / 64), st));
  if (gmax_in) {  // the producer of G already left the row maxima (sparch_spike_post_bwd)
    gmax = const_cast<float*>(gmax_in);
  } else {
    const int64_t warps = (int64_t)Be * T;
    gmax_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(G, Be, T, H, gmax);
    SPARCH_LAUNCH_OK();
  }
  TcMaps maps;
  memset(&maps, 0, sizeof maps);
  if (int e = make_map(&maps.hi, panel_hi, (long long)2 * groups * TC_ROWS, Hp, Hp, TC_ROWS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16))
    return e;
  if (int e = make_map(&maps.lo, panel_lo, (long long)2 * groups * TC_ROWS, Hp, Hp, TC_ROWS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16))
    return e;
  RecBwdTcArgs p{G, U, W, alpha, beta, a, b, u0, w0, s0, reinterpret_cast<const uint32_t*>(img), meta, gmax, theta,
                 dI, p_alpha, p_beta, p_a, p_b, panel_hi, panel_lo, cmax, Be, T, H, Hp, KB, reduced ? 1 : 0, recur_debug_buffer(), recur_debug_flags()};
  const void* fn = adapt ? (const void*)rec_bwd_tc_kernel<true> : (const void*)rec_bwd_tc_kernel<false>;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = TC_CL;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;  // every CTA of a launch must be resident: they wait on each other
  attrs[1].val.cooperative = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attrs;
  // Nsight Compute (2025.2) cannot replay a launch that is both clustered and cooperative; SPARCH_B200_TC_COOP=0
  // drops the cooperative attribute for profiling runs (co-residency then rests on the occupancy query below alone).
  static const bool coop = !(getenv("SPARCH_B200_TC_COOP") && getenv("SPARCH_B200_TC_COOP")[0] == '0');
  cfg.numAttrs = coop ? 2 : 1;
  cfg.gridDim = dim3(slices, 1);
  int max_clusters = 0;
  SPARCH_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, fn, &cfg));
  const int gmaxl = max_clusters * TC_CL / slices;
  SPARCH_REQUIRE(gmaxl >= 1, "hidden size needs more co-resident 4-CTA clusters than the GPU can hold");
  for (int g0 = 0; g0 < groups; g0 += gmaxl) {
    int gn = groups - g0 < gmaxl ? groups - g0 : gmaxl;
    cfg.gridDim = dim3(slices, gn);
    int group0 = g0, ngt = groups;
    void* args[] = {(void*)&maps, (void*)&p, (void*)&group0, (void*)&ngt, (void*)&counters};
    SPARCH_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  }
  return SPARCH_OK;
}

}  // extern "C"

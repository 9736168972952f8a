// Reverse-time recurrence of the recurrent kinds with dI_{t+1} @ V0^T on tcgen05 (alternative to
// rec_bwd_persist_kernel in recur.cu; same decomposition, same BPTT update, same hand-over idea).
//
//   CTA (slice, group) = 32 presynaptic neurons j x 64 batch rows, persistent over all T steps.
//   B operand: the slice of V0 (rows j, K = all neurons c) as fp16 hi + lo, K-major SWIZZLE_128B tiles,
//              resident in shared memory (KB x 2 x 4 KB).
//   A operand: dI_{t+1} of the 64 rows, handed over through L2 as two plain row-major fp16 matrices
//              (hi, lo), fetched by TMA (64 rows x 64 K boxes, SWIZZLE_128B) into a 4-stage ring.
//   D: 64 x 32 fp32 in 32 TMEM columns; per K16 three UMMAs (hi*hi, hi*lo, lo*hi), 192 per step, issued
//      by one thread; the update warps read D with tcgen05.ld (M = 64 layout: rows 16q..16q+15 live in
//      TMEM lanes 32q..32q+15).
//
// STATUS (measured on B200, cfg4 layer shape Be 256, T 100, H 1024): results agree with the mma.sync
// kernel to 4e-7 of the gradient maximum, but a step takes 22k cycles against 17.6k: the issuing thread
// spends ~58 cycles per tcgen05.mma at M = 64, N = 32 (704 cycles per 64-wide K block whether 4 or 12 of
// the 12 UMMA slots execute, with or without TMA traffic, one or six TMEM accumulators), i.e. 11k
// cycles per step for the 192 UMMAs the three fp16 terms need -- the per-instruction cost, not the
// tensor pipe (16-cycle floor) or L2, is the bound, and V0's fp16 hi+lo residency (128 KB for 32
// neurons) rules out a wider N.  Kept as an opt-in alternative (SPARCH_B200_BWD=tc), not the default.
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

#include "cell_math.cuh"
#include "common.cuh"
#include "tcgen05_utils.cuh"

namespace sparch {

constexpr int TC_ROWS = 64;            // batch rows per CTA
constexpr int TC_COLS = 32;            // neurons per CTA
constexpr int TC_STAGES = 4;           // A ring depth
constexpr int TC_STAGE_BYTES = 16384;  // 64 rows x 64 K x (hi, lo) fp16
constexpr int TC_THREADS = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 update
constexpr int TC_VSCALE_EXP = 13;

struct TcMaps {
  CUtensorMap hi, lo;
};

struct RecBwdTcArgs {
  const float *G, *U, *W, *alpha, *beta, *a, *b, *u0, *w0, *s0;
  const uint32_t* img;   // [slice][kb][part][32 rows x 128 B] swizzled fp16 tiles of V0 (vprep_umma_kernel)
  const int* meta;
  const float* gmax;     // [Be][T] row maxima of |G|
  float theta;
  float *dI, *p_alpha, *p_beta, *p_a, *p_b;
  __half *panel_hi, *panel_lo;  // [2][groups*64][Hp]
  float* cmax;                  // [2][groups][Hp/32][64] chunk maxima of |dI_t|
  int Be, T, H, Hp, KB;
  long long* dbg;  // optional [T][8] phase clocks of CTA (0,0) (profiling aid), normally NULL
  int dbg_flags;   // profiling experiments (results invalid): 1 hi*hi UMMA only, 4 no TMA traffic, 16 no UMMA
};

// V0 slice as UMMA B tiles: element (n, k) of k-block kb: halves offset n*64 + ((k/8) ^ (n&7))*8 + k%8.
__global__ void vprep_umma_kernel(const float* __restrict__ V, int H, int Hp, int KB, const int* __restrict__ meta,
                                  __half* __restrict__ img) {
  const int64_t per_slice = (int64_t)KB * 2 * 2048;
  const int64_t total = per_slice * (Hp / TC_COLS);
  const float sc = ldexpf(1.0f, TC_VSCALE_EXP - meta[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int slice = (int)(i / per_slice);
    int r = (int)(i % per_slice);
    const int e = r & 7, chunk_sw = (r >> 3) & 7, n = (r >> 6) & 31, part = (r >> 11) & 1, kb = r >> 12;
    const int k = ((chunk_sw ^ (n & 7)) << 3) + e;
    const int row = slice * TC_COLS + n, col = kb * 64 + k;  // V[row = presynaptic j][col = neuron c]
    float x = 0.f;
    if (row < H && col < H && row != col) x = V[(int64_t)row * H + col] * sc;
    const __half hi = __float2half_rn(x);
    img[i] = part ? __float2half_rn(x - __half2float(hi)) : hi;
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
  const uint32_t vimg = base;                                   // KB * 8 KB
  const uint32_t ring = base + (uint32_t)p.KB * 8192;           // TC_STAGES * 16 KB
  const uint32_t bars = ring + TC_STAGES * TC_STAGE_BYTES;      // full[4], empty[4], acc_full, acc_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tsm + (size_t)p.KB * 8192 + TC_STAGES * TC_STAGE_BYTES + 128);
  float* spart = reinterpret_cast<float*>(tsm + (size_t)p.KB * 8192 + TC_STAGES * TC_STAGE_BYTES + 256);  // [64][4]
  __shared__ float sprm[6][TC_COLS];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, group = group0 + blockIdx.y, row0 = group * TC_ROWS;
  const int nslices = gridDim.x, NCH = p.Hp / 32;
  int* ctr = counters + group;
  const bool dbg_cta = p.dbg && blockIdx.x == 0 && blockIdx.y == 0;

  {  // resident V0 tiles
    const uint4* src = reinterpret_cast<const uint4*>(p.img + (size_t)slice * p.KB * 2048);
    uint4* dst = reinterpret_cast<uint4*>(tsm);
    for (int i = tid; i < p.KB * 512; i += TC_THREADS) {
      uint32_t s = smem_u32(dst + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(src + i));
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
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (TC_STAGES + s), 1);
    }
    mbar_init(bars + 8 * (2 * TC_STAGES), 1);      // acc_full: one tcgen05.commit
    mbar_init(bars + 8 * (2 * TC_STAGES + 1), 8);  // acc_empty: one arrival per update warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // make the cp.async-written tiles visible to the tensor core's (async proxy) reads
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const float rs = ldexpf(1.0f, p.meta[0] - TC_VSCALE_EXP);

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the steps; lane 0 does the work) =====
    int it = 0;
    for (int t = p.T - 2; t >= 0; --t) {
      const int rbuf = (t + 1) & 1;
      if (lane == 0) {  // panel(t+1) published by every slice of the group
        const int target = nslices * (p.T - 1 - t);
        const long long t0 = clock64();
        while (true) {
          int v;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
          if (v >= target) break;
          if (clock64() - t0 > 4000000000LL) __trap();
        }
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of other SMs -> TMA reads
        if (dbg_cta) p.dbg[t * 8 + 0] = clock64();
      }
      __syncwarp();
      asm volatile("bar.arrive 2, 288;" ::: "memory");    // the update warps may read the chunk maxima now
      if (lane == 0) {
        const int y = (rbuf * ngroups_total + group) * TC_ROWS;
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(bars + 8 * (TC_STAGES + s), ph ^ 1);
          if (p.dbg_flags & 4) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * s) : "memory");
            continue;
          }
          mbar_expect_tx(bars + 8 * s, TC_STAGE_BYTES);
          const uint32_t sa = ring + s * TC_STAGE_BYTES;
          tma_load_2d(sa, &maps.hi, kb * 64, y, bars + 8 * s);
          tma_load_2d(sa + 8192, &maps.lo, kb * 64, y, bars + 8 * s);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // kind::f16, fp16 x fp16 -> fp32, both K-major, M = 64, N = 32
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_COLS >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
      int it = 0, step = 0;
      for (int t = p.T - 2; t >= 0; --t, ++step) {
        mbar_wait(bars + 8 * (2 * TC_STAGES + 1), (step & 1) ^ 1);  // D drained by the update warps
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < p.KB; ++kb, ++it) {
          const int s = it % TC_STAGES;
          const uint32_t ph = (it / TC_STAGES) & 1;
          mbar_wait(bars + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (dbg_cta && kb == 0) p.dbg[t * 8 + 1] = clock64();
          if (dbg_cta && t == p.T / 2) p.dbg[p.T * 8 + kb * 4] = clock64();
          const uint32_t sa = ring + s * TC_STAGE_BYTES;
          const uint64_t a_hi = make_desc_k_sw128(sa), a_lo = make_desc_k_sw128(sa + 8192);
          const uint64_t b_hi = make_desc_k_sw128(vimg + (uint32_t)(kb * 2 + 0) * 4096);
          const uint64_t b_lo = make_desc_k_sw128(vimg + (uint32_t)(kb * 2 + 1) * 4096);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (p.dbg_flags & 16) continue;
            umma_f16(tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (p.dbg_flags & 1) continue;
            umma_f16(tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
            umma_f16(tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
          }
          if (dbg_cta && t == p.T / 2) p.dbg[p.T * 8 + kb * 4 + 1] = clock64();
          umma_commit(bars + 8 * (TC_STAGES + s));
          if (dbg_cta && t == p.T / 2) p.dbg[p.T * 8 + kb * 4 + 2] = clock64();
        }
        umma_commit(bars + 8 * (2 * TC_STAGES));
        if (dbg_cta) p.dbg[t * 8 + 2] = clock64();
      }
    }
  } else {
    // ===== update warps: BPTT for the 64 x 32 block, 8 neurons of one row per thread =====
    const int uw = warp - 2, q = warp & 3, half = uw >> 2;  // TMEM lane quarter of this warp, column half
    const int r = 16 * q + (lane & 15), cq = 2 * half + (lane >> 4);  // row of the block, quarter of its 32 columns
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
      spart[r * 4 + cq] = m_next;
      upd_sync();
      m_next = fmaxf(fmaxf(spart[r * 4 + 0], spart[r * 4 + 1]), fmaxf(spart[r * 4 + 2], spart[r * 4 + 3]));
      const float g_row = row < p.Be ? p.gmax[(size_t)row * p.T + t] : 0.f;
      const float s_t = scale_from_max(fmaxf(m_next, 0.25f * g_row));
      if (t < p.T - 1) {
        // ---- D = dI_{t+1} (scaled) @ V0^T (scaled): read this thread's 8 values
        mbar_wait(bars + 8 * (2 * TC_STAGES), step & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_on) p.dbg[t * 8 + 3] = clock64();
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * half);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 8 * (2 * TC_STAGES + 1)) : "memory");
        // lanes 0..15 hold row r's 16 columns; lanes 16..31 take columns 8..15 from lane-16
        const float k = inv_s_next * rs;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t hi8 = __shfl_sync(0xffffffffu, v[8 + j], lane & 15);
          recb[j] = __uint_as_float(lane < 16 ? v[j] : hi8) * k;
        }
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
      if (t > 0) {
        // ---- hand dI_t over: chunk maximum (for the next scale) and fp16 hi/lo rows scaled by s_t
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(d[i]));
        upd_sync();                       // spart was read above by everyone
        spart[r * 4 + cq] = m;
        upd_sync();
        if (cq == 0) {
          const float cmx = fmaxf(fmaxf(spart[r * 4 + 0], spart[r * 4 + 1]), fmaxf(spart[r * 4 + 2], spart[r * 4 + 3]));
          p.cmax[(((size_t)wbuf * ngroups_total + group) * NCH + slice) * TC_ROWS + r] = cmx;
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
        upd_sync();                       // every update thread's panel stores are issued
        if (tid == 64) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
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
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
  }
}

static size_t rec_bwd_tc_smem(int KB) { return (size_t)KB * 8192 + TC_STAGES * TC_STAGE_BYTES + 256 + 1024 + 1024; }

}  // namespace sparch

using namespace sparch;

extern "C" {

// Hidden size padded to the 64-neuron K blocks of the tcgen05 reverse kernel.
int sparch_recur_tc_padded(int H) { return ((H + 63) / 64) * 64; }

size_t sparch_recur_bwd_tc_image_bytes(int H) {
  const int Hp = sparch_recur_tc_padded(H);
  return (size_t)(Hp / TC_COLS) * (Hp / 64) * 2 * 2048 * sizeof(__half);
}

size_t sparch_recur_bwd_tc_workspace(int Be, int T, int H) {
  const int Hp = sparch_recur_tc_padded(H);
  const size_t groups = (size_t)(Be + TC_ROWS - 1) / TC_ROWS;
  return 2 * (2 * groups * TC_ROWS * Hp * sizeof(__half))  // panels hi, lo
         + 2 * groups * (Hp / 32) * TC_ROWS * sizeof(float)  // chunk maxima
         + (size_t)Be * T * sizeof(float)                    // gmax
         + groups * sizeof(int) + 256;
}

// V (H,H) raw recurrent weight -> swizzled fp16 hi/lo UMMA tiles of V0 (zero diagonal); meta from
// sparch_recur_prepare (meta[0] = E, max|V0| < 2^E).
int sparch_recur_prepare_tc(const float* V, int H, void* img, const int* meta, sparch_stream_t st_) {
  SPARCH_REQUIRE(V && H > 0 && img && meta, "null pointer");
  const int Hp = sparch_recur_tc_padded(H), KB = Hp / 64;
  int64_t total = (int64_t)(Hp / TC_COLS) * KB * 2 * 2048;
  int nb = (int)((total + 255) / 256);
  if (nb > sm_count() * 16) nb = sm_count() * 16;
  vprep_umma_kernel<<<nb, 256, 0, as_stream(st_)>>>(V, H, Hp, KB, meta, reinterpret_cast<__half*>(img));
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_recur_bwd_tc(int kind, const float* G, const float* U, const float* W, const float* alpha,
                        const float* beta, const float* a, const float* b, const void* img, const int* meta,
                        const float* u0, const float* w0, const float* s0, float theta, float* dI, float* p_alpha,
                        float* p_beta, float* p_a, float* p_b, void* workspace, int Be, int T, int H,
                        sparch_stream_t st_) {
  SPARCH_REQUIRE(kind == SPARCH_RLIF || kind == SPARCH_RADLIF, "recurrent kinds only");
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0, "bad shape");
  if (Be == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && U && alpha && img && meta && u0 && s0 && dI && p_alpha && workspace, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (W && beta && a && b && w0 && p_beta && p_a && p_b),
                 "adaptive kind needs W, beta, a, b, w0 and the partial buffers");
  const int Hp = sparch_recur_tc_padded(H), KB = Hp / 64;
  const size_t smem = rec_bwd_tc_smem(KB);
  SPARCH_REQUIRE(smem <= 225 * 1024, "hidden size too large for the resident V0 tiles");
  cudaStream_t st = as_stream(st_);
  static bool attr_set = false;
  if (!attr_set) {
    SPARCH_CUDA(cudaFuncSetAttribute(rec_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    SPARCH_CUDA(cudaFuncSetAttribute(rec_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    attr_set = true;
  }
  const int groups = (Be + TC_ROWS - 1) / TC_ROWS, slices = Hp / TC_COLS;
  SPARCH_REQUIRE(slices <= sm_count(), "hidden size needs more co-resident CTAs than the GPU has SMs");
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t panel_bytes = (size_t)2 * groups * TC_ROWS * Hp * sizeof(__half);
  __half* panel_hi = reinterpret_cast<__half*>(ws);
  __half* panel_lo = reinterpret_cast<__half*>(ws + panel_bytes);
  float* cmax = reinterpret_cast<float*>(ws + 2 * panel_bytes);
  float* gmax = cmax + (size_t)2 * groups * (Hp / 32) * TC_ROWS;
  int* counters = reinterpret_cast<int*>(gmax + (size_t)Be * T);
  SPARCH_CUDA(cudaMemsetAsync(ws, 0, 2 * panel_bytes, st));  // rows beyond Be / columns beyond H read as zero
  SPARCH_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * groups, st));
  {
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
                 dI, p_alpha, p_beta, p_a, p_b, panel_hi, panel_lo, cmax, Be, T, H, Hp, KB, recur_debug_buffer(), recur_debug_flags()};
  const int gmaxl = sm_count() / slices;
  for (int g0 = 0; g0 < groups; g0 += gmaxl) {
    int gn = groups - g0 < gmaxl ? groups - g0 : gmaxl;
    dim3 cgrid(slices, gn);
    int group0 = g0, ngt = groups;
    void* args[] = {(void*)&maps, (void*)&p, (void*)&group0, (void*)&ngt, (void*)&counters};
    const void* fn = adapt ? (const void*)rec_bwd_tc_kernel<true> : (const void*)rec_bwd_tc_kernel<false>;
    SPARCH_CUDA(cudaLaunchCooperativeKernel(fn, cgrid, dim3(TC_THREADS), args, smem, st));
  }
  return SPARCH_OK;
}

}  // extern "C"

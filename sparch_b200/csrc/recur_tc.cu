// Reverse-time recurrence of the recurrent kinds (BPTT of snns.py:554-578 / 696-727) with recb_t = dI_{t+1} @ V0^T on
// tcgen05, split-K over a thread-block cluster, tagged-word hand-over between the steps.
//
//   Cluster of 4 CTAs = 64 batch rows x 128 presynaptic neurons j, persistent over all T steps.  CTA rank r of the
//   cluster multiplies over its QUARTER of K (neurons c in [r Hp/4, (r+1) Hp/4)) for all 128 neurons j of the cluster,
//   then the four partial products are reduce-scattered through distributed shared memory: rank d sums and owns 32 of
//   the 128 neurons.  The neurons a CTA owns lie INSIDE the K quarter it consumes (cluster c, rank r owns neurons
//   r Hp/4 + 32 c .. + 31), so the all-gather of dI_{t+1} closes over groups of Hp/128 CTAs (rank r of every cluster):
//   a CTA waits only for the producers whose data it multiplies.
//
//   The product is issued TRANSPOSED: D^T[j][row] = sum_c V0[j][c] dI[row][c], i.e. A = the resident V0 tile (M = 128
//   neurons, full-rate UMMA shape), B = the dI panel (N = 64 rows), D = 128 TMEM lanes x 64 fp32 columns.  Every TMEM
//   lane is used, a warp's tcgen05.ld returns one neuron per lane (32 rows each), and the lane quarter q of TMEM holds
//   exactly the neurons of cluster rank q: the reduce-scatter is 8 st.async per update warp.  (The first version
//   computed D = dI V0^T with M = 64: half-rate UMMA, half of the TMEM lanes idle, 16 stores per warp.)
//
//   A operand: V0[128 j][Hp/4 c] as fp16 hi + lo, resident in TENSOR MEMORY (lane = neuron, 32-bit column = two K;
//              Hp/256 x 2 x 32 columns = 256 of the 512 columns at H = 1024), copied once per launch from the swizzled
//              global image with tcgen05.st.  With A in TMEM the tensor core reads only the 1-2 KB of dI per UMMA from
//              shared memory: 32 cycles per M128 N64 K16 UMMA instead of 48 (tools/ubench/umma_f16_rate.cu).
//   B operand: dI_{t+1} of the 64 rows x this rank's K quarter, fp16 hi + lo, written into shared memory by the update
//              warps themselves (K-major SWIZZLE_128B, one 16 KB stage per 64-neuron K block, no ring).
//   Two CHAINS per CTA: rows 0..31 and 32..63 of the group are different batch rows, i.e. independent recurrences.
//              Each half has its own four update warps, barriers, accumulator (N = 32) and step counter; the MMA warp
//              serves whichever chain has a K block ready.
//   Hand-over through L2 without flags, fences or TMA: a producer stores one 32-bit word {hi, lo} per element with
//   st.relaxed.gpu; the least significant bit of lo is a TAG that toggles every time the (double-buffered) word is
//   rewritten, so a consumer polls the data itself (ld.relaxed.gpu.v4, 16 loads in flight per thread) and knows from
//   the word alone whether it is this step's value -- one L2 round trip from "dI_t computed" to "operand in shared
//   memory" (the first version: panel stores, red.release, ld.acquire poll, fence.proxy.async, TMA = three round trips,
//   ~5.5 k of its 13.5 k cycles per step).  lo keeps 10 of its 11 significant bits: 21 bits per element.
//
// Scaling.  fp16 needs a scale, and the tensor core accumulates over the whole K range without intervention, so the
// scale is per ROW.  The row maximum of |dI_t| is only known once all slices have produced their part, so the scale of
// step t is derived one step late: M_t = max(rowmax|dI_{t+1}|, 0.25 * rowmax|g_t|), s_t = 2^(4 - exp(M_t)).
// |dI_t| can exceed M_t by the growth of one step (bounded by a few hundred), far inside fp16's range above 16; values
// below it only lose ABSOLUTE precision.  rowmax|dI_{t+1}| costs no extra exchange: each rank takes the maximum of the
// K quarter it has just loaded and sends it to its three cluster peers with the reduce-scatter (4 ranks = 4 quarters =
// the whole row); every CTA of the row group derives the same s_t from the same numbers.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "cell_math.cuh"
#include "common.cuh"
#include "tcgen05_utils.cuh"

namespace sparch {

constexpr int TC_ROWS = 64;            // batch rows per CTA (UMMA N)
constexpr int TC_COLS = 32;            // neurons updated per CTA
constexpr int TC_CL = 4;               // CTAs per cluster = K quarters
constexpr int TC_N = TC_COLS * TC_CL;  // neurons per cluster (UMMA M)
constexpr int TC_STAGE_BYTES = 16384;  // 64 rows x 64 K x (hi, lo) fp16
constexpr int TC_RECV_BYTES = TC_CL * TC_COLS * TC_ROWS * 4;  // [source rank][row quad][neuron][4 rows] fp32
constexpr int TC_QMAX_BYTES = TC_CL * TC_ROWS * 4;            // [source rank][row] quarter maxima of |dI_{t+1}|
constexpr int TC_THREADS = 320;        // warp 0 spare, warp 1 MMA, warps 2..9 load + update
constexpr int TC_VSCALE_EXP = 13;
constexpr uint32_t TC_RECV_TX = (TC_RECV_BYTES + TC_QMAX_BYTES) / 2;   // per chain (32 of the 64 rows) and step
constexpr int TC_NBUF = 2;               // hand-over buffers in rotation (see tc_tag)
constexpr uint32_t TC_D_COL = 256;      // TMEM: V0 hi / lo in columns [0, 256), D of chain c in [256 + 32 c, 288 + 32 c)

struct RecBwdTcArgs {
  const float *G, *U, *W, *alpha, *beta, *a, *b, *u0, *w0, *s0;
  const uint32_t* img;   // [CTA][kb][part][128 rows x 128 B] swizzled fp16 tiles of V0 (vprep_umma_kernel)
  const int* meta;
  const float* gmax;     // [Be][T] row maxima of |G|
  float theta;
  float *dI, *p_alpha, *p_beta, *p_a, *p_b;
  uint32_t* panel;       // [2][groups*64][Hp] words {fp16 hi, fp16 lo with the tag in its least significant bit}
  int Be, T, H, Hp, KB;  // Hp: H padded to 256; KB = Hp / 256 k-blocks of 64 per CTA
  int reduced;
  int w_every;     // C > 0: W is the checkpoint tape (Be, ceil(T / C), H) of sparch_recur_fwd_tc_bidir: w at the end of
                   // every chunk of C steps; the steps in between are recomputed here (see the update's second half)
  long long* dbg;  // optional [2T][8] phase clocks of CTA (0,0) (profiling aid), normally NULL
  int dbg_flags;   // profiling experiments (results invalid): 16 no UMMA
};

// First neuron owned by CTA `slice` = (cluster slice / 4, rank slice % 4): inside the K quarter the rank consumes.
__host__ __device__ __forceinline__ int tc_own0(int slice, int Hp) {
  return (slice % TC_CL) * (Hp / TC_CL) + TC_COLS * (slice / TC_CL);
}

// V0 as UMMA tiles.  CTA bx = (cluster bx / 4, rank bx % 4) holds tile rows n = 32 d + i <-> neuron i of cluster rank
// d (presynaptic j = tc_own0(4 (bx / 4) + d) + i) and K = (bx % 4) Hp / 4 + 64 kb + k; element (n, k) of a tile: halves
// offset n*64 + ((k/8) ^ (n&7))*8 + k%8.
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
    const int row = tc_own0((bx / TC_CL) * TC_CL + n / TC_COLS, Hp) + (n % TC_COLS);  // V[row = presynaptic j][col = neuron c]
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

// Exponent E of the power-of-two scale 2^E that brings m into [8, 16) (4 minus the binary exponent of m, clamped): pure
// bit operations, the same result in every CTA that sees the same m.  -123 <= E <= 104.
__device__ __forceinline__ int scale_exp(float m) {
  int e = 0;
  if (m > 0.f && m <= 3.0e38f) {
    const int ef = (int)((__float_as_uint(m) >> 23) & 0xffu);
    e = ef ? ef - 126 : -126;
  }
  return 4 - max(e, -100);
}

// Tape values are read once: non-coherent path, no L1 allocation (with 225 KB of shared memory in use the L1 holds
// a fraction of one step's 32 KB).
__device__ __forceinline__ float ld_stream(const float* a) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(a));
  return v;
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* a) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
  return v;
}

__device__ __forceinline__ uint4 ld_weak_v4(const uint32_t* a) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint32_t* a) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}

#define TC_LD32(taddr, v)                                                                                                   \
  asm volatile(                                                                                                            \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "      \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                             \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),         \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),              \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),             \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                           \
      : "r"(taddr))

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ uint32_t tc_mapa(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && clock64() - t0 > 4000000000LL) __trap();
  }
}

// panel(t) lives in buffer (T - 1 - t) % TC_NBUF; the n-th use of a buffer carries tag (n & 1) ^ 1 (the buffers start
// zeroed: tag 0).  Two buffers are enough: a producer overwrites panel(t + 2) only after all its consumers have published
// panel(t + 1), which they do after reading it.  (Eight buffers in rotation were measured: no faster -- the ~25 B/clk
// at which a CTA's 64 KB arrive per step is not an effect of re-reading recently read addresses, although an isolated
// SM shows one: tools/ubench/l2_gather.cu.)
__device__ __forceinline__ uint32_t tc_tag(int T, int t) { return ((((T - 1 - t) / TC_NBUF) & 1) ^ 1); }
__device__ __forceinline__ int tc_buf(int T, int t) { return (T - 1 - t) % TC_NBUF; }

template <bool ADAPT, bool WCK>
__global__ void __launch_bounds__(TC_THREADS, 1)
rec_bwd_tc_kernel(const RecBwdTcArgs p, const int group0, const int ngroups_total) {
  extern __shared__ unsigned char tsm_raw[];
  const uint32_t raw = smem_u32(tsm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* tsm = tsm_raw + (base - raw);
  const uint32_t v_bytes = (uint32_t)p.KB * 2 * (TC_N * 128);     // KB x (hi, lo) x 16 KB of V0 tiles per CTA (global image)
  const uint32_t stage = base;                                    // KB x 16 KB: dI_{t+1} tiles (hi 8 KB, lo 8 KB)
  const uint32_t recv = stage + (uint32_t)p.KB * TC_STAGE_BYTES;  // partial products from the four ranks
  const uint32_t qmax = recv + TC_RECV_BYTES;                     // quarter maxima from the four ranks
  const uint32_t bars = qmax + TC_QMAX_BYTES;                     // full[4], acc_full, recv, free
  unsigned char* tail = tsm + (size_t)p.KB * TC_STAGE_BYTES;
  const unsigned char* recv_p = tail;
  const float* qmax_p = reinterpret_cast<const float*>(tail + TC_RECV_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + TC_RECV_BYTES + TC_QMAX_BYTES + 192);
  // checkpoint-tape mode: [8 rows][256 update threads] floats w of the step last visited + the same for a freshly loaded
  // checkpoint (kept out of the register file: touched once per step, in the waiting window)
  float* wc_sm = reinterpret_cast<float*>(tail + TC_RECV_BYTES + TC_QMAX_BYTES + 256);
  // barriers of chain c (the 32-row half c of the group) at index 8 c + {0..3 full[kb], 4 acc_full, 5 recv, 6 free}
  constexpr int B_ACC_FULL = 4, B_RECV = 5, B_FREE = 6, B_CHAIN = 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, group = group0 + blockIdx.y, row0 = group * TC_ROWS;
  const int rank = blockIdx.x % TC_CL;  // == %cluster_ctarank for cluster dims (4, 1, 1)
  const int own0 = tc_own0(slice, p.Hp);
  const bool dbg_cta = p.dbg && blockIdx.x == 0 && blockIdx.y == 0;

  if (tid == 0) {
    for (int c = 0; c < 2; ++c) {
      const uint32_t cb = bars + 8 * B_CHAIN * c;
      for (int kb = 0; kb < 4; ++kb) mbar_init(cb + 8 * kb, 4);   // one arrival per update warp of the chain
      mbar_init(cb + 8 * B_ACC_FULL, 1);                          // one tcgen05.commit
      mbar_init(cb + 8 * B_RECV, 1);  // armed per step with the bytes the four ranks deliver (st.async complete_tx)
      mbar_init(cb + 8 * B_FREE, TC_CL * 4);                      // every update warp of the chain, cluster-wide, has read
      mbar_expect_tx(cb + 8 * B_RECV, TC_RECV_TX);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d = tmem + TC_D_COL;
  if (warp >= 2) {
    // ---- resident V0: this CTA's 128 neurons x K quarter, fp16 hi and lo, as the A operand IN TENSOR MEMORY (lane =
    // neuron n, 32-bit column = two consecutive K): hi at columns [0, 32 KB), lo at [32 KB, 64 KB).  The tensor core then
    // reads only the 2 KB of dI per UMMA from shared memory and runs at its full rate (A from shared memory: 6 KB per
    // UMMA, 68 cycles instead of 32 at M = 128, N = 64).  Warp (q, part) copies part (hi / lo) of the neurons of TMEM
    // lane quarter q out of the global image: each lane reads the 128 contiguous bytes of its tile row, un-swizzled by
    // address.
    const int q = warp & 3, part = (warp - 2) >> 2, n = 32 * q + lane;
    const unsigned char* img = reinterpret_cast<const unsigned char*>(p.img) + (size_t)slice * v_bytes;
    for (int kb = 0; kb < p.KB; ++kb) {
      const unsigned char* rowp = img + ((size_t)(kb * 2 + part) * TC_N + n) * 128;
      uint32_t w[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((c ^ (n & 7)) << 4));
        w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
      }
      const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(part * 32 * p.KB + kb * 32);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
          "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
          "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]),
          "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]),
          "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]),
          "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // cluster-wide: nobody may arrive on a peer's barriers before they are initialised
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const float rs = ldexpf(1.0f, p.meta[0] - TC_VSCALE_EXP);

  if (warp == 1) {
    // ===== MMA issuer: the warp stays converged, one elected lane issues each K block's UMMAs as straight-line code
    // (under a plain `if (lane == 0)` ptxas wraps every UMMA in an ELECT / PLOP3 / BRA.U.ANY retry loop, ~45 cycles
    // per instruction whatever its shape: measured with tools/ubench/umma_i8_ts.cu) =====
    // kind::f16, fp16 x fp16 -> fp32, both K-major, M = 128 (neurons), N = 64 (batch rows)
    const uint32_t idesc = (1u << 4) | ((uint32_t)((TC_ROWS / 2) >> 3) << 17) | ((uint32_t)(TC_N >> 4) << 24);
    const uint64_t x_desc = make_desc_k_sw128(stage);
    const uint32_t lo_col = 32 * p.KB;
    // The two 32-row chains of the group advance independently (different batch rows): the warp serves whichever has a
    // K block ready, so one chain's tensor work fills the other's L2 / DSMEM / update time.
    int cstep[2] = {0, 0}, ckb[2] = {0, 0};
    const int nsteps = p.T - 1;
    while (cstep[0] < nsteps || cstep[1] < nsteps) {
      bool idle = true;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (cstep[c] >= nsteps) continue;
        const uint32_t cb = bars + 8 * B_CHAIN * c;
        const int kb = ckb[c];
        uint32_t ready;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ready)
            : "r"(cb + 8 * kb), "r"((uint32_t)(cstep[c] & 1))
            : "memory");
        ready = __shfl_sync(0xffffffffu, ready, 0);   // one answer for the warp
        if (!ready) continue;   // the update warps have not yet written this K block of dI_{t+1} (and drained D)
        idle = false;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_cta && lane == 0 && c == 0 && kb == 0) p.dbg[(p.T + (p.T - 2 - cstep[0])) * 8 + 0] = clock64();
        __syncwarp();
        if (tc_elect_one()) {
          const uint64_t x_hi = x_desc + (uint64_t)(kb * (TC_STAGE_BYTES >> 4) + c * (4096 >> 4)), x_lo = x_hi + (8192 >> 4);
          const uint32_t v_hi = tmem + (uint32_t)(kb * 32), v_lo = v_hi + lo_col;
          const uint32_t dcol = tmem_d + 32 * c;
          if (!(p.dbg_flags & 16)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_ts(dcol, v_hi + 8 * k, x_hi + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              if (p.reduced) continue;  // reduced-precision mode: hi * hi only
              umma_f16_ts(dcol, v_lo + 8 * k, x_hi + 2 * k, idesc, 1u);
              umma_f16_ts(dcol, v_hi + 8 * k, x_lo + 2 * k, idesc, 1u);
            }
          }
          if (kb == p.KB - 1) umma_commit(cb + 8 * B_ACC_FULL);
        }
        __syncwarp();
        if (kb == p.KB - 1) {
          if (dbg_cta && lane == 0 && c == 0) p.dbg[(p.T + (p.T - 2 - cstep[0])) * 8 + 1] = clock64();
          ckb[c] = 0;
          ++cstep[c];
        } else {
          ckb[c] = kb + 1;
        }
      }
      // nothing ready: stay off the schedulers and the shared-memory pipe for a moment (a hot try_wait loop competes
      // with the update warps' loads and stores)
      if (idle) __nanosleep(40);
    }
  } else if (warp >= 2) {
    // ===== load + update warps.  Warp uw owns batch rows 8 uw .. 8 uw + 7 of the group; lane = neuron own0 + lane
    // (BPTT state of 8 (row, neuron) pairs per thread, coalesced 128-byte rows of G / U / W / dI per warp access).
    // As a loader the warp brings rows 8 uw .. + 7 of dI_{t+1}[:, K quarter] into the stage tiles; as TMEM reader it
    // holds lane quarter q = warp % 4 (= the neurons of cluster rank q) x rows 32 half .. 32 half + 31.
    const int uw = warp - 2, q = warp & 3, half = uw >> 2;
    const uint32_t cbar = bars + 8 * B_CHAIN * half;   // this warp's chain = the 32-row half of the group it belongs to
    const int col = own0 + lane, colc = min(col, p.H - 1);
    const bool col_live = col < p.H;
    const NeuronParams prm = load_params<ADAPT>(p.alpha, p.beta, p.a, p.b, min(col, p.H - 1));
    const float inv_oma = 1.0f / prm.oma;
    const int rowb = row0 + 8 * uw;                // first of this warp's 8 batch rows
    const int TH = p.T * p.H;                      // elements per batch row of the tapes (Be * T * H < 2^31 is checked)
    float du[8], dw[8], ut[8];
    const int NCK = WCK ? (p.T + p.w_every - 1) / p.w_every : 0;
    const float inv_beta = WCK ? 1.0f / prm.beta : 0.f;
    float* wc = wc_sm + (tid - 64);                // wc[256 k]: row k of this thread
    float pa = 0.f, pb = 0.f, pc = 0.f, pd = 0.f;  // parameter-gradient sums over this thread's 8 rows and all steps
    // The hand-over scale of row 8 uw + k is computed by lane k alone (it reads the quarter maxima and the row maximum of
    // g) and travels as an exponent: e_cur = exponent of s_t, e_next = exponent of s_{t+1} (the scale of the panel that
    // is multiplied at step t).
    int e_next = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      du[k] = dw[k] = 0.f;
      const bool live = col_live && rowb + k < p.Be;
      ut[k] = (live && p.T > 0) ? __ldcg(p.U + ((int64_t)(rowb + k) * p.T + (p.T - 1)) * p.H + col) : 0.f;
      if (WCK) wc[256 * k] = (live && p.T > 0) ? __ldcg(p.W + ((int64_t)(rowb + k) * NCK + (NCK - 1)) * p.H + col) : 0.f;
    }
    // loader geometry: load i = (kb = i / 4, j = i % 4) covers row 8 uw + 2 j + (lane / 16), columns 4 (lane % 16) ..
    // + 3 of K block kb: a half-warp reads the 256 contiguous bytes of one (row, K block)
    const int lrow = 8 * uw + (lane >> 4), c16 = lane & 15;
    const uint32_t* pan_rd0 = p.panel + ((size_t)group * TC_ROWS + lrow) * p.Hp + rank * (p.Hp / TC_CL) + 4 * c16;
    uint32_t* pan_wr0 = p.panel + ((size_t)group * TC_ROWS + 8 * uw) * p.Hp + col;
    const size_t pan_buf = (size_t)ngroups_total * TC_ROWS * p.Hp;   // words per buffer
    uint32_t sa_j[4];                              // stage address of this thread's 8 bytes of row lrow + 2 j (K block 0, hi tile)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = lrow + 2 * j;
      sa_j[j] = stage + r * 128 + ((((uint32_t)c16 >> 1) ^ ((uint32_t)r & 7)) << 4) + ((c16 & 1) << 3);
    }
    const int myrow = min(rowb + (lane & 7), p.Be - 1);     // the row whose scale this lane computes
    const bool myrow_live = rowb + (lane & 7) < p.Be;
    const bool dbg_on = dbg_cta && tid == 64;
    int step = 0;
    // t % C and t / C of the checkpoint tape as counters (no division inside the loop)
    int tm = WCK ? (p.T - 1) % p.w_every : 0, tq = WCK ? (p.T - 1) / p.w_every : 0;
    for (int t = p.T - 1; t >= 0; --t) {
      // ---- this step's tape values.  Issued AFTER the panel load phase, in flight while the tensor core multiplies: the
      // loader's 64 registers of panel words are dead by then (with both live, ptxas spilled loaded values at once,
      // which serialises the loads: 4.7 k cycles to issue 32 of them).  Always-valid addresses (clamped rows / column),
      // the select comes after the load: straight-line code; 32-bit element offsets.
      float gq[8], up[8], wp[8], grow_l;
      const int toff = t * p.H + colc;
      auto load_tape = [&]() {
        const float* up_base = t > 0 ? p.U + (toff - p.H) : p.u0 + colc;
        // w_{t-1}: from the full tape; or (checkpoint tape) loaded only where step t-1 closes a chunk -- everywhere else
        // the second half of the update recomputes it from w_t -- and w0 at t = 0
        const bool w_ck = WCK && t > 0;
        const bool w_load = ADAPT && (!w_ck || tm == 0);
        const float* wp_base = !ADAPT ? nullptr
                               : t == 0 ? p.w0 + colc
                               : w_ck   ? p.W + (tq - 1) * p.H + colc      // tm == 0: (t - 1) / C = t / C - 1
                                        : p.W + (toff - p.H);
        const int rs_prev = t > 0 ? TH : p.H;
        const int rs_w = t == 0 ? p.H : (w_ck ? NCK * p.H : TH);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rowc = min(rowb + k, p.Be - 1);
          gq[k] = ld_stream(p.G + toff + rowc * TH);
          up[k] = ld_stream(up_base + rowc * rs_prev);
          wp[k] = w_load ? ld_stream(wp_base + rowc * rs_w) : 0.f;
        }
        grow_l = ld_stream(p.gmax + myrow * p.T + t);
      };
      float recb[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) recb[k] = 0.f;
      float m_next_l = 0.f;                        // lanes 0..7: maximum of |dI_{t+1}| over row 8 uw + lane
      if (dbg_on) p.dbg[t * 8 + 0] = clock64();
      if (t < p.T - 1) {
        // ---- dI_{t+1}[rows of this warp][K quarter] -> stage tiles: poll the tagged words themselves
        const uint32_t* src = pan_rd0 + (size_t)tc_buf(p.T, t + 1) * pan_buf;
        const uint32_t em = tc_tag(p.T, t + 1) << 16;
        // Polling all 64 KB would load the L2 with 8 MB per round over the grid; the warp polls one word of every 128-byte
        // LINE instead (a line = one row of one producer warp = one store instruction of that warp: 8 rows x 8
        // producers = 64 lines, two per lane: lane l takes producer l % 8, rows l / 8 and l / 8 + 4).  Fresh lines say
        // "the data is there"; every data word is still checked by its own tag (and re-read until fresh), so nothing
        // rests on the line assumption.
        const uint32_t* smp = p.panel + (size_t)tc_buf(p.T, t + 1) * pan_buf + ((size_t)group * TC_ROWS + 8 * uw + (lane >> 3)) * p.Hp +
                              rank * (p.Hp / TC_CL) + 32 * (lane & 7);
        uint4 v[16];
        __half2 mj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mj[j] = __float2half2_rn(0.f);
        int issued = 0;
        const long long t0 = clock64();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb < p.KB) {
            while (true) {
              while (issued <= kb) {
                bool ok = true;
                if ((lane & 7) < 2 * p.KB) {
                  const uint32_t w0 = ld_relaxed_u32(smp), w1 = ld_relaxed_u32(smp + 4 * p.Hp);
                  ok = (((w0 ^ em) | (w1 ^ em)) & 0x10000u) == 0;
                }
                const uint32_t fresh = __ballot_sync(0xffffffffu, ok);
                int n = issued;
                while (n < p.KB && ((fresh >> (2 * n)) & 0x03030303u) == 0x03030303u) ++n;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if ((i >> 2) >= issued && (i >> 2) < n) {
                    const uint32_t* q_ = src + (2 * (i & 3)) * p.Hp + (i >> 2) * 64;
                    v[i] = ld_relaxed_v4(q_);
                  }
                issued = n;
                if (dbg_on) p.dbg[2 * p.T * 8 + kb] += 1;          // sample rounds entered at K block kb
                if (clock64() - t0 > 4000000000LL) __trap();
              }
              // every word carries its own tag (a balanced tree: the update warps are ALU-latency bound here)
              uint32_t bj[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 x = v[kb * 4 + j];
                bj[j] = ((x.x ^ em) | (x.y ^ em)) | ((x.z ^ em) | (x.w ^ em));
              }
              const bool bad = (((bj[0] | bj[1]) | (bj[2] | bj[3])) & 0x10000u) != 0;
              if (!__any_sync(0xffffffffu, bad)) break;
              issued = kb;                                         // a stale word after fresh lines: read the K block again
              if (dbg_on) p.dbg[2 * p.T * 8 + 4 + kb] += 1;
            }
            if (dbg_on && kb == 0) p.dbg[t * 8 + 1] = clock64();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // {hi, lo | tag} words -> hi pairs and lo pairs (the tag stays in lo's last bit: 2^-22 of the element, with
              // the 21 bits that survive the tagging anyway)
              const uint4 x = v[kb * 4 + j];
              const uint32_t h01 = __byte_perm(x.x, x.y, 0x5410), h23 = __byte_perm(x.z, x.w, 0x5410);
              const uint32_t l01 = __byte_perm(x.x, x.y, 0x7632), l23 = __byte_perm(x.z, x.w, 0x7632);
              const uint32_t sa = sa_j[j] + kb * TC_STAGE_BYTES;
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sa), "r"(h01), "r"(h23) : "memory");
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sa + 8192), "r"(l01), "r"(l23) : "memory");
              const __half2 a01 = __habs2(*reinterpret_cast<const __half2*>(&h01));
              const __half2 a23 = __habs2(*reinterpret_cast<const __half2*>(&h23));
              mj[j] = __hmax2(mj[j], __hmax2(a01, a23));
            }
            // generic-proxy stores -> the tensor core's async-proxy reads.  (The tcgen05.ld of D of the previous step was
            // fenced right after the drain: tcgen05.fence::before_thread_sync precedes this arrival in program order.)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cbar + 8 * kb) : "memory");
            if (dbg_on) p.dbg[(p.T + t) * 8 + 4 + kb] = clock64();
          }
        }
        if (dbg_on) p.dbg[t * 8 + 2] = clock64();
        // ---- maxima of |dI_{t+1}| over this K quarter: mj[j] covers row 2 j + lane / 16 of the warp (in units of
        // s_{t+1}); lanes 0 and 16 send them to all four ranks
        float mf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mf[j] = fmaxf(__low2float(mj[j]), __high2float(mj[j]));
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) mf[j] = fmaxf(mf[j], __shfl_xor_sync(0xffffffffu, mf[j], o));
        }
        // the peers' buffers are free once every update warp of the cluster has read the previous step's data
        if (step > 0) mbar_wait_cluster(cbar + 8 * B_FREE, (step - 1) & 1);
        {
          // rows 8 uw .. + 7 in units of 1 (not of s_{t+1}); lane d < 4 sends the 32 bytes to rank d: two 16-byte remote
          // stores per destination (one 4-byte store per row and destination cost ~2 k cycles of DSMEM transactions)
          float val[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            val[2 * j] = __shfl_sync(0xffffffffu, mf[j], 0);
            val[2 * j + 1] = __shfl_sync(0xffffffffu, mf[j], 16);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            val[k] *= __uint_as_float((uint32_t)(127 - __shfl_sync(0xffffffffu, e_next, k)) << 23);
          if (lane < TC_CL) {
            const uint32_t ra = tc_mapa(qmax + (uint32_t)(rank * TC_ROWS + 8 * uw) * 4, lane);
            const uint32_t rb = tc_mapa(cbar + 8 * B_RECV, lane);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
              asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                               ra + 16 * hh),
                           "r"(__float_as_uint(val[4 * hh])), "r"(__float_as_uint(val[4 * hh + 1])),
                           "r"(__float_as_uint(val[4 * hh + 2])), "r"(__float_as_uint(val[4 * hh + 3])), "r"(rb)
                           : "memory");
          }
        }
        // ---- partial D^T = V0[128 neurons][K quarter] dI_{t+1}[rows][K quarter]^T: this warp reads neurons 32 q + lane
        // (= neuron `lane` of cluster rank q), rows 32 half .. 32 half + 31, and sends them to their owner
        if (dbg_on) p.dbg[t * 8 + 7] = clock64();
        if (WCK) {
          // ---- checkpoint adaptation tape, while the tensor core multiplies: w_t from w_{t+1} by solving
          // w_{t+1} = beta w_t + a u_t + b s_t (snns.py:718) -- or from the tape where step t closes a chunk -- and the
          // term of d(beta) that step t+1 left open, dw_{t+1} w_t.  beta >= 0.967: a rounding error grows by <= 1.034
          // per step back and is dropped at the next checkpoint; w only enters d(beta).  Everything it needs (u_t, the
          // adjoint of step t+1) is in registers, w itself waits in shared memory.  Placed HERE because no global load
          // is in flight: behind load_tape() the shared-memory loads share scoreboards with the 24 tape loads and wait
          // for DRAM (measured: +1.4 k cycles per step).
          const bool from_tape = tm == p.w_every - 1;      // (t + 1) % C == 0
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float s_t = spike_of(__fsub_rn(ut[k], p.theta));
            const float w_t = from_tape ? wc[256 * (8 + k)] : (wc[256 * k] - prm.a * ut[k] - prm.b * s_t) * inv_beta;
            pb += dw[k] * w_t;
            wc[256 * k] = w_t;
          }
        }
        mbar_wait_sleep(cbar + 8 * B_ACC_FULL, step & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_on) p.dbg[t * 8 + 3] = clock64();
        {
          uint32_t x[32];
          const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * half);
          TC_LD32(taddr, x);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // rows 32 half + 4 jj .. + 3 of neuron `lane` -> rank q's receive buffer, laid out [source rank][row quad][neuron]
          // [4 rows]: the 32 lanes of a store cover 512 contiguous bytes (DSMEM moves contiguous bytes as whole packets;
          // a layout with one 16-byte piece per lane and 256-byte stride took twice as long)
          const uint32_t la = recv + (uint32_t)((rank * 16 + 8 * half) * TC_COLS + lane) * 16;
          const uint32_t ra = tc_mapa(la, q), rb = tc_mapa(cbar + 8 * B_RECV, q);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)  // each store reports its 16 bytes to the owner's barrier: no fence, no arrival
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                             ra + (uint32_t)(jj * TC_COLS * 16)),
                         "r"(x[4 * jj]), "r"(x[4 * jj + 1]), "r"(x[4 * jj + 2]), "r"(x[4 * jj + 3]), "r"(rb)
                         : "memory");
        }
        if (dbg_on) p.dbg[t * 8 + 4] = clock64();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        // the tape values travel while the partial products cross the cluster (issued earlier they queue behind the
        // 64 KB of panel words in the SM's load path and delay this warp's arrival at the accumulator by ~2 k cycles)
        load_tape();
        // ---- this CTA's neurons from the four ranks, summed in rank order; the quarter maxima with them
        mbar_wait_cluster(cbar + 8 * B_RECV, step & 1);
        if ((uw & 3) == 0 && lane == 0 && t > 0) mbar_expect_tx(cbar + 8 * B_RECV, TC_RECV_TX);  // arm the next phase
        if (dbg_on) p.dbg[t * 8 + 5] = clock64();
        float acc[8];
#pragma unroll
        for (int s = 0; s < TC_CL; ++s) {
          const unsigned char* rp = recv_p + (size_t)((s * 16 + 2 * uw) * TC_COLS + lane) * 16;
          const float4 x0 = *reinterpret_cast<const float4*>(rp);
          const float4 x1 = *reinterpret_cast<const float4*>(rp + TC_COLS * 16);
          const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = s == 0 ? xs[k] : acc[k] + xs[k];
          m_next_l = fmaxf(m_next_l, qmax_p[s * TC_ROWS + 8 * uw + (lane & 7)]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int e_row = __shfl_sync(0xffffffffu, e_next, k);
          recb[k] = acc[k] * (__uint_as_float((uint32_t)(127 - e_row) << 23) * rs);
        }
        // the receive buffers are free for the next step once every update warp of the cluster has read them.  A relaxed
        // arrival: it signals completed shared-memory READS (their values are consumed above), there is nothing to
        // release -- mbarrier.arrive.release.cluster compiles to MEMBAR.ALL.GPU per arrival.
        __syncwarp();
        if (lane == 0 && t > 0) {
#pragma unroll
          for (int d = 0; d < TC_CL; ++d)
            asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tc_mapa(cbar + 8 * B_FREE, d)) : "memory");
        }
        if (dbg_on) p.dbg[(p.T + t) * 8 + 2] = clock64();
        ++step;
      } else {
        load_tape();
      }
      // ---- scale of this step's hand-over (lane k: row k): s_t = 2^e_cur brings max(rowmax|dI_{t+1}|, rowmax|g_t| / 4)
      // into [8, 16)
      const int e_cur = scale_exp(fmaxf(m_next_l, myrow_live ? 0.25f * grow_l : 0.f));
      // ---- BPTT update of the 8 (row, neuron) pairs (cell_math.cuh: step_bwd, same operations in the same order), in
      // two parts: first what dI_t needs, then its hand-over, and only then what nobody waits for (parameter-gradient
      // sums, the adjoint of w, the dI tape) -- that part runs while the published words travel to the L2.
      const uint32_t tag = tc_tag(p.T, t) << 16;
      uint32_t* dst = pan_wr0 + (size_t)tc_buf(p.T, t) * pan_buf;
      float d[8], sc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sc[k] = __uint_as_float((uint32_t)(127 + __shfl_sync(0xffffffffu, e_cur, k)) << 23);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool live = col_live && rowb + k < p.Be;
        gq[k] = live ? gq[k] : 0.f;
        up[k] = live ? up[k] : 0.f;
        wp[k] = live ? wp[k] : 0.f;
        float ds = gq[k] - prm.alpha * du[k] + recb[k];
        if (ADAPT) ds += prm.b * dw[k];
        float du_t = (window_of(__fsub_rn(ut[k], p.theta)) ? ds : 0.0f) + prm.alpha * du[k];
        if (ADAPT) du_t += prm.a * dw[k];
        du[k] = du_t;
        d[k] = prm.oma * du_t;
      }
      if (t > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float x = d[k] * sc[k];
          const __half hh = __float2half_rn(x);
          const __half hl = p.reduced ? __float2half_rn(0.f) : __float2half_rn(x - __half2float(hh));
          const uint32_t word = (uint32_t)__half_as_ushort(hh) | (((uint32_t)__half_as_ushort(hl) << 16) & 0xFFFE0000u) | tag;
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(dst), "r"(word) : "memory");
          dst += p.Hp;
        }
      }
      e_next = e_cur;
      if (dbg_on) p.dbg[t * 8 + 6] = clock64();
      float sp[8];
      if (t > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) sp[k] = spike_of(__fsub_rn(up[k], p.theta));
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sp[k] = (col_live && rowb + k < p.Be) ? p.s0[(int64_t)(rowb + k) * p.H + col] : 0.f;
      }
      if (WCK && t > 0 && tm == 0) {
        // step t-1 closes a chunk: its w came from the checkpoint tape (load_tape); the next step's waiting window
        // takes it from shared memory instead of recomputing it
#pragma unroll
        for (int k = 0; k < 8; ++k) wc[256 * (8 + k)] = wp[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float dd = up[k] - sp[k];
        pa += du[k] * ((dd - ut[k]) * inv_oma);
        if (ADAPT) {
          const float dw_t = prm.beta * dw[k] - d[k];
          if (!WCK || t == 0) pb += dw_t * wp[k];   // (checkpoint tape: added a step later, in the waiting window)
          pc += dw_t * up[k];
          pd += dw_t * sp[k];
          dw[k] = dw_t;
        }
        ut[k] = up[k];
        if (col_live && rowb + k < p.Be) __stcg(p.dI + toff + (rowb + k) * TH, d[k]);
      }
      if (dbg_on) p.dbg[(p.T + t) * 8 + 3] = clock64();
      if (WCK) {
        if (tm == 0) { tm = p.w_every - 1; --tq; } else { --tm; }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int row = rowb + k;
      if (col_live && row < p.Be) {
        // the sums cover the thread's 8 rows: they go to the first one, the buffers are reduced over the batch anyway
        const int64_t o1 = (int64_t)row * p.H + col;
        p.p_alpha[o1] = k == 0 ? pa : 0.f;
        if (ADAPT) {
          p.p_beta[o1] = k == 0 ? pb : 0.f;
          p.p_a[o1] = k == 0 ? pc : 0.f;
          p.p_b[o1] = k == 0 ? pd : 0.f;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  // cluster-wide: no CTA leaves while a peer may still store into its receive buffer
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static size_t rec_bwd_tc_smem(int KB, bool wck = true) {
  return (size_t)KB * TC_STAGE_BYTES + TC_RECV_BYTES + TC_QMAX_BYTES + 256 + 1024 + (wck ? 16 * 256 * sizeof(float) : 0);
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
  return TC_NBUF * groups * TC_ROWS * Hp * sizeof(uint32_t)  // rotating panels of {hi, lo | tag} words
         + (size_t)Be * T * sizeof(float)              // gmax
         + 256;
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
  return sparch_recur_bwd_tc_ck(kind, G, U, W, alpha, beta, a, b, img, meta, u0, w0, s0, theta, dI, p_alpha, p_beta, p_a,
                                p_b, workspace, reduced, Be, T, H, gmax_in, 0, st_);
}

int sparch_recur_bwd_tc_ck(int kind, const float* G, const float* U, const float* W, const float* alpha,
                           const float* beta, const float* a, const float* b, const void* img, const int* meta,
                           const float* u0, const float* w0, const float* s0, float theta, float* dI, float* p_alpha,
                           float* p_beta, float* p_a, float* p_b, void* workspace, int reduced, int Be, int T, int H,
                           const float* gmax_in, int w_every, sparch_stream_t st_) {
  SPARCH_REQUIRE(w_every >= 0, "w_every");
  SPARCH_REQUIRE(kind == SPARCH_RLIF || kind == SPARCH_RADLIF, "recurrent kinds only");
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0, "bad shape");
  if (Be == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && U && alpha && img && meta && u0 && s0 && dI && p_alpha && workspace, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (W && beta && a && b && w0 && p_beta && p_a && p_b),
                 "adaptive kind needs W, beta, a, b, w0 and the partial buffers");
  const int Hp = sparch_recur_tc_padded(H), KB = Hp / (64 * TC_CL);
  const size_t smem = rec_bwd_tc_smem(KB, adapt && w_every > 0);
  SPARCH_REQUIRE(KB <= 4 && smem <= 128 * 1024, "hidden size too large for the resident V0 tiles");
  SPARCH_REQUIRE((long long)Be * T * H < (1LL << 31), "tape larger than 2^31 elements: split the batch");
  cudaStream_t st = as_stream(st_);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    // 98.5 KB at H = 1024 (V0 lives in tensor memory): the rest of the SM's 256 KB stays L1, whose lines buffer the
    // 64 KB of panel words a CTA has in flight per step
    const void* fns[3] = {(const void*)rec_bwd_tc_kernel<true, false>, (const void*)rec_bwd_tc_kernel<true, true>,
                          (const void*)rec_bwd_tc_kernel<false, false>};
    for (const void* f : fns) {
      SPARCH_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      SPARCH_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    }
  }
  const int groups = (Be + TC_ROWS - 1) / TC_ROWS, slices = Hp / TC_COLS;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t panel_bytes = (size_t)TC_NBUF * groups * TC_ROWS * Hp * sizeof(uint32_t);
  uint32_t* panel = reinterpret_cast<uint32_t*>(ws);
  float* gmax = reinterpret_cast<float*>(ws + panel_bytes);
  SPARCH_CUDA(cudaMemsetAsync(ws, 0, panel_bytes, st));  // tag 0 everywhere: the first use of a buffer writes tag 1
  if (gmax_in) {  // the producer of G already left the row maxima (sparch_spike_post_bwd)
    gmax = const_cast<float*>(gmax_in);
  } else {
    const int64_t warps = (int64_t)Be * T;
    gmax_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(G, Be, T, H, gmax);
    SPARCH_LAUNCH_OK();
  }
  RecBwdTcArgs p{G, U, W, alpha, beta, a, b, u0, w0, s0, reinterpret_cast<const uint32_t*>(img), meta, gmax, theta,
                 dI, p_alpha, p_beta, p_a, p_b, panel, Be, T, H, Hp, KB, reduced ? 1 : 0, (kind & 1) ? w_every : 0, recur_debug_buffer(),
                 recur_debug_flags()};
  const bool wck = adapt && w_every > 0;
  const void* fn = wck     ? (const void*)rec_bwd_tc_kernel<true, true>
                   : adapt ? (const void*)rec_bwd_tc_kernel<true, false>
                           : (const void*)rec_bwd_tc_kernel<false, false>;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = TC_CL;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attrs;
  // The CTAs of a launch wait on each other, so all of them must be resident.  The grid of one launch never exceeds
  // what cudaOccupancyMaxActiveClusters reports for an empty GPU, and whatever else occupies SMs when it starts
  // (the tail of the previous kernel, a collective on another stream) finishes without needing this kernel: the
  // launch needs no cooperative attribute (Nsight Compute cannot replay a launch that is both clustered and
  // cooperative).  SPARCH_B200_TC_COOP=1 adds the attribute (gang scheduling by the driver).
  static const bool coop = getenv("SPARCH_B200_TC_COOP") && getenv("SPARCH_B200_TC_COOP")[0] == '1';
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
    void* args[] = {(void*)&p, (void*)&group0, (void*)&ngt};
    SPARCH_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  }
  return SPARCH_OK;
}

}  // extern "C"

// Forward recurrence of the recurrent kinds (RLIF, RadLIF: snns.py:554-578, 696-727) with s_{t-1} @ V0 on
// tcgen05: integer tensor cores, the spike operand in TENSOR MEMORY, one persistent kernel for all T steps.
//
//   CTA (row group g, slice s) = 128 batch rows x 16 neurons, resident for the whole layer pass.
//   B operand: its 16 columns of V0 as THREE int8 digit planes of a 23-bit fixed-point image
//              q = rint(V0[j][c] * 2^(22 - E_c)), q = d2 2^16 + d1 2^8 + d0 (balanced digits, E_c per column),
//              stacked along N (N = 48: one UMMA serves all three planes), K-major SWIZZLE_128B tiles resident in
//              shared memory (Hp x 48 bytes = 48 KB at H = 1024).
//   A operand: the spikes of step t-1 of the group's 128 rows as uint8 {0,1}, written into TMEM by tcgen05.st
//              (TMEM lane = batch row, 32-bit column = 4 K positions); tcgen05.mma.kind::i8 reads A from TMEM, so
//              the spike matrix never exists in shared or global memory -- it travels as bits.
//   D:         int32 128 x 48 in TMEM (two buffers, alternating steps): EXACT products and sums; the three planes
//              are recombined in fp32 with one rounding -- s @ V0 is as accurate as an fp32 matrix product.
//   Exchange:  each step a CTA publishes, per batch row, one 32-bit word holding its 16 spikes as 2-BIT FIELDS
//              (spike i at bit 2i; written as two 16-bit halves by the two warps that own 8 neurons each).  The odd
//              bits of a published half are zero, a buffer memset to 0xFF per step means "not yet written": the
//              consumers of the row group poll the words directly (ld.relaxed.gpu) -- one L2 round trip, no fence, no
//              flag.  The 2-bit fields are prmt selectors as they are: a nibble {0, s_odd, 0, s_even} picks byte
//              s_even out of (00 01 .. 00 01) or s_odd out of (00 00 .. 01 01), so 16 spikes become 4 TMEM columns in
//              5 integer instructions (integer ALU throughput is what bounds the expansion: every CTA of a row group
//              expands the same 128 x H spikes).  The published words ARE the packed spike tensor
//              [T][group][slice][128 rows].
//   Warps:     0-7 workers, 8 MMA issue.  Worker w = (TMEM lane quarter w % 4, half w / 4) alternates two jobs per
//              step: EXPANSION of its half of the 256-K batches (poll 16 producers' words -> 64 TMEM columns -> arrive
//              on the batch's mbarrier; its first batch is polled alone and the next is fetched while the first is
//              expanded and multiplied), then the NEURON UPDATE of its 32 rows x 8 neurons (tcgen05.ld D, state u, w, s
//              in registers for all T, publish, tapes).  The MMA warp issues 8 UMMAs 128 x 48 x 32 per batch as a
//              converged warp (elect.sync).
// Per-step chain: publish -> L2 -> poll -> expand -> UMMA -> tcgen05.ld -> update -> publish.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "cell_math.cuh"
#include "common.cuh"
#include "tcgen05_utils.cuh"

namespace sparch {

constexpr int FT_ROWS = 128;          // batch rows per CTA = TMEM lanes
constexpr int FT_NEUR = 16;           // neurons per CTA
constexpr int FT_N = 48;              // UMMA N: 3 digit planes x 16 neurons
constexpr int FT_BATCH = 16;          // producer slices per expansion batch: 256 K = 64 TMEM columns = 8 UMMAs
constexpr int FT_MAXB = 6;            // batches per step: H <= 1536 (A columns 0..383, D at 384 and 448)
constexpr int FT_DCOL0 = 384, FT_DCOL1 = 448;
constexpr int FT_WORKERS = 16;          // worker warps: 4 TMEM lane quarters x 4 (batch residue / neuron quad)
constexpr int FT_THREADS = (FT_WORKERS + 1) * 32;
constexpr int FT_MMA_WARP = FT_WORKERS;
constexpr int FT_QBITS = 22;
constexpr uint32_t FT_ODD = 0xAAAAAAAAu;   // bits that are zero in a published word, one in the 0xFF fill
constexpr long long FT_SPIN_LIMIT = 4000000000LL;

// K position kappa as the tensor core sees it (TMEM column kappa/4, byte kappa%4) -> presynaptic neuron.  Inside a
// producer's group of 16: kappa = 4 c + e (column c of the four a word expands to, byte e)  <->  neuron
// 8 (c >> 1) + 2 e + (c & 1): columns 0/1 hold the even/odd spikes of the word's low half, 2/3 of its high half.
__host__ __device__ __forceinline__ int ft_neuron_of(int kappa) {
  const int r = kappa & 15, c = r >> 2, e = r & 3;
  return (kappa & ~15) + 8 * (c >> 1) + 2 * e + (c & 1);
}

__device__ __forceinline__ int ft_exp_of(float m) {  // m < 2^e
  int e = 0;
  if (m > 0.f && m <= 3.0e38f) frexpf(m, &e);
  return e;
}

// colmax[c] = bit pattern of max_j |V[j][c]|, j != c (zeroed by the caller)
__global__ void ft_colmax_kernel(const float* __restrict__ V, int H, unsigned* __restrict__ colmax) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (c >= H) return;
  float m = 0.f;
  for (int r = blockIdx.y * 8 + threadIdx.y; r < H; r += gridDim.y * 8)
    if (r != c) m = fmaxf(m, fabsf(V[(int64_t)r * H + c]));
  atomicMax(&colmax[c], __float_as_uint(m));
}

// img: [slice][kb < Hp/128][n < 48][128 bytes], byte (n, k) of a tile at n*128 + (((k/16) ^ (n&7)) * 16) + k%16;
// row n = plane * 16 + local neuron (plane 0 = most significant digit).  One thread per 16-byte swizzle chunk.
__global__ void ft_vprep_kernel(const float* __restrict__ V, int H, int Hp, int nslices,
                                const unsigned* __restrict__ colmax, uint8_t* __restrict__ img,
                                float* __restrict__ colscale) {
  const int KB = Hp / 128;
  const int64_t total = (int64_t)nslices * KB * FT_N * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int chunk_sw = (int)(i & 7);
    const int n = (int)((i >> 3) % FT_N);
    const int64_t rest = (i >> 3) / FT_N;
    const int kb = (int)(rest % KB), slice = (int)(rest / KB);
    const int plane = n / FT_NEUR, c = slice * FT_NEUR + n % FT_NEUR;
    const int kappa0 = kb * 128 + ((chunk_sw ^ (n & 7)) << 4);
    const int ec = c < H ? ft_exp_of(__uint_as_float(colmax[c])) : 0;
    const float sc = ldexpf(1.0f, FT_QBITS - ec);
    __align__(16) uint8_t out[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int nu = ft_neuron_of(kappa0 + e);
      float x = 0.f;
      if (nu < H && c < H && nu != c) x = V[(int64_t)nu * H + c] * sc;
      const int q = __float2int_rn(x);
      const int d0 = (int)(int8_t)(q & 0xff);
      const int q1 = (q - d0) >> 8;
      const int d1 = (int)(int8_t)(q1 & 0xff);
      const int d2 = (q1 - d1) >> 8;
      out[e] = (uint8_t)(plane == 0 ? d2 : plane == 1 ? d1 : d0);
    }
    *reinterpret_cast<uint4*>(img + (((int64_t)slice * KB + kb) * FT_N + n) * 128 + chunk_sw * 16) =
        *reinterpret_cast<const uint4*>(out);
    if (plane == 0 && kb == 0 && chunk_sw == 0) colscale[c] = ldexpf(1.0f, ec - FT_QBITS);
  }
}

struct RecFwdTcArgs {
  const float *Z, *scale, *shift, *alpha, *beta, *a, *b, *rec0, *u0, *w0, *s0;
  const uint8_t* img;
  const float* colscale;
  float theta;
  float *S, *U, *W;
  uint32_t* bits;  // [T][groups][slices][128]
  int Be, T, H, Hp, NB, ngroups_total;
  int reduced;     // 1: two digit planes only (15-bit image of V0)
  int use_tma;     // 1: Z tiles in / S, U, W tiles out by TMA (needs H % 4 == 0); 0: per-thread global accesses
  int w_every;     // 0: W is the full (Be, T, H) adaptation tape.  C > 0: W is (Be, ceil(T / C), H) and receives w_t only at
                   // the last step of every chunk of C steps (and at T - 1): the reverse pass recomputes the steps in
                   // between from u and s by inverting the update (csrc/recur_tc.cu), 4 / C instead of 4 B/elt each way
  int rev_from;    // bidirectional layers (snns.py:666-668): rows >= rev_from run the sequence backwards and read the
                   // input of row - rev_from -- Z is (rev_from, T, H), no flipped copy; 0 = off.  TMA mode only.
  long long* dbg;  // optional [2T][8] phase clocks of CTA (0,0), normally NULL
  int dbg_flags;
};

struct FtMaps {
  CUtensorMap z, s, u, w;  // (Be, T, H) fp32 tensors through boxes of 32 rows x 1 step x 16 neurons
};

constexpr int FT_TILE_BYTES = FT_ROWS * FT_NEUR * 4;  // one 128 x 16 fp32 tile
constexpr int FT_TMA_SMEM = 2 * FT_TILE_BYTES + 2 * 3 * FT_TILE_BYTES;  // Z double buffer + (S, U, W) double buffer

#define FT_ST32(taddr, v, o)                                                                                             \
  asm volatile(                                                                                                          \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),                                               \
      "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]),            \
      "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]),        \
      "r"(v[o + 14]), "r"(v[o + 15]), "r"(v[o + 16]), "r"(v[o + 17]), "r"(v[o + 18]), "r"(v[o + 19]), "r"(v[o + 20]),     \
      "r"(v[o + 21]), "r"(v[o + 22]), "r"(v[o + 23]), "r"(v[o + 24]), "r"(v[o + 25]), "r"(v[o + 26]), "r"(v[o + 27]),     \
      "r"(v[o + 28]), "r"(v[o + 29]), "r"(v[o + 30]), "r"(v[o + 31])                                                      \
      : "memory")

#define FT_LD4(taddr, v)                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"                 \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])                             \
               : "r"(taddr))

#define FT_LD8(taddr, v)                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                     \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
               : "r"(taddr))

// The MMA warp stays converged; elect.sync picks the lane that issues a batch's UMMAs.  (Measured with
// tools/ubench/umma_i8_ts.cu: under a plain divergent `if (lane == 0)` ptxas wraps every UMMA in an ELECT / PLOP3 /
// BRA.U.ANY retry loop and a K = 1024 step of 32 UMMAs takes 1440 cycles whatever N is; with the election visible
// to the compiler it takes the tensor pipe's 128 * N / 256 cycles per UMMA -- 773 cycles at N = 48.)
__device__ __forceinline__ bool elect_one() {
  uint32_t e;
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(e));
  return e != 0;
}
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ bool ft_valid(uint32_t w) { return (w & FT_ODD) == 0u; }

// exact int32 -> fp32 for |x| < 2^22 without the conversion pipe
__device__ __forceinline__ float ft_i2f(uint32_t x) { return __int_as_float((int)x + 0x4B400000) - 12582912.0f; }

template <bool ADAPT>
__global__ void __launch_bounds__(FT_THREADS, 1)
rec_fwd_tc_kernel(const __grid_constant__ FtMaps maps, const RecFwdTcArgs p, const int group0) {
  extern __shared__ unsigned char fsm_raw[];
  const uint32_t raw = smem_u32(fsm_raw);
  const uint32_t vimg = (raw + 1023u) & ~1023u;
  unsigned char* vsm = fsm_raw + (vimg - raw);
  __shared__ float4 sprm_a[FT_NEUR];  // alpha, 1 - alpha, beta, a
  __shared__ float4 sprm_b[FT_NEUR];  // b, BatchNorm scale, shift, column scale of V0's image
  __shared__ __align__(8) unsigned long long sbar[1 + 2 + 8];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x, nsl = gridDim.x, group = group0 + blockIdx.y, row0 = group * FT_ROWS;
  const int NB = p.NB, rot = slice % NB;
  const uint32_t bar_a = smem_u32(&sbar[0]);             // a_ready: the whole A operand of a step is in TMEM (4 NB warp arrivals)
  const uint32_t bar_acc = smem_u32(&sbar[1]);           // acc_full[d]: D buffer d complete (tcgen05.commit)
  const uint32_t bar_z = smem_u32(&sbar[3]);             // z_full[quarter][b]: the quarter's Z tile of a step has landed
  const uint32_t ztile = vimg + (uint32_t)p.Hp * FT_N;   // 2 x 8 KB: Z tiles [128 rows][16 neurons] of steps t, t+1
  const uint32_t stile = ztile + 2 * FT_TILE_BYTES;      // 2 x (S, U, W) tiles, the same layout
  const bool dbg_cta = p.dbg && blockIdx.x == 0 && blockIdx.y == 0;

  {  // resident digit planes of this slice's 16 columns of V0
    const size_t v_bytes = (size_t)p.Hp * FT_N;
    const uint4* src = reinterpret_cast<const uint4*>(p.img + (size_t)slice * v_bytes);
    uint4* dst = reinterpret_cast<uint4*>(vsm);
    for (int i = tid; i < (int)(v_bytes / 16); i += FT_THREADS) {
      const uint32_t sa = smem_u32(dst + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i));
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
  }
  if (tid < FT_NEUR) {
    const int colr = slice * FT_NEUR + tid, col = min(colr, p.H - 1);
    const NeuronParams q0 = load_params<ADAPT>(p.alpha, p.beta, p.a, p.b, col);
    sprm_a[tid] = make_float4(q0.alpha, q0.oma, q0.beta, q0.a);
    sprm_b[tid] = make_float4(q0.b, p.scale ? p.scale[col] : 1.0f, p.scale ? p.shift[col] : 0.0f,
                              colr < p.H ? p.colscale[colr] : 0.0f);
  }
  if (tid == 0) {
    mbar_init(bar_a, 4 * NB);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc + 8, 1);
    for (int b = 0; b < 8; ++b) mbar_init(bar_z + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == FT_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async-written tiles -> tensor core reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == FT_MMA_WARP) {
    // ===== MMA issue (whole warp, converged; one elected lane per instruction) =====
    // kind::i8: D = S32, A = U8 (TMEM), B = S8 (smem, K-major), M = 128
    const uint32_t n_mma = p.reduced ? 32u : (uint32_t)FT_N;
    const uint32_t idesc = (2u << 4) | (1u << 10) | ((n_mma >> 3) << 17) | ((uint32_t)(FT_ROWS >> 4) << 24);
    const uint64_t desc0 = make_desc_k_sw128(vimg);
    for (int t = 1; t < p.T; ++t) {
      const int k = t - 1;
      const uint32_t dcol = tmem + ((k & 1) ? FT_DCOL1 : FT_DCOL0);
      // ONE wait per step: the 16 workers expand their batches in parallel and finish within a few hundred cycles of
      // each other, so waiting per batch only added the wait loop's ~400 cycles per batch to the chain
      mbar_wait_sleep(bar_a, k & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (dbg_cta && lane == 0) p.dbg[t * 8 + 2] = clock64();
      __syncwarp();
      // One election per step, then the UMMAs as straight-line code of the elected lane: operand addresses are
      // (batch base + compile-time constant), 3 uniform adds + 1 UTCIMMA per UMMA in SASS.  (A per-instruction
      // elect.sync cost ~19 instructions per UMMA.)
      if (elect_one()) {
        for (int bt = 0; bt < NB; ++bt) {
          const uint32_t a_base = tmem + (uint32_t)(64 * bt);
          const uint64_t d_base = desc0 + (uint64_t)(bt * (2 * FT_N * 128 / 16));   // two 128-K tiles per batch
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t bdesc = d_base + (uint64_t)((ks >> 2) * (FT_N * 128 / 16) + 2 * (ks & 3));
            if (!(p.dbg_flags & 16)) umma_i8_ts(dcol, a_base + 8 * ks, bdesc, idesc, (bt > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_acc + 8 * (k & 1));
      }
      __syncwarp();
      if (dbg_cta && lane == 0) p.dbg[t * 8 + 3] = clock64();
    }
  } else {
    // ===== workers: expansion of the batches i = r, r + 4 (rotated order), then the update of 32 rows x 4 neurons =====
    const int q = warp & 3, r = warp >> 2;
    const int lrow = 32 * q + lane, row = row0 + lrow;
    const bool rlive = row < p.Be;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    const bool dbg_w = dbg_cta && tid == 0;

    // ---------- expansion helpers ----------
    // 16 words of batch bt (one per producer slice) for this thread's row; rows beyond Be and slices beyond the last
    // read as "published, no spikes"
    auto load_batch = [&](const uint32_t* base, int bt, uint32_t (&wv)[FT_BATCH]) {
      const uint32_t* src = base + (size_t)(bt * FT_BATCH) * FT_ROWS;
      const int nq = rlive ? min(FT_BATCH, nsl - bt * FT_BATCH) : 0;   // producers of this batch that exist
#pragma unroll
      for (int i = 0; i < FT_BATCH; ++i) wv[i] = 0u;
      if (nq == FT_BATCH) {
#pragma unroll
        for (int i = 0; i < FT_BATCH; ++i)
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(wv[i]) : "l"(src + (size_t)i * FT_ROWS) : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < FT_BATCH; ++i)
          if (i < nq)
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(wv[i]) : "l"(src + (size_t)i * FT_ROWS) : "memory");
      }
    };
    // re-poll until every word of the batch is there (whole warp: the TMEM store that follows is warp-wide)
    auto ensure = [&](const uint32_t* base, int bt, uint32_t (&wv)[FT_BATCH]) {
      const long long t0 = clock64();
      for (;;) {
        uint32_t any = 0u;
#pragma unroll
        for (int i = 0; i < FT_BATCH; ++i) any |= wv[i];
        const bool ok = ft_valid(any) || (p.dbg_flags & 1);
        if (__all_sync(0xffffffffu, ok)) break;
        if (!ok) {  // spin on the first word that is still missing (one light load per round), then re-read the batch
          int bad = 0;
#pragma unroll
          for (int i = FT_BATCH - 1; i >= 0; --i) bad = ft_valid(wv[i]) ? bad : i;
          const uint32_t* a = base + (size_t)(bt * FT_BATCH + bad) * FT_ROWS;
          uint32_t x;
          do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(a) : "memory");
            if (clock64() - t0 > FT_SPIN_LIMIT) __trap();  // a lost store must not hang the GPU
          } while (!ft_valid(x));
        }
        __syncwarp();
        load_batch(base, bt, wv);
      }
    };
    // 16 producers x 16 spikes -> 64 columns of 4 bytes: per word two prmt on its low half (even / odd spikes), a
    // shift, two prmt on its high half; the word's nibbles {0, s_odd, 0, s_even} are the selectors
    auto expand_store_signal = [&](int bt, const uint32_t (&wv)[FT_BATCH]) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t x = wv[8 * half + i], xh = x >> 16;
          v[4 * i + 0] = __byte_perm(0x00000100u, 0x00000100u, x);
          v[4 * i + 1] = __byte_perm(0x00000000u, 0x00000101u, x);
          v[4 * i + 2] = __byte_perm(0x00000100u, 0x00000100u, xh);
          v[4 * i + 3] = __byte_perm(0x00000000u, 0x00000101u, xh);
        }
        FT_ST32(lane_addr + (uint32_t)(64 * bt + 32 * half), v, 0);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
    };
    auto batch_of = [&](int i) {
      const int bt = rot + i;
      return bt >= NB ? bt - NB : bt;
    };

    // ---------- update state: neurons 4r .. 4r+3 of this thread's row ----------
    const int nl0 = 4 * r, col0 = slice * FT_NEUR + nl0;
    const int nv = rlive ? max(0, min(4, p.H - col0)) : 0;
    const bool vec = ((p.H & 3) == 0) && nv == 4;
    const uint32_t live_mask = ((1u << (2 * nv)) - 1u) & 0x55u;
    const int64_t idx0 = (int64_t)row * p.H + col0;
    float u[4], w[4], s[4], zn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = w[i] = s[i] = zn[i] = 0.f;
    auto ld4 = [&](const float* src, float (&v)[4]) {
      if (vec) {
        const float4 x0 = *reinterpret_cast<const float4*>(src);
        v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = i < nv ? src[i] : 0.f;
      }
    };
    auto st4 = [&](float* dst, const float (&v)[4]) {
      if (vec) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nv) dst[i] = v[i];
      }
    };
    const bool tma = p.use_tma != 0;
    if (nv > 0) {
      ld4(p.u0 + idx0, u);
      ld4(p.s0 + idx0, s);
      if (ADAPT) ld4(p.w0 + idx0, w);
      if (p.T > 0 && !tma) ld4(p.Z + (int64_t)row * p.T * p.H + col0, zn);
    }
    // TMA mode: the Z tile of step t+1 (128 rows x 16 neurons) is fetched by one bulk-tensor copy while step t runs,
    // and the S / U / W tiles of step t leave through shared memory by bulk-tensor stores: per-thread accesses of a
    // (row, 4 neurons) thread would touch 32 different 128-byte lines per warp instruction, and 4 such accesses per
    // step and thread kept the load/store unit busy for ~2000 cycles per step.
    const uint32_t tile_off = (uint32_t)(lrow * (FT_NEUR * 4) + r * 16);
    const bool issuer = tma && r == 0 && lane == 0;   // one thread per quarter drives its 32-row tiles
    const uint32_t qz = bar_z + 16 * q;               // z_full[q][2]
    const uint32_t qoff = (uint32_t)(32 * q) * (FT_NEUR * 4);
    // bidirectional second half: this CTA's rows read Z[row - rev_from][T - 1 - t] (a row group lies in one half)
    const bool rev = p.rev_from > 0 && row0 >= p.rev_from;
    const int zrow0 = rev ? row0 - p.rev_from : row0;
    auto zt = [&](int t_) { return rev ? p.T - 1 - t_ : t_; };
    if (issuer) {
      for (int t0 = 0; t0 < 2 && t0 < p.T; ++t0) {
        mbar_expect_tx(qz + 8 * t0, FT_TILE_BYTES / 4);
        tma_load_3d(ztile + (uint32_t)t0 * FT_TILE_BYTES + qoff, &maps.z, slice * FT_NEUR, zt(t0), zrow0 + 32 * q, qz + 8 * t0);
      }
    }
    uint8_t* pub = reinterpret_cast<uint8_t*>(p.bits) + ((((size_t)group) * nsl + slice) * FT_ROWS + lrow) * 4 + r;
    const size_t pub_step = (size_t)p.ngroups_total * nsl * FT_ROWS * 4;

    // TMA mode: hand the tapes of step ts (still in this thread's registers) to the quarter's issuing thread, which
    // stores the three 32-row tiles and fetches the Z tile of step ts + 2.  Called right after the expansion of step
    // ts + 1, i.e. while the tensor core works and the workers would only wait: the fence, the quarter's barrier and
    // the bulk copies stay off the publish -> poll chain (placed right after the publish they cost ~2000 cycles there).
    auto hand_over = [&](int ts) {
      const uint32_t sb = stile + (uint32_t)(ts & 1) * (3 * FT_TILE_BYTES) + tile_off;
      if (p.S)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb), "f"(s[0]), "f"(s[1]), "f"(s[2]), "f"(s[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + FT_TILE_BYTES), "f"(u[0]), "f"(u[1]), "f"(u[2]), "f"(u[3])
                   : "memory");
      const bool w_step = ADAPT && (!p.w_every || ts % p.w_every == p.w_every - 1 || ts == p.T - 1);
      if (w_step)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + 2 * FT_TILE_BYTES), "f"(w[0]), "f"(w[1]), "f"(w[2]),
                     "f"(w[3])
                     : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> the TMA's reads
      if (issuer) tma_store_wait_read0();                             // the stores of the step before have read their tiles
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");      // the quarter's four warps (32 rows x 16 neurons)
      if (issuer) {
        const uint32_t tb = stile + (uint32_t)(ts & 1) * (3 * FT_TILE_BYTES) + qoff;
        if (!(p.dbg_flags & 2)) {
          if (p.S) tma_store_3d(&maps.s, slice * FT_NEUR, ts, row0 + 32 * q, tb);
          tma_store_3d(&maps.u, slice * FT_NEUR, ts, row0 + 32 * q, tb + FT_TILE_BYTES);
          if (w_step) tma_store_3d(&maps.w, slice * FT_NEUR, p.w_every ? ts / p.w_every : ts, row0 + 32 * q, tb + 2 * FT_TILE_BYTES);
        }
        tma_store_commit();
        if (ts + 2 < p.T) {  // Z tile of step ts + 2: its buffer held step ts, read before this barrier by all four warps
          const uint32_t b = qz + 8 * (ts & 1);
          mbar_expect_tx(b, FT_TILE_BYTES / 4);
          tma_load_3d(ztile + (uint32_t)(ts & 1) * FT_TILE_BYTES + qoff, &maps.z, slice * FT_NEUR, zt(ts + 2), zrow0 + 32 * q, b);
        }
      }
    };
    uint32_t wa[FT_BATCH];   // the words of this worker's first batch: polled right after the publish of the step before
    for (int t = 0; t < p.T; ++t) {
      const int64_t o0 = ((int64_t)row * p.T + t) * p.H + col0;
      float z[4], rec[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { z[i] = zn[i]; rec[i] = 0.f; }
      if (!tma && nv > 0 && t + 1 < p.T) ld4(p.Z + o0 + p.H, zn);  // next step's input is in flight during this step
      if (t == 0) {
        if (nv > 0) ld4(p.rec0 + idx0, rec);
      } else {
        // ---- expansion: this worker's batches (i = r, r + 4) of the spike words of step t-1.  (The UMMAs of step
        // t-2 have read A: this thread saw their commit before its update of step t-1.)
        const uint32_t* base = p.bits + (((size_t)(t - 1) * p.ngroups_total + group) * nsl) * FT_ROWS + lrow;
        for (int i = r; i < NB; i += 4) {
          if (i > r) load_batch(base, batch_of(i), wa);
          ensure(base, batch_of(i), wa);
          if (dbg_w && i == r) p.dbg[t * 8 + 1] = clock64();
          expand_store_signal(batch_of(i), wa);
        }
        if (dbg_w) p.dbg[t * 8 + 6] = clock64();
        if (dbg_cta && lane == 0 && q == 0 && r > 0) p.dbg[(p.T + t) * 8 + 4 + r] = clock64();  // workers r = 1..3 done expanding
        if (tma) hand_over(t - 1);
      }
      if (tma) {
        mbar_wait_sleep(qz + 8 * (t & 1), (t >> 1) & 1);
        float4 zz;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(zz.x), "=f"(zz.y), "=f"(zz.z), "=f"(zz.w)
                     : "r"(ztile + (uint32_t)(t & 1) * FT_TILE_BYTES + tile_off));
        z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
      }
      // ---- everything of the update that does not need s_{t-1} @ V0 happens while the tensor core works: the
      // adaptation variable, alpha (u - s) and the normalised input (same operations and roundings as step_fwd)
      float au[4], oma[4], cs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 pa = sprm_a[nl0 + i], pb = sprm_b[nl0 + i];
        if (ADAPT) w[i] = __fadd_rn(__fadd_rn(__fmul_rn(pa.z, w[i]), __fmul_rn(pa.w, u[i])), __fmul_rn(pb.x, s[i]));
        au[i] = __fmul_rn(pa.x, __fsub_rn(u[i], s[i]));
        z[i] = __fmaf_rn(z[i], pb.y, pb.z);   // (scale, shift) = (1, 0) without normalisation: exact
        oma[i] = pa.y;
        cs[i] = pb.w;
      }
      if (t > 0) {
        // ---- D of this step
        const int k = t - 1;
        mbar_wait_sleep(bar_acc + 8 * (k & 1), (k >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (dbg_w) p.dbg[t * 8 + 4] = clock64();
        const uint32_t ta = lane_addr + ((k & 1) ? FT_DCOL1 : FT_DCOL0) + (uint32_t)nl0;
        uint32_t d2[4], d1[4], d0[4];
        FT_LD4(ta, d2);
        FT_LD4(ta + 16, d1);
        if (!p.reduced) FT_LD4(ta + 32, d0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo = p.reduced ? ft_i2f(d1[i]) * 256.0f : fmaf(ft_i2f(d1[i]), 256.0f, ft_i2f(d0[i]));
          rec[i] = fmaf(ft_i2f(d2[i]), 65536.0f, lo) * cs[i];
        }
      }
      // branch-free over the 4 neurons (their dependent chains interleave): dead neurons / rows compute on zeros and
      // are masked out of the published word and the tape stores
      uint32_t my = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x = __fadd_rn(z[i], rec[i]);
        if (ADAPT) x = __fsub_rn(x, w[i]);
        u[i] = __fadd_rn(au[i], __fmul_rn(oma[i], x));
        s[i] = spike_of(__fsub_rn(u[i], p.theta));
        my |= (s[i] > 0.f ? 1u : 0u) << (2 * i);   // 2-bit fields: the odd bits stay zero = "published"
      }
      my &= live_mask;
      if (rlive) {  // publish first: this store is what the other slices wait for (dead neurons publish zeros)
        const uint16_t b = (uint16_t)my;
        asm volatile("{\n.reg .b16 t;\nmov.b16 t, %1;\nst.relaxed.gpu.global.u8 [%0], t;\n}" ::"l"(pub + (size_t)t * pub_step), "h"(b)
                     : "memory");
      }
      if (dbg_w) p.dbg[t * 8 + 5] = clock64();
      if (!tma && nv > 0 && !(p.dbg_flags & 2)) {
        if (p.S) st4(p.S + o0, s);
        st4(p.U + o0, u);
        if (ADAPT) {
          if (!p.w_every)
            st4(p.W + o0, w);
          else if (t % p.w_every == p.w_every - 1 || t == p.T - 1)
            st4(p.W + ((int64_t)row * ((p.T + p.w_every - 1) / p.w_every) + t / p.w_every) * p.H + col0, w);
        }
      }
      // the first poll of the next step
      if (t + 1 < p.T && r < NB) {
        const uint32_t* base = p.bits + (((size_t)t * p.ngroups_total + group) * nsl) * FT_ROWS + lrow;
        if (dbg_w) p.dbg[(t + 1) * 8 + 0] = clock64();
        // pilot: one word per thread (a 128-byte line per warp and round) until the batch's first producer has
        // published, then the 16 words at once -- re-reading all 16 in every round made a round cost 1300-1900 cycles
        if (rlive && !(p.dbg_flags & 1)) {
          const uint32_t* a = base + (size_t)(batch_of(r) * FT_BATCH) * FT_ROWS;
          const long long t0 = clock64();
          uint32_t x;
          do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(a) : "memory");
            if (clock64() - t0 > FT_SPIN_LIMIT) __trap();
          } while (!ft_valid(x));
        }
        __syncwarp();
        load_batch(base, batch_of(r), wa);
      }
    }
    if (tma && p.T > 0) hand_over(p.T - 1);
    if (issuer) tma_store_wait_all();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == FT_MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace sparch

using namespace sparch;

extern "C" {

// number of 16-neuron slices / 256-K batches of a hidden size
static inline int ft_slices(int H) { return (H + FT_NEUR - 1) / FT_NEUR; }
static inline int ft_batches(int H) { return (ft_slices(H) + FT_BATCH - 1) / FT_BATCH; }

int sparch_recur_fwd_tc_max_h(void) { return FT_MAXB * FT_BATCH * FT_NEUR; }

size_t sparch_recur_fwd_tc_image_bytes(int H) {
  // digit-plane image [slice][Hp/128][48][128] + column scales (slices * 16 floats) + column maxima (same count)
  const size_t Hp = (size_t)ft_batches(H) * 256, ns = ft_slices(H);
  return ns * Hp * FT_N + 2 * ns * FT_NEUR * sizeof(float);
}

size_t sparch_recur_fwd_tc_bits_bytes(int Be, int T, int H) {
  const size_t groups = (size_t)(Be + FT_ROWS - 1) / FT_ROWS;
  return (size_t)T * groups * ft_slices(H) * FT_ROWS * sizeof(uint32_t);
}

int sparch_recur_prepare_fwd_tc(const float* V, int H, void* img, sparch_stream_t st_) {
  SPARCH_REQUIRE(V && H > 0 && img, "null pointer");
  SPARCH_REQUIRE(H <= sparch_recur_fwd_tc_max_h(), "hidden size too large for the tcgen05 forward recurrence");
  cudaStream_t st = as_stream(st_);
  const int ns = ft_slices(H), Hp = ft_batches(H) * 256;
  uint8_t* planes = reinterpret_cast<uint8_t*>(img);
  float* colscale = reinterpret_cast<float*>(planes + (size_t)ns * Hp * FT_N);
  unsigned* colmax = reinterpret_cast<unsigned*>(colscale + (size_t)ns * FT_NEUR);
  SPARCH_CUDA(cudaMemsetAsync(colscale, 0, 2 * (size_t)ns * FT_NEUR * sizeof(float), st));
  dim3 g((H + 31) / 32, 16);
  ft_colmax_kernel<<<g, dim3(32, 8), 0, st>>>(V, H, colmax);
  SPARCH_LAUNCH_OK();
  const int64_t total = (int64_t)ns * (Hp / 128) * FT_N * 8;
  int nb = (int)((total + 255) / 256);
  if (nb > sm_count() * 16) nb = sm_count() * 16;
  ft_vprep_kernel<<<nb, 256, 0, st>>>(V, H, Hp, ns, colmax, planes, colscale);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

int sparch_recur_fwd_tc(int kind, const float* Z, const float* scale, const float* shift, const float* alpha,
                        const float* beta, const float* a, const float* b, const float* rec0, const void* img,
                        const float* u0, const float* w0, const float* s0, float theta, float* S, float* U, float* W,
                        uint32_t* bits, int reduced, int Be, int T, int H, sparch_stream_t st_) {
  return sparch_recur_fwd_tc_bidir(kind, Z, scale, shift, alpha, beta, a, b, rec0, img, u0, w0, s0, theta, S, U, W, bits,
                                   reduced, Be, T, H, 0, 0, st_);
}

int sparch_recur_fwd_tc_bidir(int kind, const float* Z, const float* scale, const float* shift, const float* alpha,
                              const float* beta, const float* a, const float* b, const float* rec0, const void* img,
                              const float* u0, const float* w0, const float* s0, float theta, float* S, float* U,
                              float* W, uint32_t* bits, int reduced, int Be, int T, int H, int rev_from, int w_every,
                              sparch_stream_t st_) {
  SPARCH_REQUIRE(kind == SPARCH_RLIF || kind == SPARCH_RADLIF, "recurrent kinds only");
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0, "bad shape");
  SPARCH_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  SPARCH_REQUIRE(H <= sparch_recur_fwd_tc_max_h(), "hidden size too large for the tcgen05 forward recurrence");
  if (Be == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(Z && alpha && rec0 && img && u0 && s0 && U && bits, "null pointer");   // S may be NULL: bits only
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (beta && a && b && w0 && W), "adaptive kind needs beta, a, b, w0, W");
  const int ns = ft_slices(H), NB = ft_batches(H), Hp = NB * 256;
  const int groups = (Be + FT_ROWS - 1) / FT_ROWS;
  const size_t smem = (size_t)Hp * FT_N + FT_TMA_SMEM + 1024;
  cudaStream_t st = as_stream(st_);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    SPARCH_CUDA(cudaFuncSetAttribute(rec_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SPARCH_CUDA(cudaFuncSetAttribute(rec_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const int max_ctas = sm_count();  // one CTA per SM: each allocates all 512 TMEM columns
  SPARCH_REQUIRE(ns <= max_ctas, "hidden size needs more co-resident CTAs than the GPU has SMs");
  const uint8_t* planes = reinterpret_cast<const uint8_t*>(img);
  const float* colscale = reinterpret_cast<const float*>(planes + (size_t)ns * Hp * FT_N);
  static const bool tma_off = getenv("SPARCH_B200_FWD_TMA") && getenv("SPARCH_B200_FWD_TMA")[0] == '0';
  const int use_tma = ((H & 3) == 0 && !tma_off) ? 1 : 0;   // TMA needs 16-byte global strides
  SPARCH_REQUIRE(w_every >= 0, "w_every");
  if (rev_from) {
    SPARCH_REQUIRE(rev_from > 0 && Be == 2 * rev_from && rev_from % FT_ROWS == 0 && use_tma,
                   "reversed second half: Be = 2 * rev_from, rev_from a multiple of 128 rows, H a multiple of 4");
  }
  const int z_rows = rev_from ? rev_from : Be;
  FtMaps maps;
  memset(&maps, 0, sizeof maps);
  if (use_tma) {
    if (int e = make_map3d_f32(&maps.z, Z, z_rows, T, H, 32, FT_NEUR)) return e;
    if (S)
      if (int e = make_map3d_f32(&maps.s, S, Be, T, H, 32, FT_NEUR)) return e;
    if (int e = make_map3d_f32(&maps.u, U, Be, T, H, 32, FT_NEUR)) return e;
    if (int e = make_map3d_f32(&maps.w, adapt ? W : U, Be, (adapt && w_every) ? (T + w_every - 1) / w_every : T, H, 32, FT_NEUR)) return e;
  }
  RecFwdTcArgs p{Z, scale, shift, alpha, beta, a, b, rec0, u0, w0, s0, planes, colscale, theta, S, U, W, bits,
                 Be, T, H, Hp, NB, groups, reduced ? 1 : 0, use_tma, w_every, rev_from, recur_debug_buffer(), recur_debug_flags()};
  // a word of all ones means "not yet published" (published halves have zero odd bits): every step has its own words
  SPARCH_CUDA(cudaMemsetAsync(bits, 0xFF, sparch_recur_fwd_tc_bits_bytes(Be, T, H), st));
  const int gmax = max_ctas / ns;  // row groups per cooperative launch (all its CTAs wait on each other)
  const void* fn = adapt ? (const void*)rec_fwd_tc_kernel<true> : (const void*)rec_fwd_tc_kernel<false>;
  for (int g0 = 0; g0 < groups; g0 += gmax) {
    const int gn = groups - g0 < gmax ? groups - g0 : gmax;
    int group0 = g0;
    void* args[] = {(void*)&maps, (void*)&p, (void*)&group0};
    SPARCH_CUDA(cudaLaunchCooperativeKernel(fn, dim3(ns, gn), dim3(FT_THREADS), args, smem, st));
  }
  return SPARCH_OK;
}

}  // extern "C"

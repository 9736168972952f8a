// Recurrent kinds (RLIF, RadLIF): the per-timestep product s_{t-1} @ V0 on the tensor pipe, fused
// with the membrane update (snns.py:572, 718-724).
//
// Decomposition: a CTA owns 32 neurons (columns of V0) and runs TWO independent 4-warp teams, each
// owning 32 batch rows (its own named barrier, exchange buffers and arrival counter), so that one
// team's exchange latency overlaps the other team's tensor work; the teams share only the V0 image.  Its slice
// of V0 (Hp x 32) sits in shared memory as fp16 hi + fp16 lo (V0 * 2^k split in two, 22 mantissa
// bits: products with a spike are exact, accumulation is fp32), already arranged in
// mma.m16n8k16 B-fragment order so a warp loads its fragments with conflict-free 16-byte reads.
// The A operand is never materialised: spikes travel as 1-bit planes (32 neurons per word) and
// each thread synthesises its A fragments from the words with one rotate and one mask per
// register.  A spike is encoded as 2.0 (fp16 pattern 0x4000, a single set bit), and the K order
// inside a 32-neuron word is permuted so that the two halves of every fragment register are 16
// bits apart in the word (the same permutation is baked into the V0 image; the factor 2 into its
// scale).  The 4 warps of a team are 4 K-quarters (2 row tiles x 4 column tiles each); the K-quarters
// are reduced through shared memory, then the team's 128 threads apply the neuron update to its 32 x 32
// block (8 neurons per thread, state in registers) and emit the new spike words and the fp32 tapes.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "cell_math.cuh"
#include "common.cuh"

namespace sparch {

constexpr int RB = 32;       // batch rows per team (two teams per CTA)
constexpr int TEAMS = 2;
constexpr int TT = 128;      // threads per team
constexpr int RC = 32;       // neurons (V0 columns) per CTA
constexpr int RED_RS = 40;   // row stride (floats) of the K-quarter reduction buffer: conflict-free float2 stores
constexpr int VSCALE_EXP = 13;

// meta[0] = E with max|V0| < 2^E (int), set by vmax_kernel.
__global__ void vmax_kernel(const float* __restrict__ V, int H, int* __restrict__ meta) {
  // one warp per row (no per-element division); the diagonal does not count (snns.py:566/712 zeroes it)
  float m = 0.f;
  const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < H; row += gridDim.x * wpb) {
    const float* __restrict__ r = V + (int64_t)row * H;
#pragma unroll 4
    for (int c = lane; c < H; c += 32)
      if (c != row) m = fmaxf(m, fabsf(r[c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) atomicMax(&meta[1], __float_as_int(m));
}

__global__ void vmax_finish_kernel(int* meta) {
  float m = __int_as_float(meta[1]);
  int e = 0;
  if (m > 0.f && isfinite(m)) frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)
  meta[0] = e;
}

// Physical bit (neuron index inside a 32-neuron word) feeding logical K position `kpos` (0..15) of
// k-step `ks` (0/1) of that word.  kpos = 2q + 8hs + e.
__host__ __device__ __forceinline__ int spike_bit_of(int ks, int kpos) {
  int e = kpos & 1, hs = (kpos >> 3) & 1, q = (kpos >> 1) & 3;
  return (14 + 16 * e + 4 * q + 2 * ks + hs) & 31;
}

// Build the fragment-ordered fp16 hi/lo image of V0 (diagonal zeroed, snns.py:566/712).
//   transposed == 0 (forward):  B[k = presynaptic j][n = neuron c] = V[j][c], K bit-permuted
//   transposed == 1 (backward): B[k = neuron c][n = presynaptic j] = V[j][c], natural K order
// Image words per slice: (((kk*2 + nh)*2 + blk)*32 + lane)*4 + wd, kk < Hp/16.
__global__ void vprep_kernel(const float* __restrict__ V, int H, int Hp, int transposed,
                             const int* __restrict__ meta, uint32_t* __restrict__ img) {
  const int KK = Hp / 16;
  const int64_t words_per_slice = (int64_t)KK * 512;
  const int64_t total = words_per_slice * (Hp / RC);
  const int E = meta[0];
  const float sc = ldexpf(1.0f, (transposed ? VSCALE_EXP : VSCALE_EXP - 1) - E);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int slice = (int)(i / words_per_slice);
    int r = (int)(i % words_per_slice);
    int wd = r & 3, lane = (r >> 2) & 31, blk = (r >> 7) & 1, nh = (r >> 8) & 1, kk = r >> 9;
    int part = wd >> 1, reg = wd & 1, g = lane >> 2, q = lane & 3;
    int n = slice * RC + 8 * (2 * nh + blk) + g;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      int kpos = 2 * q + 8 * reg + e;
      int k = transposed ? 16 * kk + kpos : 32 * (kk >> 1) + spike_bit_of(kk & 1, kpos);
      int row = transposed ? n : k, col = transposed ? k : n;  // V[row][col], row = presynaptic j
      float x = 0.f;
      if (row < H && col < H && row != col) x = V[(int64_t)row * H + col] * sc;
      __half hi = __float2half_rn(x);
      v[e] = part ? (x - __half2float(hi)) : __half2float(hi);
    }
    __half2 h2 = __floats2half2_rn(v[0], v[1]);
    img[i] = *reinterpret_cast<uint32_t*>(&h2);
  }
}

static long long* g_dbg = nullptr;  // see sparch_recur_debug_clocks
static int g_dbg_flags = 0;
long long* recur_debug_buffer() { return g_dbg; }
int recur_debug_flags() { return g_dbg_flags; }

struct RecFwdArgs {
  const float *Z, *scale, *shift, *alpha, *beta, *a, *b, *rec0, *u0, *w0, *s0;
  const uint32_t* img;
  const int* meta;
  float theta;
  float *S, *U, *W;
  uint2* bits;     // [T][Be][Hp/32] {32 spikes, step tag t+1}: tagged 64-bit words (LL exchange)
  int Be, T, H, Hp;
  long long* dbg;  // optional [T][8] phase clocks of CTA (0,0) (profiling aid), normally NULL
  int dbg_flags;   // profiling experiments (results invalid): 1 no tag wait, 2 no tape stores, 4 no Z loads
  int reduced;     // 1: reduced-precision mode, hi term of V0 only (11 mantissa bits)
};

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// 8 consecutive floats: two 16-byte accesses when aligned, scalar with a bound otherwise.
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&v)[8], bool vec, int nv) {
  if (vec) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = i < nv ? p[i] : 0.f;
  }
}
__device__ __forceinline__ void store8(float* __restrict__ p, const float (&v)[8], bool vec, int nv) {
  if (vec) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) p[i] = v[i];
  }
}

// ------------------------------------------------------------------ reverse step
// recb[b, j] = sum_c dI_{t+1}[b, c] * V0[j, c] for the CTA's 32 presynaptic neurons j, then the
// BPTT update of SURVEY.md 8a for its 64 x 32 block.  The A operand is real-valued here, so the
// producer (the same block of step t+1) hands it over already split for the tensor pipe as
// block floating point: per (row, 32-column chunk) a power-of-two scale brings the chunk's
// largest |dI| into [8, 16), then hi = fp16(x), lo = fp16(x - hi) (22 mantissa bits inside the
// chunk, full fp32 range across chunks), stored in mma A-fragment order so the consumer's loads
// are linear 16-byte copies.  hi*Vhi + hi*Vlo + lo*Vhi are accumulated per chunk and folded into
// the fp32 accumulators with the chunk's inverse scale.  The V0^T image is resident in shared
// memory; the A panels stream through a cp.async ring (32 KB per stage: 4 chunks).
struct RecBwdArgs {
  const float *G, *U, *W, *alpha, *beta, *a, *b, *u0, *w0, *s0;
  const uint32_t* img;  // V0^T image (vprep transposed = 1)
  const int* meta;
  float theta;
  float *dI, *p_alpha, *p_beta, *p_a, *p_b;
  uint32_t* panel;  // 2 x [groups][Hp/32 chunks][1024 words]   (groups of 32 rows)
  float* pscale;    // 2 x [groups][Hp/32][32]
  int Be, T, H, Hp;
  long long* dbg;   // optional [T][8] phase clocks of CTA (0,0) (profiling aid), normally NULL
  int reduced;      // 1: reduced-precision mode, hi x hi product only, lo halves of the panel not moved
  int dbg_flags;    // profiling experiments (results invalid): 8 team 0 alone, 32 no ping-pong
};

// ------------------------------------------------------------------ persistent kernels
// One cooperative launch runs all T steps.  The V0 image is loaded into shared memory once and the
// per-neuron state (u, w, previous spike; adjoints and parameter-gradient sums in the reverse
// kernel) stays in registers across timesteps.  The only inter-CTA traffic per step is the
// exchange of the step's spike words (forward: tagged 8-byte words polled directly) or dI panels
// (reverse: ordered by one monotonically increasing counter per 32-row group).
__device__ __forceinline__ void team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TT) : "memory");
}
// Ping-pong of the two teams' tensor phases (named barriers 3 and 4, all 256 threads counted): a team
// enters its MMA loop only after the other team has left its own, so each finds the tensor pipe free.
__device__ __forceinline__ void pingpong_wait(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(3 + team), "n"(TEAMS * TT) : "memory");
}
__device__ __forceinline__ void pingpong_pass(int team) {  // lets the OTHER team go
  asm volatile("bar.arrive %0, %1;" ::"r"(3 + (team ^ 1)), "n"(TEAMS * TT) : "memory");
}

__device__ __forceinline__ void group_wait(const int* ctr, int target, int lt, int team) {
  if (lt == 0) {
    const long long t0 = clock64();
    while (true) {
      int v;
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 4000000000LL) __trap();  // a lost arrival must not hang the GPU
    }
  }
  team_sync(team);
}
__device__ __forceinline__ void group_arrive(int* ctr, int lt, int team) {
  // The team barrier orders every thread's panel stores before thread 0's release at gpu scope
  // (release is cumulative over what the barrier made visible to thread 0): one fence, not 128.
  team_sync(team);
  if (lt == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
}

// Forward teams are 8 warps (4 K-quarters x 2 column halves, 512 threads per CTA): a warp's HMMA issue
// rate bounds a team's MMA phase, so the forward kernel -- whose register budget allows it -- halves
// each warp's share; the neuron update then has 4 neurons per thread.
constexpr int FW = 8;                 // warps per forward team
constexpr int FT = FW * 32;           // threads per forward team
constexpr int FKQ = 4;                // K-quarters; the 8 warps are 4 K-quarters x 2 column halves
constexpr int FWD_TEAM_WORDS_FIXED = FKQ * RB * RED_RS;  // reduction buffer (floats); spike tile follows

__device__ __forceinline__ void fteam_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(FT) : "memory");
}
__device__ __forceinline__ void fpingpong_wait(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(3 + team), "n"(TEAMS * FT) : "memory");
}
__device__ __forceinline__ void fpingpong_pass(int team) {
  asm volatile("bar.arrive %0, %1;" ::"r"(3 + (team ^ 1)), "n"(TEAMS * FT) : "memory");
}
__device__ __forceinline__ void load4(const float* __restrict__ p, float (&v)[4], bool vec, int nv) {
  if (vec) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = i < nv ? p[i] : 0.f;
  }
}
__device__ __forceinline__ void store4(float* __restrict__ p, const float (&v)[4], bool vec, int nv) {
  if (vec) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < nv) p[i] = v[i];
  }
}

template <bool ADAPT>
__global__ void __launch_bounds__(TEAMS* FT, 1) rec_fwd_persist_kernel(const RecFwdArgs p, const int group0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NW = p.Hp / 32;   // spike words per row
  const int RSB = NW + 1;     // padded row stride of the spike-word tile
  const int tid = threadIdx.x, lane = tid & 31, team = tid / FT, lt = tid % FT;
  const int kq = ((tid >> 5) % FW) >> 1, nh = (tid >> 5) & 1, g = lane >> 2, q = lane & 3;
  uint32_t* simg = reinterpret_cast<uint32_t*>(smem_raw);
  const size_t team_words = (size_t)FWD_TEAM_WORDS_FIXED + (size_t)RB * RSB;
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)p.Hp * 128) + team * team_words;
  uint32_t* sbits = reinterpret_cast<uint32_t*>(red + FWD_TEAM_WORDS_FIXED);

  const int slice = blockIdx.x, group = group0 + TEAMS * blockIdx.y + team, row0 = group * RB;

  {
    const uint4* src = reinterpret_cast<const uint4*>(p.img + (size_t)slice * p.Hp * 32);
    uint4* dst = reinterpret_cast<uint4*>(simg);
    for (int i = tid; i < p.Hp * 8; i += TEAMS * FT) cp_async16(dst + i, src + i);
  }
  __shared__ float sprm[8][RC];  // alpha, 1-alpha, beta, a, b, 1/(1-alpha), scale, shift of the slice
  if (tid < RC) {
    const int col = min(slice * RC + tid, p.H - 1);
    const NeuronParams q0 = load_params<ADAPT>(p.alpha, p.beta, p.a, p.b, col);
    sprm[0][tid] = q0.alpha; sprm[1][tid] = q0.oma; sprm[2][tid] = q0.beta; sprm[3][tid] = q0.a;
    sprm[4][tid] = q0.b; sprm[5][tid] = 1.0f / q0.oma;
    sprm[6][tid] = p.scale ? p.scale[col] : 1.0f;
    sprm[7][tid] = p.scale ? p.shift[col] : 0.0f;
  }
  const int r = lt >> 3, cg = lt & 7;   // 256 threads on the team's 32 x 32 block: 4 neurons of one row each
  const int row = row0 + r;
  const int col0 = slice * RC + cg * 4;
  const bool live = row < p.Be && col0 < p.H;
  const bool vec = ((p.H & 3) == 0) && (col0 + 4 <= p.H);
  const int nv = live ? min(4, p.H - col0) : 0;
  const int64_t idx0 = (int64_t)row * p.H + col0;
  const float rs = ldexpf(1.0f, p.meta[0] - VSCALE_EXP);
  float u[4], w[4], s[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = w[i] = s[i] = 0.f;
  if (live) {
    load4(p.u0 + idx0, u, vec, nv);
    load4(p.s0 + idx0, s, vec, nv);
    if (ADAPT) load4(p.w0 + idx0, w, vec, nv);
  }
  cp_async_wait_all();
  __syncthreads();            // V0 image and parameter table visible to both teams
  if (row0 >= p.Be) return;   // odd number of row groups: the last CTA's second team has no rows
  if (team == 1 && (p.dbg_flags & 8)) return;  // experiment: team 0 alone on the SM (team 1's rows are not computed)
  if (team == 1 && p.dbg_flags >= 16) {  // experiment: start team 1 (dbg_flags / 16) * 256 cycles late
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)(p.dbg_flags / 16) * 256) {}
  }
  // (Starting team 1 half a step late, to run the two pipelines in anti-phase, was measured: no gain --
  // a single warp issues one HMMA.16816 per ~18 cycles, so a team's MMA phase lasts ~4.7 k cycles with
  // or without the other team competing for the tensor pipe.)

  const uint4* bimg = reinterpret_cast<const uint4*>(simg);
  bool tapes_pending = false;
  // both teams of this CTA have rows (and the experiment flags leave both running): alternate MMA phases
  const bool pingpong = (group0 + TEAMS * blockIdx.y + 1) * RB < p.Be && !(p.dbg_flags & (8 | 32));
  if (pingpong && team == 1) fpingpong_pass(1);  // team 0 goes first
  for (int t = 0; t < p.T; ++t) {
    const int64_t o0 = ((int64_t)row * p.T + t) * p.H + col0;
    float z[4], rec[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = rec[i] = 0.f;
    const bool dbg_on = p.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    if (live && !(p.dbg_flags & 4)) load4(p.Z + o0, z, vec, nv);  // independent of the exchange: issued before the wait
    if (t > 0) {
      // Wait for and fetch the spike words of step t-1 in one go: each word travels with its step
      // tag in a single 8-byte store, so a matching tag means the data is there (no fence, no
      // separate flag).  All of a thread's loads are issued before any result is looked at.
      const uint2* bsrc = p.bits + (size_t)(t - 1) * p.Be * NW;
      const long long t0 = clock64();
      if (dbg_on) p.dbg[t * 8 + 4] = t0;
      for (int base = 0; base < RB * NW; base += FT * 4) {
        const uint2* src[4];
        int dst[4];
        bool need[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = base + k * FT + lt;
          const int rr = i / NW, wi = i - rr * NW;
          need[k] = i < RB * NW && row0 + rr < p.Be;
          dst[k] = i < RB * NW ? rr * RSB + wi : -1;
          src[k] = bsrc + (size_t)(row0 + rr) * NW + wi;
        }
        uint32_t bv[4], tg[4];
        bool ok;
        do {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            bv[k] = 0;
            tg[k] = (uint32_t)t;
            if (need[k])
              asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];"
                           : "=r"(bv[k]), "=r"(tg[k])
                           : "l"(src[k])
                           : "memory");
          }
          if (tapes_pending) {  // step t-1's tape stores ride in the shadow of the L2 round trip
            tapes_pending = false;
            const int64_t op = o0 - p.H;
            store4(p.S + op, s, vec, nv);
            store4(p.U + op, u, vec, nv);
            if (ADAPT) store4(p.W + op, w, vec, nv);
          }
          ok = true;
#pragma unroll
          for (int k = 0; k < 4; ++k) ok = ok && (tg[k] == (uint32_t)t || (p.dbg_flags & 1));
          if (!ok && clock64() - t0 > 4000000000LL) __trap();  // a lost store must not hang the GPU
        } while (!ok);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (dst[k] >= 0) sbits[dst[k]] = bv[k];
      }
      fteam_sync(team);
      if (pingpong) fpingpong_wait(team);
      if (dbg_on) p.dbg[t * 8 + 0] = clock64();
      const bool dbg_on1 = p.dbg && tid == FT && blockIdx.x == 0 && blockIdx.y == 0;  // team 1's view
      if (dbg_on1) p.dbg[t * 8 + 5] = clock64();

      float acc[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
      for (int wi = kq; wi < NW; wi += FKQ) {
        uint32_t wa[2], wb[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t x0 = sbits[(16 * mt + g) * RSB + wi], x1 = sbits[(16 * mt + g + 8) * RSB + wi];
          wa[mt] = __funnelshift_r(x0, x0, 4 * q);
          wb[mt] = __funnelshift_r(x1, x1, 4 * q);
        }
        // both k-steps of the word: 4 passes over the 8 accumulators (hi ks0, hi ks1, lo ks0, lo ks1), so
        // two HMMAs on the same accumulator are always 8 instructions apart
        uint4 f[2][2];
        uint32_t af[2][2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const int kk = 2 * wi + ks;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) f[ks][nt] = bimg[((kk * 2 + nh) * 2 + nt) * 32 + lane];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t M = 0x40004000u;
            af[ks][mt][0] = __funnelshift_r(wa[mt], wa[mt], 2 * ks) & M;
            af[ks][mt][1] = __funnelshift_r(wb[mt], wb[mt], 2 * ks) & M;
            af[ks][mt][2] = __funnelshift_r(wa[mt], wa[mt], 2 * ks + 1) & M;
            af[ks][mt][3] = __funnelshift_r(wb[mt], wb[mt], 2 * ks + 1) & M;
          }
        }
        const int nparts = p.reduced ? 1 : 2;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          if (part >= nparts) break;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt)
                mma16816(acc[mt][nt], af[ks][mt][0], af[ks][mt][1], af[ks][mt][2], af[ks][mt][3],
                         part ? f[ks][nt].z : f[ks][nt].x, part ? f[ks][nt].w : f[ks][nt].y);
        }
      }
      if (pingpong) fpingpong_pass(team);
      if (dbg_on) p.dbg[t * 8 + 1] = clock64();
      if (dbg_on1) p.dbg[t * 8 + 6] = clock64();
      float* myred = red + kq * RB * RED_RS;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int col = 16 * nh + 8 * nt + 2 * q;
          *reinterpret_cast<float2*>(&myred[(16 * mt + g) * RED_RS + col]) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
          *reinterpret_cast<float2*>(&myred[(16 * mt + g + 8) * RED_RS + col]) =
              make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
      fteam_sync(team);
      {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < FKQ; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(&red[(k * RB + r) * RED_RS + cg * 4]);
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        rec[0] = sum.x * rs; rec[1] = sum.y * rs; rec[2] = sum.z * rs; rec[3] = sum.w * rs;
      }
      if (dbg_on) p.dbg[t * 8 + 2] = clock64();
    } else if (live) {
      load4(p.rec0 + idx0, rec, vec, nv);
    }
    uint32_t my = 0;
    if (live) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < nv) {
          const int lc = cg * 4 + i;
          NeuronParams npi;
          npi.alpha = sprm[0][lc]; npi.oma = sprm[1][lc]; npi.beta = sprm[2][lc]; npi.a = sprm[3][lc];
          npi.b = sprm[4][lc];
          float cur = p.scale ? __fmaf_rn(z[i], sprm[6][lc], sprm[7][lc]) : z[i];
          cur = __fadd_rn(cur, rec[i]);
          float wi_ = ADAPT ? w[i] : 0.f;
          step_fwd<ADAPT>(npi, cur, p.theta, u[i], wi_, s[i]);
          w[i] = wi_;
          my |= (s[i] > 0.f ? 1u : 0u) << (cg * 4 + i);
        }
      }
    }
    my |= __shfl_xor_sync(0xffffffffu, my, 1);
    my |= __shfl_xor_sync(0xffffffffu, my, 2);
    my |= __shfl_xor_sync(0xffffffffu, my, 4);
    if (cg == 0 && row < p.Be)   // publish first: this store is what the other slices wait for
      asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p.bits + ((size_t)t * p.Be + row) * NW + slice),
                   "r"(my), "r"((uint32_t)(t + 1))
                   : "memory");
    // The tapes are only read after the kernel: their stores are deferred until the next step's
    // spike-word loads are in flight (u, w, s stay unchanged in registers until then).
    tapes_pending = live && !(p.dbg_flags & 2);
    if (tapes_pending && t == p.T - 1) {
      store4(p.S + o0, s, vec, nv);
      store4(p.U + o0, u, vec, nv);
      if (ADAPT) store4(p.W + o0, w, vec, nv);
    }
    fteam_sync(team);            // sbits / red are rewritten by the next step
    if (dbg_on) p.dbg[t * 8 + 3] = clock64();
  }
}

constexpr int PB_CHUNK_WORDS = 1024;                   // one 32-column chunk of a 32-row panel: 4 KB
constexpr int PB_STAGE_WORDS = 4 * PB_CHUNK_WORDS;     // 4 chunks (one per K-quarter) = 16 KB
constexpr int PB_STAGES = 2;

template <bool ADAPT>
__global__ void __launch_bounds__(TEAMS* TT, 1)
rec_bwd_persist_kernel(const RecBwdArgs p, const int group0, const int ngroups_total, int* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NCH = p.Hp / 32;
  const int tid = threadIdx.x, lane = tid & 31, team = tid >> 7, lt = tid & (TT - 1);
  const int kq = (tid >> 5) & 3, g = lane >> 2, q = lane & 3;
  uint32_t* simg = reinterpret_cast<uint32_t*>(smem_raw);                          // V0^T image (shared by both teams)
  const size_t team_words = (size_t)PB_STAGES * PB_STAGE_WORDS + (size_t)NCH * RB;
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + (size_t)p.Hp * 128) + team * team_words;  // A stages
  float* sscale = reinterpret_cast<float*>(ring + PB_STAGES * PB_STAGE_WORDS);
  float* red = reinterpret_cast<float*>(ring);                                     // aliases the ring after the K loop

  const int slice = blockIdx.x, group = group0 + TEAMS * blockIdx.y + team, row0 = group * RB;
  const int nslices = gridDim.x;
  int* ctr = counters + group;

  {
    const uint4* src = reinterpret_cast<const uint4*>(p.img + (size_t)slice * p.Hp * 32);
    uint4* dst = reinterpret_cast<uint4*>(simg);
    for (int i = tid; i < p.Hp * 8; i += TEAMS * TT) cp_async16(dst + i, src + i);
  }
  __shared__ float sprm[6][RC];  // alpha, 1-alpha, beta, a, b, 1/(1-alpha) of the slice
  if (tid < RC) {
    const int col = min(slice * RC + tid, p.H - 1);
    const NeuronParams q0 = load_params<ADAPT>(p.alpha, p.beta, p.a, p.b, col);
    sprm[0][tid] = q0.alpha; sprm[1][tid] = q0.oma; sprm[2][tid] = q0.beta; sprm[3][tid] = q0.a;
    sprm[4][tid] = q0.b; sprm[5][tid] = 1.0f / q0.oma;
  }
  const int r = lt >> 2, cg = lt & 3;
  const int row = row0 + r;
  const int col0 = slice * RC + cg * 8;
  const bool live = row < p.Be && col0 < p.H;
  const bool vec = ((p.H & 3) == 0) && (col0 + 8 <= p.H);
  const int nv = live ? min(8, p.H - col0) : 0;
  const int64_t idx0 = (int64_t)row * p.H + col0;
  const float rs = ldexpf(1.0f, p.meta[0] - VSCALE_EXP);
  float du[8], dw[8], pa[8], pb[8], pc[8], pd[8], ut[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) du[i] = dw[i] = pa[i] = pb[i] = pc[i] = pd[i] = ut[i] = 0.f;
  if (live && p.T > 0) load8(p.U + ((int64_t)row * p.T + (p.T - 1)) * p.H + col0, ut, vec, nv);
  cp_async_wait_all();
  __syncthreads();            // V0^T image and parameter table visible to both teams
  if (row0 >= p.Be) return;   // odd number of row groups: the last CTA's second team has no rows
  if (team == 1 && (p.dbg_flags & 8)) return;  // experiment: team 0 alone on the SM
  // (Alternating the two teams' panel-stream + MMA phases like the forward kernel does was measured:
  // a team alone needs 10.5 k cycles for that phase, 13.7 k when both overlap -- serialising them
  // (2 x 10.5 k) is slower than overlapping, 22.5 k vs 21.1 k per step.  The switch is kept for experiments.)
  const bool pingpong = (p.dbg_flags & 64) && (group0 + TEAMS * blockIdx.y + 1) * RB < p.Be;
  if (pingpong && team == 1) pingpong_pass(1);

  const uint4* bimg = reinterpret_cast<const uint4*>(simg);
  const int NSC = (NCH + 3) / 4;

  for (int t = p.T - 1; t >= 0; --t) {
    const int64_t o0 = ((int64_t)row * p.T + t) * p.H + col0;
    const int rbuf = (t + 1) & 1, wbuf = t & 1;
    float gq[8], up[8], wp[8], sp[8], recb[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gq[i] = up[i] = wp[i] = sp[i] = recb[i] = d[i] = 0.f;
    if (live) {  // tape reads are independent of the exchange: issued before the wait
      load8(p.G + o0, gq, vec, nv);
      if (t > 0) {
        load8(p.U + o0 - p.H, up, vec, nv);
        if (ADAPT) load8(p.W + o0 - p.H, wp, vec, nv);
      } else {
        load8(p.u0 + idx0, up, vec, nv);
        load8(p.s0 + idx0, sp, vec, nv);
        if (ADAPT) load8(p.w0 + idx0, wp, vec, nv);
      }
    }
    const bool dbg_on = p.dbg && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
    if (dbg_on) p.dbg[t * 8 + 0] = clock64();
    if (t < p.T - 1) {
      group_wait(ctr, nslices * (p.T - 1 - t), lt, team);
      if (pingpong) pingpong_wait(team);
      if (dbg_on) p.dbg[t * 8 + 1] = clock64();
      const uint32_t* apanel = p.panel + ((size_t)rbuf * ngroups_total + group) * NCH * PB_CHUNK_WORDS;
      const float* gsc = p.pscale + ((size_t)rbuf * ngroups_total + group) * NCH * RB;
      // Each warp streams ITS chunks (c = 4j + kq) through a private 2-slot ring: no team barrier inside
      // the K loop, so the four warps drift apart and one warp's L2 wait is covered by the others' MMAs.
      uint32_t* wring = ring + (size_t)kq * (2 * PB_CHUNK_WORDS);
      auto issue = [&](int j) {
        const int c = 4 * j + kq;
        if (c < NCH) {
          const uint4* sa = reinterpret_cast<const uint4*>(apanel + (size_t)c * PB_CHUNK_WORDS);
          uint4* da = reinterpret_cast<uint4*>(wring + (size_t)(j & 1) * PB_CHUNK_WORDS);
          // 16-byte unit i: bit 5 selects the hi (0) / lo (1) block of a (row tile, k-step); the reduced
          // mode never touches the lo blocks
          for (int i = lane; i < PB_CHUNK_WORDS / 4; i += 32)
            if (!(p.reduced && (i & 32))) cp_async16(da + i, sa + i);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
      };
      issue(0);
      for (int i = lt; i < NCH * RB; i += TT) sscale[i] = __ldcg(&gsc[i]);
      team_sync(team);   // scales visible (the ring itself is warp-private)
      float acc[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
      for (int j = 0; j < NSC; ++j) {
        issue(j + 1);
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        const int c = 4 * j + kq;
        if (c < NCH) {
          const uint4* a4 = reinterpret_cast<const uint4*>(wring + (size_t)(j & 1) * PB_CHUNK_WORDS);
          float tacc[2][4][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int i = 0; i < 4; ++i) tacc[mt][nt][i] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int kk = 2 * c + ks;
            uint4 f[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) f[nt] = bimg[(kk * 4 + nt) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint4 ah = a4[((mt * 2 + ks) * 2 + 0) * 32 + lane];
              const uint4 al = a4[((mt * 2 + ks) * 2 + 1) * 32 + lane];
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) mma16816(tacc[mt][nt], ah.x, ah.y, ah.z, ah.w, f[nt].x, f[nt].y);
              if (!p.reduced) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma16816(tacc[mt][nt], ah.x, ah.y, ah.z, ah.w, f[nt].z, f[nt].w);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma16816(tacc[mt][nt], al.x, al.y, al.z, al.w, f[nt].x, f[nt].y);
              }
            }
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const float s_lo = sscale[c * RB + 16 * mt + g], s_hi = sscale[c * RB + 16 * mt + g + 8];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              acc[mt][nt][0] = fmaf(tacc[mt][nt][0], s_lo, acc[mt][nt][0]);
              acc[mt][nt][1] = fmaf(tacc[mt][nt][1], s_lo, acc[mt][nt][1]);
              acc[mt][nt][2] = fmaf(tacc[mt][nt][2], s_hi, acc[mt][nt][2]);
              acc[mt][nt][3] = fmaf(tacc[mt][nt][3], s_hi, acc[mt][nt][3]);
            }
          }
        }
        __syncwarp();    // every lane is done with the slot before the next iteration refills it
      }
      team_sync(team);   // all warps are done with the ring: the reduction buffer aliases it
      if (pingpong) pingpong_pass(team);
      if (dbg_on) p.dbg[t * 8 + 2] = clock64();
      float* myred = red + kq * RB * RED_RS;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = 8 * nt + 2 * q;
          *reinterpret_cast<float2*>(&myred[(16 * mt + g) * RED_RS + col]) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
          *reinterpret_cast<float2*>(&myred[(16 * mt + g + 8) * RED_RS + col]) =
              make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
      team_sync(team);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cg * 8 + i;
        recb[i] = ((red[(0 * RB + r) * RED_RS + c] + red[(1 * RB + r) * RED_RS + c]) +
                   (red[(2 * RB + r) * RED_RS + c] + red[(3 * RB + r) * RED_RS + c])) * rs;
      }
    }
    if (live) {
      if (t > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sp[i] = spike_of(__fsub_rn(up[i], p.theta));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < nv) {
          const int lc = cg * 8 + i;
          NeuronParams npi;
          npi.alpha = sprm[0][lc]; npi.oma = sprm[1][lc]; npi.beta = sprm[2][lc]; npi.a = sprm[3][lc];
          npi.b = sprm[4][lc];
          float dwi = dw[i], pbi = pb[i], pci = pc[i], pdi = pd[i];
          d[i] = step_bwd<ADAPT>(npi, sprm[5][lc], p.theta, gq[i], recb[i], ut[i], up[i], sp[i], wp[i],
                                 du[i], dwi, pa[i], pbi, pci, pdi);
          dw[i] = dwi; pb[i] = pbi; pc[i] = pci; pd[i] = pdi;
          ut[i] = up[i];
        }
      }
    }
    if (t > 0) {
      // hand dI_t to step t-1: block floating point fp16 hi/lo in mma A-fragment order
      float m = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(d[i]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      int e = 0;
      if (m > 0.f && m <= 3.0e38f) frexpf(m, &e);
      e = max(e, -100);
      const float up_scale = ldexpf(1.0f, 4 - e), inv_scale = ldexpf(1.0f, e - 4);
      uint32_t* wpanel = p.panel + (((size_t)wbuf * ngroups_total + group) * NCH + slice) * PB_CHUNK_WORDS;
      const int mt = r >> 4, rr = r & 15, gg = rr & 7, upper = rr >> 3, ks = cg >> 1, hs = cg & 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x0 = d[2 * j] * up_scale, x1 = d[2 * j + 1] * up_scale;
        __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        __half2 hi = __halves2half2(h0, h1);
        __half2 lo = __floats2half2_rn(x0 - __half2float(h0), x1 - __half2float(h1));
        const int wd = ((mt * 2 + ks) * 2 * 32 + (4 * gg + j)) * 4 + upper + 2 * hs;
        wpanel[wd] = *reinterpret_cast<uint32_t*>(&hi);
        wpanel[wd + 128] = *reinterpret_cast<uint32_t*>(&lo);
      }
      if (cg == 0) p.pscale[(((size_t)wbuf * ngroups_total + group) * NCH + slice) * RB + r] = inv_scale;
      if (dbg_on) p.dbg[t * 8 + 3] = clock64();
      group_arrive(ctr, lt, team);
      if (dbg_on) p.dbg[t * 8 + 4] = clock64();
    }
    if (live) store8(p.dI + o0, d, vec, nv);  // tape store after the hand-over: off the critical path
  }
  if (live) {
    store8(p.p_alpha + idx0, pa, vec, nv);
    if (ADAPT) {
      store8(p.p_beta + idx0, pb, vec, nv);
      store8(p.p_a + idx0, pc, vec, nv);
      store8(p.p_b + idx0, pd, vec, nv);
    }
  }
}

static size_t rec_bwd_persist_smem(int Hp) {
  return (size_t)Hp * 128 + TEAMS * ((size_t)PB_STAGES * PB_STAGE_WORDS * 4 + (size_t)(Hp / 32) * RB * 4);
}

static size_t rec_fwd_smem(int Hp) {
  return (size_t)Hp * 128 + TEAMS * ((size_t)FWD_TEAM_WORDS_FIXED * 4 + (size_t)RB * (Hp / 32 + 1) * 4);
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_recur_padded(int H) { return ((H + 31) / 32) * 32; }

int sparch_recur_debug_clocks(long long* buf) {
  g_dbg = buf;
  const char* f = getenv("SPARCH_B200_DEBUG_FLAGS");
  g_dbg_flags = (buf && f) ? atoi(f) : 0;
  return SPARCH_OK;
}

int sparch_recur_prepare(const float* V, int H, uint32_t* img_fwd, uint32_t* img_bwd, int* meta,
                         sparch_stream_t st_) {
  SPARCH_REQUIRE(V && H > 0 && meta, "null pointer");  // both images NULL: meta (max|V0|) only
  cudaStream_t st = as_stream(st_);
  const int Hp = sparch_recur_padded(H);
  SPARCH_CUDA(cudaMemsetAsync(meta, 0, 2 * sizeof(int), st));
  int nb = (H + 7) / 8;                       // 8 warps per block, one row per warp
  if (nb > sm_count() * 8) nb = sm_count() * 8;
  vmax_kernel<<<nb, 256, 0, st>>>(V, H, meta);
  SPARCH_LAUNCH_OK();
  vmax_finish_kernel<<<1, 1, 0, st>>>(meta);
  SPARCH_LAUNCH_OK();
  int64_t total = (int64_t)Hp * Hp;  // words
  int pb = (int)((total + 255) / 256);
  if (pb > sm_count() * 16) pb = sm_count() * 16;
  if (img_fwd) {
    vprep_kernel<<<pb, 256, 0, st>>>(V, H, Hp, 0, meta, img_fwd);
    SPARCH_LAUNCH_OK();
  }
  if (img_bwd) {
    vprep_kernel<<<pb, 256, 0, st>>>(V, H, Hp, 1, meta, img_bwd);
    SPARCH_LAUNCH_OK();
  }
  return SPARCH_OK;
}

int sparch_recur_fwd(int kind, const float* Z, const float* scale, const float* shift,
                     const float* alpha, const float* beta, const float* a, const float* b,
                     const float* rec0, const uint32_t* img_fwd, const int* meta, const float* u0,
                     const float* w0, const float* s0, float theta, float* S, float* U, float* W,
                     uint32_t* bits, int reduced, int Be, int T, int H, sparch_stream_t st_) {
  SPARCH_REQUIRE(kind == SPARCH_RLIF || kind == SPARCH_RADLIF, "recurrent kinds only");
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0, "bad shape");
  SPARCH_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
  if (Be == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(Z && alpha && rec0 && img_fwd && meta && u0 && s0 && S && U && bits, "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (beta && a && b && w0 && W), "adaptive kind needs beta, a, b, w0, W");
  const int Hp = sparch_recur_padded(H);
  const size_t smem = rec_fwd_smem(Hp);
  SPARCH_REQUIRE(smem <= 225 * 1024, "hidden size too large for the resident V0 slice");
  RecFwdArgs p{Z, scale, shift, alpha, beta, a, b, rec0, u0, w0, s0, img_fwd, meta, theta, S, U, W,
               reinterpret_cast<uint2*>(bits), Be, T, H, Hp, g_dbg, g_dbg_flags, reduced ? 1 : 0};
  cudaStream_t st = as_stream(st_);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    SPARCH_CUDA(cudaFuncSetAttribute(rec_fwd_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    SPARCH_CUDA(cudaFuncSetAttribute(rec_fwd_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  }
  const int max_ctas = sm_count();  // one CTA per SM (shared memory bound)
  const int slices = Hp / RC, groups = (Be + RB - 1) / RB;
  SPARCH_REQUIRE(slices <= max_ctas, "hidden size needs more co-resident CTAs than the GPU has SMs");
  const int gmax = TEAMS * (max_ctas / slices);   // row groups (of 32) per cooperative launch
  // tags are t+1 >= 1: a zeroed buffer means "nothing published yet"
  SPARCH_CUDA(cudaMemsetAsync(bits, 0, sizeof(uint2) * (size_t)T * Be * (Hp / 32), st));
  for (int g0 = 0; g0 < groups; g0 += gmax) {
    int gn = groups - g0 < gmax ? groups - g0 : gmax;
    dim3 cgrid(slices, (gn + TEAMS - 1) / TEAMS);
    int group0 = g0;
    void* args[] = {(void*)&p, (void*)&group0};
    const void* fn = adapt ? (const void*)rec_fwd_persist_kernel<true> : (const void*)rec_fwd_persist_kernel<false>;
    SPARCH_CUDA(cudaLaunchCooperativeKernel(fn, cgrid, dim3(TEAMS * FT), args, smem, st));
  }
  return SPARCH_OK;
}

int sparch_recur_sync_words(int Be) { return (Be + RB - 1) / RB + 1; }

size_t sparch_recur_bwd_workspace(int Be, int H) {
  const int Hp = sparch_recur_padded(H);
  const size_t groups = (size_t)(Be + RB - 1) / RB;
  return 2 * groups * (Hp / 32) * PB_CHUNK_WORDS * 4 + 2 * groups * (Hp / 32) * RB * 4;
}

int sparch_recur_bwd(int kind, const float* G, const float* U, const float* W, const float* alpha,
                     const float* beta, const float* a, const float* b, const uint32_t* img_bwd,
                     const int* meta, const float* u0, const float* w0, const float* s0, float theta,
                     float* dI, float* p_alpha, float* p_beta, float* p_a, float* p_b, void* workspace,
                     int* sync_ws, int reduced, int Be, int T, int H, sparch_stream_t st_) {
  SPARCH_REQUIRE(kind == SPARCH_RLIF || kind == SPARCH_RADLIF, "recurrent kinds only");
  SPARCH_REQUIRE(Be >= 0 && T >= 0 && H > 0, "bad shape");
  if (Be == 0 || T == 0) return SPARCH_OK;
  SPARCH_REQUIRE(G && U && alpha && img_bwd && meta && u0 && s0 && dI && p_alpha && workspace && sync_ws,
                 "null pointer");
  const bool adapt = kind & 1;
  SPARCH_REQUIRE(!adapt || (W && beta && a && b && w0 && p_beta && p_a && p_b),
                 "adaptive kind needs W, beta, a, b, w0 and the partial buffers");
  const int Hp = sparch_recur_padded(H);
  const size_t groups = (size_t)(Be + RB - 1) / RB;
  uint32_t* panel = reinterpret_cast<uint32_t*>(workspace);
  float* pscale = reinterpret_cast<float*>(panel + 2 * groups * (Hp / 32) * PB_CHUNK_WORDS);
  RecBwdArgs p{G, U, W, alpha, beta, a, b, u0, w0, s0, img_bwd, meta, theta, dI,
               p_alpha, p_beta, p_a, p_b, panel, pscale, Be, T, H, Hp, g_dbg, reduced ? 1 : 0, g_dbg_flags};
  cudaStream_t st = as_stream(st_);
  const size_t psmem = rec_bwd_persist_smem(Hp);
  SPARCH_REQUIRE(psmem <= 225 * 1024, "hidden size too large for the resident V0^T slice");
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    SPARCH_CUDA(cudaFuncSetAttribute(rec_bwd_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    SPARCH_CUDA(cudaFuncSetAttribute(rec_bwd_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  }
  const int max_ctas = sm_count();
  const int slices = Hp / RC;
  SPARCH_REQUIRE(slices <= max_ctas, "hidden size needs more co-resident CTAs than the GPU has SMs");
  const int gmax = TEAMS * (max_ctas / slices);
  SPARCH_CUDA(cudaMemsetAsync(sync_ws, 0, sizeof(int) * groups, st));
  for (int g0 = 0; g0 < (int)groups; g0 += gmax) {
    int gn = (int)groups - g0 < gmax ? (int)groups - g0 : gmax;
    dim3 cgrid(slices, (gn + TEAMS - 1) / TEAMS);
    int group0 = g0, ngt = (int)groups;
    void* args[] = {(void*)&p, (void*)&group0, (void*)&ngt, (void*)&sync_ws};
    const void* fn = adapt ? (const void*)rec_bwd_persist_kernel<true> : (const void*)rec_bwd_persist_kernel<false>;
    SPARCH_CUDA(cudaLaunchCooperativeKernel(fn, cgrid, dim3(TEAMS * TT), args, psmem, st));
  }
  return SPARCH_OK;
}

}  // extern "C"

// tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY (written by tcgen05.st from registers), B = int8
// digit planes in shared memory (K-major SWIZZLE_128B), D = int32 in TMEM: the building block of the tcgen05
// forward recurrence (csrc/recur_fwd_tc.cu).  Checks the operand layouts against a CPU product and measures
// the issue rate of a K = 1024 step for several N.
//
//   A[128 rows][K] uint8 {0,1}: TMEM lane = row, 32-bit column c holds K positions 4c .. 4c+3 (byte e = K 4c+e)
//   B[N][K] int8: tile kb (128 K) at kb * N * 128 bytes, byte (n, k) at n*128 + (((k/16) ^ (n&7)) * 16) + k%16
//   D[128][N] int32: lane = row, column = n
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o umma_i8_ts.bin umma_i8_ts.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}

#define ST16(taddr, v, o)                                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),  \
               "r"(v[o + 6]), "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]),              \
               "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]) : "memory")

#define LD16(taddr, v, o)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),        \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),      \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                     \
               : "r"(taddr))

// spike bits -> bytes: 16 bits of word w -> 4 columns; column j holds bits j, j+4, j+8, j+12 (bytes 0..3)
__device__ __forceinline__ uint32_t expand4(uint32_t m16) {  // m16: nibbles in {0,1}
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0x00000100u), "r"(0u), "r"(m16));
  return r;
}

constexpr int K = 1024, M = 128, KCOLS = K / 4;

// bits: [128 rows][K/16] uint16 words (bit i of word p = spike of K-neuron 16p + i)
// bimg: B image as laid out above (N * K bytes), dout: [128][N] int32
template <int N>
__global__ void __launch_bounds__(160, 1) umma_kernel(const uint16_t* __restrict__ bits, const uint8_t* __restrict__ bimg,
                                                       int* __restrict__ dout, long long* __restrict__ cyc, int reps, uint32_t d_col) {
  extern __shared__ unsigned char sm_raw[];
  const uint32_t raw = smem_u32(sm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* sm = sm_raw + (base - raw);
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bars[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * K / 16; i += 160)
    reinterpret_cast<uint4*>(sm)[i] = reinterpret_cast<const uint4*>(bimg)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(128u));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[1])), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t a_col = 0;
  if (warp < 4) {
    // row = tid: expand this row's spike bits into TMEM columns
    const uint16_t* rb = bits + (size_t)tid * (K / 16);
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < KCOLS; c0 += 16) {  // 16 columns = 4 words of 16 spikes
      uint32_t v[16];
#pragma unroll
      for (int wq = 0; wq < 4; ++wq) {
        const uint32_t w = rb[c0 / 4 + wq];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[4 * wq + j] = expand4((w >> j) & 0x1111u);
      }
      ST16(lane_addr + a_col + c0, v, 0);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
    mbar_wait(smem_u32(&bars[1]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      LD16(lane_addr + d_col + n0, v, 0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) dout[(size_t)tid * N + n0 + i] = (int)v[i];
    }
  } else {
    // the WHOLE warp runs the issue loop (converged); one elected lane issues each UMMA: no per-instruction
    // election loop in SASS (a divergent `if (lane == 0)` costs ~45 cycles per UMMA in ELECT/PLOP3/BRA.U.ANY)
    mbar_wait(smem_u32(&bars[0]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // kind::i8: D = S32 (2 << 4), A = U8 (0 << 7), B = S8 (1 << 10), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (2u << 4) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long t0 = clock64();
    // one election per K = 1024 step, then the 32 UMMAs as straight-line code of the elected lane: addresses and
    // descriptors are base + compile-time constants
    const uint64_t desc0 = make_desc_k_sw128(base);
    for (int r = 0; r < reps; ++r) {
      uint32_t elected;
      asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(elected));
      if (elected) {
#pragma unroll
        for (int ks = 0; ks < K / 32; ++ks) {
          const uint64_t bdesc = desc0 + (uint64_t)((ks / 4) * (N * 128 / 16) + 2 * (ks % 4));
          const uint32_t acc = (r > 0 || ks > 0) ? 1u : 0u;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                       "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                       ::"r"(tmem + d_col), "r"(tmem + a_col + 8 * ks), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bars[1]), 0);
    const long long t2 = clock64();
    if (lane == 0) {
      cyc[0] = t1 - t0;
      cyc[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N>
int run(int reps, uint32_t d_col = 256) {
  std::vector<uint16_t> bits((size_t)M * K / 16);
  std::vector<int8_t> B((size_t)N * K);
  std::vector<uint8_t> img((size_t)N * K);
  srand(1234 + N);
  for (auto& b : bits) b = (uint16_t)(rand() & rand() & 0xffff);   // ~25 % density
  for (auto& v : B) v = (int8_t)((rand() % 256) - 128);
  // logical K position kappa <-> (word p = kappa / 16, within: position 4 j + e <-> bit j + 4 e)
  auto spike = [&](int row, int kappa) {
    const int p = kappa / 16, q = kappa % 16, j = q / 4, e = q % 4;
    return (bits[(size_t)row * (K / 16) + p] >> (j + 4 * e)) & 1;
  };
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const int kb = k / 128, kk = k % 128;
      img[(size_t)kb * N * 128 + n * 128 + (((kk / 16) ^ (n & 7)) * 16) + kk % 16] = (uint8_t)B[(size_t)n * K + k];
    }
  std::vector<int> ref((size_t)M * N, 0);
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      long long s = 0;
      for (int k = 0; k < K; ++k) s += spike(r, k) * (int)B[(size_t)n * K + k];
      ref[(size_t)r * N + n] = (int)s;
    }
  uint16_t* dbits; uint8_t* dimg; int* dd; long long* dc;
  CK(cudaMalloc(&dbits, bits.size() * 2)); CK(cudaMalloc(&dimg, img.size())); CK(cudaMalloc(&dd, ref.size() * 4)); CK(cudaMalloc(&dc, 16));
  CK(cudaMemcpy(dbits, bits.data(), bits.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  const size_t smem = (size_t)N * K + 1024;
  CK(cudaFuncSetAttribute(umma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int pass = 0; pass < 2; ++pass) {
    const int rp = pass == 0 ? 1 : reps;
    CK(cudaMemset(dd, 0xff, ref.size() * 4));
    umma_kernel<N><<<1, 160, smem>>>(dbits, dimg, dd, dc, rp, d_col);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> out(ref.size());
    long long cyc[2];
    CK(cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost));
    long long bad = 0, first = -1;
    for (size_t i = 0; i < out.size(); ++i)
      if (out[i] != ref[i] * rp) { if (first < 0) first = (long long)i; ++bad; }
    printf("i8 TS  M=128 N=%3d K=1024 D at column %3u reps=%3d: mismatches %lld of %zu%s | issue %.1f cyc/UMMA, to completion %.1f cyc/UMMA (%lld cycles per K=1024 step)\n",
           N, d_col, rp, bad, out.size(), bad ? " (FAIL)" : " (exact)", (double)cyc[0] / (rp * 32), (double)cyc[1] / (rp * 32), cyc[1] / rp);
    if (bad && pass == 0) printf("   first mismatch at row %lld col %lld: got %d want %d\n", first / N, first % N, out[first], ref[first]);
  }
  cudaFree(dbits); cudaFree(dimg); cudaFree(dd); cudaFree(dc);
  return 0;
}

int main() {
  if (run<32>(50)) return 1;
  if (run<48>(50)) return 1;
  if (run<64>(50)) return 1;
  if (run<96>(50)) return 1;
  if (run<48>(50, 384)) return 1;
  if (run<48>(50, 448)) return 1;
  if (run<48>(1, 448)) return 1;
  return 0;
}

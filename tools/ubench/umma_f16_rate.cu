// Issue rate of tcgen05.mma kind::f16 (fp16 x fp16 -> fp32) for the shapes of the reverse recurrence
// (csrc/recur_tc.cu): M = 128, N in {32, 64, 128, 256}, K = 16 per instruction, A in TENSOR MEMORY or in shared
// memory, one or two accumulators.  Timing only (operands are whatever the memories hold).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o umma_f16_rate.bin umma_f16_rate.cu
#include <cstdio>
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

// MODE 0: A in TMEM, one accumulator; 1: A in TMEM, two accumulators alternating; 2: A in shared memory, one
// accumulator; 3: A in TMEM, kind::i8 (K = 32) for reference
template <int N, int MODE>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* __restrict__ cyc, int reps) {
  extern __shared__ unsigned char sm_raw[];
  const uint32_t raw = smem_u32(sm_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(sm_raw + (base - raw))[i] = 0x3c003c00u;  // 1.0
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = MODE == 3 ? ((2u << 4) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                                     : ((1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
    const uint64_t bdesc0 = make_desc_k_sw128(base), adesc0 = make_desc_k_sw128(base + 32768);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint32_t elected;
      asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(elected));
      if (elected) {
#pragma unroll
        for (int ks = 0; ks < 48; ++ks) {   // 48 UMMAs = one step of the reverse recurrence at H = 1024
          const uint64_t bdesc = bdesc0 + (uint64_t)(((ks / 4) % 4) * (N * 128 / 16) + 2 * (ks % 4));
          const uint32_t d = tmem + 256 + ((MODE == 1) ? 64u * (ks & 1) : 0u);
          if (MODE == 2) {
            const uint64_t adesc = adesc0 + (uint64_t)(((ks / 4) % 2) * 1024 + 2 * (ks % 4));
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                         ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          } else if (MODE == 3) {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                         ::"r"(d), "r"(tmem + 8 * (ks % 32)), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          } else {
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                         ::"r"(d), "r"(tmem + 8 * (ks % 32)), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int MODE>
int run(const char* what) {
  long long* dc;
  CK(cudaMalloc(&dc, 16));
  const size_t smem = 96 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int reps : {1, 20}) {
    rate_kernel<N, MODE><<<1, 64, smem>>>(dc, reps);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc[2];
    CK(cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost));
    printf("%-34s M=128 N=%3d reps=%2d: issue %.1f cyc/UMMA, to completion %.1f cyc/UMMA\n", what, N, reps,
           (double)cyc[0] / (48.0 * reps), (double)cyc[1] / (48.0 * reps));
  }
  cudaFree(dc);
  return 0;
}

int main() {
  if (run<64, 0>("f16, A in TMEM, one accumulator")) return 1;
  if (run<64, 1>("f16, A in TMEM, two accumulators")) return 1;
  if (run<64, 2>("f16, A in shared memory")) return 1;
  if (run<64, 3>("i8,  A in TMEM, one accumulator")) return 1;
  if (run<32, 0>("f16, A in TMEM, one accumulator")) return 1;
  if (run<32, 1>("f16, A in TMEM, two accumulators")) return 1;
  if (run<32, 3>("i8,  A in TMEM, one accumulator")) return 1;
  if (run<128, 0>("f16, A in TMEM, one accumulator")) return 1;
  if (run<128, 2>("f16, A in shared memory")) return 1;
  return 0;
}

// How fast does an SM pull 64 KB out of the L2 when 128 SMs do the same at once?  The all-gather of the tcgen05 BPTT
// (csrc/recur_tc.cu): per step every CTA reads 64 rows x 1 KB (its K quarter of dI_{t+1}) that 7 other CTAs read too.
// grid 128 x 256 threads, each thread 16 x 16-byte loads (the kernel's pattern: a half-warp = 256 contiguous bytes of a
// row), strong (ld.relaxed.gpu) or weak, all in flight; cycles from first issue to last use, per CTA (max over warps).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o l2_gather.bin l2_gather.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <int STRONG>
__global__ void __launch_bounds__(256) gather_kernel(const uint32_t* __restrict__ panel, int Hp, int share, int kbytes_per_thread,
                                                     long long* __restrict__ cyc, uint32_t* __restrict__ sink, int reps) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // CTA b reads quarter (b % 4) of group (b / 32 if share else b): 8 CTAs share a (group, quarter) when share = 1
  const int quarter = blockIdx.x % 4, group = share ? blockIdx.x / 32 : blockIdx.x / 4;
  uint32_t acc = 0;
  long long best = 1LL << 60, tot = 0;
  for (int r = 0; r < reps; ++r) {
    // a different panel every repetition: nothing of it is in this SM's L1 (32 groups of 256 KB rotate)
    const uint32_t* src = panel + ((size_t)((group + 5 * r) % 32) * 64 + 8 * warp + (lane >> 4)) * Hp + quarter * (Hp / 4) + 4 * (lane & 15);
    __syncthreads();
    const long long t0 = clock64();
    uint4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < kbytes_per_thread) {
        const uint32_t* a = src + (2 * (i & 3)) * Hp + (i >> 2) * 64;
        if (STRONG == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(a) : "memory");
        else if (STRONG == 2) asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(a) : "memory");
        else if (STRONG == 3) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(a) : "memory");
        else asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(a) : "memory");
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < kbytes_per_thread) acc += v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    __syncthreads();
    const long long dt = clock64() - t0;
    if (r > 0) { best = dt < best ? dt : best; tot += dt; }
  }
  if (tid == 0) { cyc[2 * blockIdx.x] = best; cyc[2 * blockIdx.x + 1] = tot / (reps - 1); }
  if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
  const int Hp = 1024, groups = 32;
  uint32_t* panel; long long* dc; uint32_t* sink;
  CK(cudaMalloc(&panel, (size_t)groups * 64 * Hp * 4)); CK(cudaMemset(panel, 1, (size_t)groups * 64 * Hp * 4));
  CK(cudaMalloc(&dc, 2 * 148 * 8)); CK(cudaMalloc(&sink, 4));
  long long h[2 * 148];
  const char* names[4] = {"ld.global.nc.L1::no_allocate", "ld.relaxed.gpu            ", "ld.global (weak)          ", "ld.volatile.global        "};
  CK(cudaFuncSetAttribute(gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int smem_kb : {0, 100, 200}) {
    for (int n : {16})
      for (int grid : {128}) {
        gather_kernel<1><<<grid, 256, smem_kb * 1024>>>(panel, Hp, 1, n, dc, sink, 50);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, dc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int b = 0; b < grid; ++b) avg += h[2 * b + 1];
        avg /= grid;
        printf("ld.relaxed.gpu, %3d KB of dynamic shared memory per CTA, 64 KB per CTA, 128 CTAs: avg %6.0f cycles = %5.1f B/clk/SM\n", smem_kb, avg, n * 4096.0 / avg);
      }
  }
  for (int strong = 0; strong < 4; ++strong)
    for (int share = 1; share < 2; ++share)
      for (int n : {4, 16})
        for (int grid : {1, 128}) {
          if (strong == 1) gather_kernel<1><<<grid, 256>>>(panel, Hp, share, n, dc, sink, 50);
          else if (strong == 2) gather_kernel<2><<<grid, 256>>>(panel, Hp, share, n, dc, sink, 50);
          else if (strong == 3) gather_kernel<3><<<grid, 256>>>(panel, Hp, share, n, dc, sink, 50);
          else gather_kernel<0><<<grid, 256>>>(panel, Hp, share, n, dc, sink, 50);
          CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(h, dc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
          double avg = 0, mn = 1e18, mx = 0;
          for (int b = 0; b < grid; ++b) { avg += h[2 * b + 1]; mn = h[2 * b] < mn ? h[2 * b] : mn; mx = h[2 * b + 1] > mx ? h[2 * b + 1] : mx; }
          avg /= grid;
          printf("%s, %s, %2d KB per CTA, %3d CTAs: avg %6.0f cycles (best %5.0f, slowest CTA avg %6.0f) = %5.1f B/clk/SM\n",
                 names[strong], share ? "8 CTAs share a panel" : "private panels     ", n * 4, grid, avg, mn, mx, n * 4096.0 / avg);
        }
  return 0;
}

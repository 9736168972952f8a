// Issue-rate microbenchmark of the legacy tensor path on sm_100a: mma.sync m16n8k16 f16 (HMMA) vs m16n8k32 s8 (IMMA).
// One CTA per SM, W warps per CTA, each warp issues N independent-accumulator MMAs in a loop; prints cycles per MMA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void rate_kernel(long long* out, int iters) {
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
  float f[8][4] = {};
  int c[8][4] = {};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(f[j][0]), "+f"(f[j][1]), "+f"(f[j][2]), "+f"(f[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  long long t1 = clock64();
  float s = 0; int si = 0;
  for (int j = 0; j < 8; ++j) for (int k = 0; k < 4; ++k) { s += f[j][k]; si += c[j][k]; }
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)s + si; }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int kind = 0; kind < 2; ++kind)
    for (int warps : {4, 8, 16}) {
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) {
        if (kind == 0) rate_kernel<0><<<148, warps * 32>>>(d, iters); else rate_kernel<1><<<148, warps * 32>>>(d, iters);
        cudaDeviceSynchronize();
      }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      double per = (double)h[0] / (iters * 8.0 * warps);
      printf("%s warps/SM %2d: %.2f cycles per MMA per SM (%.0f MAC/clk/SM)  err=%s\n", kind ? "IMMA m16n8k32 s8 " : "HMMA m16n8k16 f16",
             warps, per, (kind ? 8192.0 : 4096.0) / per, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}

// Latency floor of the recurrence's per-timestep hand-over ("measured empty-kernel exchange round
// trip"): an EMPTY persistent kernel with the forward recurrence's grid and its tagged-word exchange -- every
// CTA (row group g, neuron slice s) publishes one 32-bit word {16 payload bits, 16-bit step tag} per batch row and
// step, and may only publish step t+1 after it has seen the step-t words of ALL slices of its row group.  No tensor
// work, no neuron update: what remains is store -> L2 -> polling load, i.e. the chain no kernel of this
// decomposition can beat.  Prints ns per step for the forward geometry (words polled directly: one L2 round trip)
// and for a release/acquire flag hand-over (the reverse kernel's: stores, red.release, ld.acquire, then data).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o exchange_rtt.bin exchange_rtt.cu
// usage: exchange_rtt.bin [slices=64] [groups=2] [steps=2000]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ROWS = 128;

// words: [T][groups][slices][ROWS].  DEPTH = how many of a thread's polling loads are in flight at once: 16 = the
// slices are polled in dependent batches of 16 (4 L2 round trips per step at 64 slices), 64 = all at once.
template <int DEPTH>
__global__ void __launch_bounds__(ROWS, 1) tagged_kernel(uint32_t* __restrict__ words, int T, int* __restrict__ sink) {
  const int slice = blockIdx.x, group = blockIdx.y, nsl = gridDim.x, row = threadIdx.x;
  uint32_t acc = 0;
  for (int t = 0; t < T; ++t) {
    uint32_t* base = words + ((size_t)t * gridDim.y + group) * nsl * ROWS;
    const uint32_t tag = (uint32_t)((t + 1) & 0xffff) << 16;
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(base + (size_t)slice * ROWS + row), "r"(tag | (acc & 0xffffu)) : "memory");
    const long long t0 = clock64();
    for (int b0 = 0; b0 < nsl; b0 += DEPTH) {
      uint32_t v[DEPTH];
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) {
          const int s = (b0 + i + slice) % nsl;  // every consumer starts at a different producer
          v[i] = tag;
          if (b0 + i < nsl)
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[i]) : "l"(base + (size_t)s * ROWS + row) : "memory");
        }
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) ok = ok && ((v[i] ^ tag) >> 16) == 0;
        ok = __all_sync(0xffffffffu, ok);
        if (!ok && clock64() - t0 > 2000000000LL) __trap();
      } while (!ok);
#pragma unroll
      for (int i = 0; i < DEPTH; ++i) acc += v[i];
    }
    __syncthreads();  // stands for the join of the step (MMA complete -> update)
  }
  if (acc == 0x12345678u) sink[0] = 1;
}

// data + flag: every CTA writes its 128 words, then one release-add on the group's counter; consumers acquire-poll the
// counter (>= slices * (t+1)) and then read the data.
__global__ void __launch_bounds__(ROWS, 1) flag_kernel(uint32_t* __restrict__ words, int* __restrict__ ctr, int T, int* __restrict__ sink) {
  const int slice = blockIdx.x, group = blockIdx.y, nsl = gridDim.x, row = threadIdx.x;
  uint32_t acc = 0;
  for (int t = 0; t < T; ++t) {
    uint32_t* base = words + ((size_t)(t & 1) * gridDim.y + group) * nsl * ROWS;
    base[(size_t)slice * ROWS + row] = acc + t;
    __syncthreads();
    if (row == 0) {
      asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr + group) : "memory");
      const int target = nsl * (t + 1);
      const long long t0 = clock64();
      while (true) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr + group) : "memory");
        if (v >= target) break;
        if (clock64() - t0 > 2000000000LL) __trap();
      }
    }
    __syncthreads();
    for (int s = 0; s < nsl; ++s) acc += __ldcg(base + (size_t)s * ROWS + row);
    __syncthreads();
  }
  if (acc == 0x12345678u) sink[0] = 1;
}

int main(int argc, char** argv) {
  const int slices = argc > 1 ? atoi(argv[1]) : 64, groups = argc > 2 ? atoi(argv[2]) : 2, T = argc > 3 ? atoi(argv[3]) : 2000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  if (slices * groups > prop.multiProcessorCount) { printf("grid %d x %d exceeds %d SMs\n", slices, groups, prop.multiProcessorCount); return 1; }
  uint32_t* words; int *ctr, *sink;
  const size_t nw = (size_t)T * groups * slices * ROWS;
  CK(cudaMalloc(&words, nw * 4)); CK(cudaMalloc(&ctr, 4 * groups)); CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int Tk = T;
  for (int depth : {16, 64}) {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(words, 0, nw * 4));
      void* args[] = {(void*)&words, (void*)&Tk, (void*)&sink};
      const void* fn = depth == 16 ? (const void*)tagged_kernel<16> : (const void*)tagged_kernel<64>;
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel(fn, dim3(slices, groups), dim3(ROWS), args, 0, 0));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("tagged words (%2d loads in flight) grid %d x %d, %d steps: %.3f ms = %.1f ns/step\n", depth, slices, groups, T, ms,
             ms * 1e6 / T);
    }
  }
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(ctr, 0, 4 * groups));
    void* args[] = {(void*)&words, (void*)&ctr, (void*)&Tk, (void*)&sink};
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((const void*)flag_kernel, dim3(slices, groups), dim3(ROWS), args, 0, 0));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("release/acquire grid %d x %d, %d steps: %.3f ms = %.1f ns/step\n", slices, groups, T, ms, ms * 1e6 / T);
  }
  return 0;
}

// The reference draws the initial states of every forward pass from torch's default CPU generator
// (torch.rand(B, H), sparch/models/snns.py:286-287, 423-425, 558-559, 700-702, 812): a serial MT19937 stream, ~1 ms of
// host time per (256, 1024) draw, 7 draws per cfg-4 step.  This kernel produces the SAME numbers on the device from
// the generator's state and returns the state the generator would have afterwards, so the host only ships 2.5 KB each
// way and sets the state (sparch_b200/rng.py).  Algorithm (restated, not linked: ATen's at::mt19937 /
// uniform_real_distribution<float>, torch 2.x): the standard MT19937 recurrence and tempering (Matsumoto & Nishimura
// 1998); a float32 uniform in [0, 1) is (raw & (2^24 - 1)) * 2^-24, one 32-bit word per element, in element order.
#include "common.cuh"

namespace sparch {

constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ float mt_uniform(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return (float)(y & 0x00ffffffu) * 5.9604644775390625e-8f;   // exact: 24-bit integer times 2^-24
}

// One CTA.  The recurrence x[k + 624] = x[k + 397] ^ twist(x[k], x[k + 1]) reaches back 227 words, so a block of 624 words is
// 227 independent chains of three words plus one last word (old block in one buffer, new block in the other), then 624
// outputs at once.
__global__ void __launch_bounds__(256) mt19937_uniform_kernel(const uint32_t* __restrict__ state_in, int pos, long long n,
                                                              float* __restrict__ out, uint32_t* __restrict__ state_out) {
  __shared__ uint32_t buf[2][MT_N];
  const int tid = threadIdx.x;
  for (int i = tid; i < MT_N; i += 256) buf[0][i] = state_in[i];
  __syncthreads();
  int cur = 0;
  long long done = 0;
  {  // what is left of the block in hand
    const int take = (int)((long long)(MT_N - pos) < n ? (MT_N - pos) : n);
    for (int i = tid; i < take; i += 256) out[i] = mt_uniform(buf[0][pos + i]);
    done = take;
    pos += take;
  }
  while (done < n) {
    const uint32_t* o = buf[cur];
    uint32_t* w = buf[cur ^ 1];
    // words i, i + 227, i + 454 form a chain inside one thread (new[i + 227] needs new[i], the old words are all there):
    // only the last word needs other threads' results
    if (tid < MT_N - MT_M) {
      const int i1 = tid + (MT_N - MT_M), i2 = tid + 2 * (MT_N - MT_M);
      const uint32_t a = o[tid + MT_M] ^ mt_twist(o[tid], o[tid + 1]);                            // 0 .. 226
      const uint32_t b = a ^ mt_twist(o[i1], o[i1 + 1]);                                          // 227 .. 453
      w[tid] = a;
      w[i1] = b;
      if (i2 < MT_N - 1) w[i2] = b ^ mt_twist(o[i2], o[i2 + 1]);                                  // 454 .. 622
    }
    __syncthreads();
    if (tid == 0) w[MT_N - 1] = w[MT_M - 1] ^ mt_twist(o[MT_N - 1], w[0]);                        // 623
    __syncthreads();
    cur ^= 1;
    const int take = (int)(n - done < MT_N ? n - done : MT_N);
    for (int i = tid; i < take; i += 256) out[done + i] = mt_uniform(w[i]);
    done += take;
    pos = take;
  }
  __syncthreads();
  for (int i = tid; i < MT_N; i += 256) state_out[i] = buf[cur][i];
  if (tid == 0) state_out[MT_N] = (uint32_t)pos;
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_mt19937_uniform(const uint32_t* state_in, int pos, int64_t n, float* out, uint32_t* state_out,
                           sparch_stream_t st) {
  SPARCH_REQUIRE(state_in && state_out && pos >= 0 && pos <= MT_N && n >= 0 && (n == 0 || out), "bad argument");
  mt19937_uniform_kernel<<<1, 256, 0, as_stream(st)>>>(state_in, pos, (long long)n, out, state_out);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

// On-device data path of the spiking datasets: what sparch/dataloaders/spiking_datasets.py:66-78
// does per example on the host -- np.digitize of the firing times into nb_steps bins, a sparse (nb_steps, nb_units)
// tensor of ones at (bin, unit), to_dense() (duplicates SUM: the dense tensor holds spike counts) -- for a whole
// batch of event lists in one launch.  The batch crosses PCIe / NVLink-C2C as events (6-8 bytes each, ~8 k per SHD
// example) instead of as a dense fp32 tensor (280 KB per example).
#include "common.cuh"

namespace sparch {

// one thread per event: bin = #{i : bins[i] <= t} (np.digitize, right = False, increasing bins) by binary search over the
// SAME float64 bin edges the host would use, compared in float64 like numpy does; counts accumulate with atomicAdd on
// floats holding small integers (exact, order-independent)
__global__ void events_to_dense_kernel(const float* __restrict__ times, const int* __restrict__ units,
                                       const long long* __restrict__ offsets, const double* __restrict__ bins, int B,
                                       int nb_steps, int nb_units, long long nev, float* __restrict__ dense,
                                       int* __restrict__ bad) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nev; e += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = B;                       // example of this event: offsets[b] <= e < offsets[b + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (offsets[mid] <= e) lo = mid; else hi = mid;
    }
    const double t = (double)times[e];
    int a = 0, b = nb_steps;                  // bins[0 .. a) <= t < bins[b ..)
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (bins[mid] <= t) a = mid + 1; else b = mid;
    }
    const int u = units[e];
    if (a >= nb_steps || u < 0 || u >= nb_units || !(t == t)) {   // the reference's sparse constructor raises here
      atomicOr(bad, 1);
      continue;
    }
    atomicAdd(dense + ((long long)lo * nb_steps + a) * nb_units + u, 1.0f);
  }
}

}  // namespace sparch

using namespace sparch;

extern "C" {

int sparch_events_to_dense(const float* times, const int* units, const int64_t* offsets, const double* bins, int B,
                             int nb_steps, int nb_units, int64_t nev, float* dense, int* bad, sparch_stream_t st) {
  SPARCH_REQUIRE(B >= 0 && nb_steps > 0 && nb_units > 0 && nev >= 0, "bad shape");
  if (B == 0) return SPARCH_OK;
  SPARCH_REQUIRE(offsets && bins && dense && bad && (nev == 0 || (times && units)), "null pointer");
  cudaStream_t s = as_stream(st);
  SPARCH_CUDA(cudaMemsetAsync(dense, 0, (size_t)B * nb_steps * nb_units * sizeof(float), s));
  SPARCH_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
  if (nev == 0) return SPARCH_OK;
  long long nb = (nev + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  events_to_dense_kernel<<<(unsigned)(nb < cap ? nb : cap), 256, 0, s>>>(times, units, reinterpret_cast<const long long*>(offsets),
                                                                        bins, B, nb_steps, nb_units, nev, dense, bad);
  SPARCH_LAUNCH_OK();
  return SPARCH_OK;
}

}  // extern "C"

// alps_b200: the collective step of the HARMONIC partition inside a device group (api.cu, group_harmonic_eval).
//
// Every device of the group has left the un-normalised chi partials of its harmonic shard -- [omega][species][48
// complex] -- in its own HBM.  Device 0 sums them: dst[i] += src_1[i] + ... + src_k[i], reading the peers' rows
// directly through peer memory (NVLink 5 / NVSwitch loads, cudaDeviceEnablePeerAccess), in device order, so the sum is
// deterministic.  The payload is 768 B per species and omega: a latency-bound exchange, one launch, no staging copy --
// this replaces the MPI_REDUCE pair of disp() (src/ALPS_fns.f90:519-523).
#include "common.cuh"
#include "kernels.h"

namespace alps {

struct PeerRows {
  const double* p[8];
};

__global__ void __launch_bounds__(256) k_reduce_partials(double* __restrict__ dst, PeerRows src, int nsrc, size_t count2) {
  // double2 granularity: rows are 96 doubles long
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count2) return;
  double2 acc = reinterpret_cast<const double2*>(dst)[i];
  double2 v[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (k < nsrc) v[k] = __ldcv(reinterpret_cast<const double2*>(src.p[k]) + i);    // all peer loads in flight together
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (k < nsrc) {
      acc.x += v[k].x;
      acc.y += v[k].y;
    }
  reinterpret_cast<double2*>(dst)[i] = acc;
}

void launch_reduce_partials(double* dst, const double* const* src, int nsrc, size_t count, cudaStream_t st) {
  if (nsrc <= 0 || count == 0) return;
  PeerRows r{};
  for (int k = 0; k < nsrc && k < 8; k++) r.p[k] = src[k];
  const size_t c2 = count / 2;      // count = omegas * nspec * 96: even
  k_reduce_partials<<<(unsigned)((c2 + 255) / 256), 256, 0, st>>>(dst, r, nsrc, c2);
}

}  // namespace alps

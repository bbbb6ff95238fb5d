// Shared by the tensor-core depthwise kernels (dw_tc.cu, dw_tc_wgrad.cu): how a persistent CTA walks its work.
#pragma once
#include "kdcc_common.cuh"

namespace kdcc {

// Work is organised in units = (channel, split): the planes (image x 128x128 tile) of one channel, divided
// among `splits` CTAs.  A CTA keeps per-channel state (the Toeplitz operand, the k*k gradient partials) across
// the planes of a unit.  Every warp role walks the same sequence.
struct PlaneWalk {
  long pair, pairs, stride;
  int pl, planes, splits, C;
  __device__ PlaneWalk(long pairs_, int planes_, int splits_, int C_)
      : pair(blockIdx.x), pairs(pairs_), stride(gridDim.x), planes(planes_), splits(splits_), C(C_) {
    pl = (int)(pair / C);
    settle();
  }
  __device__ void settle() {  // skip units whose split owns no plane
    while (pair < pairs && pl >= planes) { pair += stride; pl = (int)(pair / C); }
  }
  __device__ bool valid() const { return pair < pairs; }
  __device__ int channel() const { return (int)(pair % C); }
  __device__ int split() const { return (int)(pair / C); }
  __device__ bool first_of_unit() const { return pl == split(); }
  __device__ bool last_of_unit() const { return pl + splits >= planes; }
  __device__ void next() {
    pl += splits;
    if (pl >= planes) next_unit();
  }
  __device__ void next_unit() {
    pair += stride;
    pl = (int)(pair / C);
    settle();
  }
};

// CTAs per channel: balance the persistent grid without shredding units into single planes
static inline int tc_unit_splits(int C, int planes) {
  int best = 1;
  long best_cost = -1;
  for (int s = 1; s <= planes && s <= 16; ++s) {
    const long rounds = ceil_div<long>((long)C * s, kNumSMs);
    const long cost = rounds * ceil_div(planes, s) * 16 + rounds;  // planes per CTA dominate; small per-unit overhead
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

}  // namespace kdcc

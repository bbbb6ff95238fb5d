// Shared by the tensor-core depthwise kernels (dw_tc.cu, dw_tc_wgrad.cu): how a persistent CTA walks its work.
#pragma once
#include "kdcc_common.cuh"

namespace kdcc {

// Work is organised in units = (channel, split): the planes (image x 128x128 tile) of one channel, divided
// among `splits` CTAs.  A CTA keeps per-channel state (the Toeplitz operand, the k*k gradient partials) across
// the planes of a unit.  Every warp role walks the same sequence.
struct PlaneWalk {
  // 32-bit arithmetic: units = channels x splits is far below 2^31, and a 64-bit division by a run-time value is a
  // ~100-instruction subroutine that the single issuing thread of a kernel would pay once or twice per plane
  unsigned pair, pairs, stride, C;
  int pl, planes, splits, sp, ch;  // sp / ch: split and channel of the current unit
  __device__ PlaneWalk(long pairs_, int planes_, int splits_, int C_)
      : pair(blockIdx.x), pairs((unsigned)pairs_), stride(gridDim.x), C((unsigned)C_), planes(planes_), splits(splits_) {
    locate();
    settle();
  }
  __device__ void locate() {
    sp = (int)(pair / C);
    ch = (int)(pair - (unsigned)sp * C);
    pl = sp;
  }
  __device__ void settle() {  // skip units whose split owns no plane
    while (pair < pairs && pl >= planes) { pair += stride; locate(); }
  }
  __device__ bool valid() const { return pair < pairs; }
  __device__ int channel() const { return ch; }
  __device__ int split() const { return sp; }
  __device__ bool first_of_unit() const { return pl == sp; }
  __device__ bool last_of_unit() const { return pl + splits >= planes; }
  __device__ void next() {
    pl += splits;
    if (pl >= planes) next_unit();
  }
  __device__ void next_unit() {
    pair += stride;
    locate();
    settle();
  }
};

// One level of a warp-wide reduce-scatter over 2*H values per lane: the lane with bit M set keeps the upper half and
// hands over the lower one (and vice versa); afterwards a[0..H) holds pair sums and `base` the index of a[0].  Five
// levels (M = 16 .. 1) leave every lane with the complete sums of its own 1/32 of the values.
template <int H, int M, int P>
__device__ __forceinline__ void halve_scatter(float (&a)[P], int lane, int &base) {
  const bool up = (lane & M) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float send = up ? a[i] : a[i + H];
    const float keep = up ? a[i + H] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
  }
  if (up) base += H;
  if constexpr (M > 1) halve_scatter<H / 2, M / 2>(a, lane, base);
}

// out[u*K+v] = sum over the warp's lanes of acc[u][v] (the k*k tap sums of the rows a warp owns)
template <int K>
__device__ __forceinline__ void warp_sum_taps(const float (&acc)[K][K], int lane, float *out) {
  constexpr int KK = K * K, P = 32 * ((KK + 31) / 32);
  float a[P];
#pragma unroll
  for (int i = 0; i < P; ++i) a[i] = i < KK ? acc[i / K][i % K] : 0.f;
  int base = 0;
  halve_scatter<P / 2, 16>(a, lane, base);
#pragma unroll
  for (int i = 0; i < P / 32; ++i)
    if (base + i < KK) out[base + i] = a[i];
}

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

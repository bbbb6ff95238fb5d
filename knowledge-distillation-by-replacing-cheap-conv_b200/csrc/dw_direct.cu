// Depthwise k x k convolution, direct (gather) kernels: any k / dilation / padding, fp32 or bf16,
// NHWC with C a multiple of the 16-byte vector.  These are the parity path for fp32 (1e-5) and the
// general-shape path; the bf16 production shapes dispatch to the TMA-staged kernels in dw_tma.cu.
// Reference semantics: models/students/transform_blocks/depthwise_separable_conv.py:7-8,12
// (F.conv2d, groups=C, stride 1, zero padding) and its autograd backward.
#include "dw_kernels.cuh"

namespace kdcc {

// One thread = one output pixel x one 16-byte channel vector.  `flip` turns the same loop into the
// input-gradient (transposed) correlation: taps are read mirrored and pad' = dil*(k-1) - pad.
template <typename T>
__global__ void __launch_bounds__(256) dw_direct_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, T *__restrict__ y,
                                                        int N, int H, int W, int C, int Ho, int Wo, int k,
                                                        int dil, int pad, int flip) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  const int CV = C / VN;
  const long total = (long)N * Ho * Wo * CV;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV);
    long r = idx / CV;
    const int j = (int)(r % Wo); r /= Wo;
    const int i = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float acc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[e] = bias ? bias[cv * VN + e] : 0.f;
    for (int u = 0; u < k; ++u) {
      const int ii = i + u * dil - pad;
      if (ii < 0 || ii >= H) continue;
      for (int v = 0; v < k; ++v) {
        const int jj = j + v * dil - pad;
        if (jj < 0 || jj >= W) continue;
        V xv;
        xv.raw = *reinterpret_cast<const decltype(xv.raw) *>(x + (((long)n * H + ii) * W + jj) * C + cv * VN);
        float xf[VN];
        xv.unpack(xf);
        const int tap = flip ? (k - 1 - u) * k + (k - 1 - v) : u * k + v;
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] = fmaf(w[(long)(cv * VN + e) * k * k + tap], xf[e], acc[e]);
      }
    }
    V o;
    o.pack(acc);
    *reinterpret_cast<decltype(o.raw) *>(y + (((long)n * Ho + i) * Wo + j) * C + cv * VN) = o.raw;
  }
}

// Weight gradient, stage 1.  Thread = (channel vector, tap); blockIdx.y = split of the N*Ho output rows.
// part[split][tap][c], tap == k*k is the bias gradient (sum of dy).
template <typename T>
__global__ void __launch_bounds__(256) dw_wgrad_direct_kernel(const T *__restrict__ x, const T *__restrict__ dy,
                                                              float *__restrict__ part, int N, int H, int W,
                                                              int C, int Ho, int Wo, int k, int dil, int pad,
                                                              int rows_per_split) {
  using V = Vec16<T>;
  constexpr int VN = V::N;
  const int CV = C / VN;
  const int KK1 = k * k + 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= CV * KK1) return;
  const int cv = t % CV, tap = t / CV;
  const int u = tap / k, v = tap % k;  // tap == k*k -> u == k (bias row)
  const bool is_bias = tap == k * k;
  const int row0 = blockIdx.y * rows_per_split;
  const int row1 = min(N * Ho, row0 + rows_per_split);
  float acc[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) acc[e] = 0.f;
  for (int row = row0; row < row1; ++row) {
    const int n = row / Ho, i = row % Ho;
    const int ii = is_bias ? 0 : i + u * dil - pad;
    if (!is_bias && (ii < 0 || ii >= H)) continue;
    for (int j = 0; j < Wo; ++j) {
      const int jj = is_bias ? 0 : j + v * dil - pad;
      if (!is_bias && (jj < 0 || jj >= W)) continue;
      V gv;
      gv.raw = *reinterpret_cast<const decltype(gv.raw) *>(dy + (((long)n * Ho + i) * Wo + j) * C + cv * VN);
      float gf[VN];
      gv.unpack(gf);
      if (is_bias) {
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] += gf[e];
      } else {
        V xv;
        xv.raw = *reinterpret_cast<const decltype(xv.raw) *>(x + (((long)n * H + ii) * W + jj) * C + cv * VN);
        float xf[VN];
        xv.unpack(xf);
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[e] = fmaf(gf[e], xf[e], acc[e]);
      }
    }
  }
  float *dst = part + ((long)blockIdx.y * KK1 + tap) * C + cv * VN;
#pragma unroll
  for (int e = 0; e < VN; ++e) dst[e] = acc[e];
}

// stage 2: dw[c][tap] = sum_split part[split][tap][c]  (fixed order, double accumulate)
__global__ void dw_wgrad_reduce_kernel(const float *__restrict__ part, float *__restrict__ dw,
                                       float *__restrict__ dbias, int splits, int C, int KK, int rows) {
  // part[split][rows][C]; rows == KK (no bias row) or KK + 1 (row KK = sum of dy)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // c fastest
  if (idx >= rows * C) return;
  const int c = idx % C, tap = idx / C;
  double acc = 0.0;
  for (int s = 0; s < splits; ++s) acc += (double)part[((long)s * rows + tap) * C + c];
  if (tap < KK) {
    if (dw) dw[(long)c * KK + tap] = (float)acc;
  } else if (dbias) {
    dbias[c] = (float)acc;
  }
}

int dw_direct_wgrad_splits(int N, int Ho, int C, int k, int vn) {
  const long threads = (long)(C / vn) * (k * k + 1);
  const long blocks_x = ceil_div<long>(threads, 256);
  long want = ceil_div<long>((long)kNumSMs * 8, blocks_x);
  want = max(1L, min(want, (long)N * Ho));
  return (int)want;
}

template <typename T>
int dw_direct_fwd(const void *x, const float *w, const float *bias, void *y, int N, int H, int W, int C, int Ho,
                  int Wo, int k, int dil, int pad, int flip, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  const long total = (long)N * Ho * Wo * (C / VN);
  if (total == 0) return KDCC_OK;
  const int grid = (int)min((long)kNumSMs * 16, ceil_div<long>(total, 256));
  dw_direct_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T *>(x), w, bias, static_cast<T *>(y), N, H, W, C,
                                            Ho, Wo, k, dil, pad, flip);
  return launch_status();
}

template <typename T>
int dw_direct_wgrad(const void *x, const void *dy, float *dw, float *dbias, float *part, int N, int H, int W,
                    int C, int Ho, int Wo, int k, int dil, int pad, cudaStream_t st) {
  constexpr int VN = Vec16<T>::N;
  const int KK = k * k;
  const int splits = dw_direct_wgrad_splits(N, Ho, C, k, VN);
  const int rows_per = ceil_div(N * Ho, splits);
  const int threads = (C / VN) * (KK + 1);
  dim3 grid(ceil_div(threads, 256), splits);
  dw_wgrad_direct_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T *>(x), static_cast<const T *>(dy), part, N,
                                                  H, W, C, Ho, Wo, k, dil, pad, rows_per);
  int rc = launch_status();
  if (rc) return rc;
  dw_wgrad_reduce_kernel<<<ceil_div((KK + 1) * C, 256), 256, 0, st>>>(part, dw, dbias, splits, C, KK, KK + 1);
  return launch_status();
}

template int dw_direct_fwd<float>(const void *, const float *, const float *, void *, int, int, int, int, int, int,
                                  int, int, int, int, cudaStream_t);
template int dw_direct_fwd<__nv_bfloat16>(const void *, const float *, const float *, void *, int, int, int, int,
                                          int, int, int, int, int, int, cudaStream_t);
template int dw_direct_wgrad<float>(const void *, const void *, float *, float *, float *, int, int, int, int, int,
                                    int, int, int, int, cudaStream_t);
template int dw_direct_wgrad<__nv_bfloat16>(const void *, const void *, float *, float *, float *, int, int, int,
                                            int, int, int, int, int, int, cudaStream_t);

}  // namespace kdcc

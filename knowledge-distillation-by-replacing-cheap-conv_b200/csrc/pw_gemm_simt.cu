// CUDA-core GEMM for the pointwise convolution: the fp32 parity path (1e-5 relative against the
// reference needs true fp32 products) and the landing spot for shapes the tcgen05 kernel refuses
// (channel counts that are not multiples of 8).  64x64 output tile, 16-deep reduction slices,
// 4x4 register micro-tile per thread.  Reference: depthwise_separable_conv.py:9,13 and autograd.
#include "pw_kernels.cuh"

namespace kdcc {

constexpr int ST = 64, SK = 16;

template <typename T>
__global__ void __launch_bounds__(256) pw_simt_kernel(const T *__restrict__ A, const T *__restrict__ B, const SimtGemm g) {
  __shared__ float As[SK][ST + 4];
  __shared__ float Bs[SK][ST + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * ST, j0 = blockIdx.x * ST;
  const int split = blockIdx.z;
  const int per = ((g.R + g.splits - 1) / g.splits + SK - 1) / SK * SK;
  const int r_begin = split * per, r_end = min(g.R, r_begin + per);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const bool a_r_contig = g.sar == 1, b_r_contig = g.sbr == 1;
  for (int r0 = r_begin; r0 < r_end; r0 += SK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * 256;
      {
        const int rr = a_r_contig ? e % SK : e / ST, ii = a_r_contig ? e / SK : e % ST;
        const int gi = i0 + ii, gr = r0 + rr;
        As[rr][ii] = (gi < g.I && gr < r_end) ? to_f32(A[gi * g.sai + gr * g.sar]) : 0.f;
      }
      {
        const int rr = b_r_contig ? e % SK : e / ST, jj = b_r_contig ? e / SK : e % ST;
        const int gj = j0 + jj, gr = r0 + rr;
        Bs[rr][jj] = (gj < g.J && gr < r_end) ? to_f32(B[gj * g.sbj + gr * g.sbr]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SK; ++r) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[r][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[r][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + ty * 4 + a;
    if (gi >= g.I) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + tx * 4 + b;
      if (gj >= g.J) continue;
      const float v = acc[a][b];
      if (g.out_f32) {
        g.out_f32[((long)split * g.I + gi) * g.J + gj] = v;
      } else {
        if (g.out_raw) static_cast<T *>(g.out_raw)[(long)gi * g.J + gj] = from_f32<T>(v);
        if (g.out_act) {
          float x = v;
          if (g.scale) x *= g.scale[gj];
          if (g.shift) x += g.shift[gj];
          if (g.residual) x += to_f32(static_cast<const T *>(g.residual)[(long)gi * g.J + gj]);
          if (g.relu) x = fmaxf(x, 0.f);
          static_cast<T *>(g.out_act)[(long)gi * g.J + gj] = from_f32<T>(x);
        }
      }
    }
  }
}

template <typename T>
int pw_simt_gemm(const void *a, const void *b, const SimtGemm &g, cudaStream_t st) {
  dim3 grid(ceil_div(g.J, ST), ceil_div(g.I, ST), g.splits);
  pw_simt_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T *>(a), static_cast<const T *>(b), g);
  return launch_status();
}
template int pw_simt_gemm<float>(const void *, const void *, const SimtGemm &, cudaStream_t);
template int pw_simt_gemm<__nv_bfloat16>(const void *, const void *, const SimtGemm &, cudaStream_t);

int pw_simt_dw_splits(long M, int K, int Nc) {
  const long tiles = (long)ceil_div(K, ST) * ceil_div(Nc, ST);
  long s = max(1L, (long)kNumSMs * 4 / tiles);
  s = min(s, max(1L, M / 256));
  const long per = ceil_div<long>(ceil_div<long>(M, s), SK) * SK;
  return (int)ceil_div<long>(M, per);
}

}  // namespace kdcc

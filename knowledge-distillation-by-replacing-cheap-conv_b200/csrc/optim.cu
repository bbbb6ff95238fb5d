// The optimizer step of the layerwise loop as ONE pass: RAdam exactly as the reference implements it
// (utils/optim/radam.py:31-99; selected by cfg/cityscapes/*.json "optimizer": {"type": "RAdam", "lr": 0.005}).
//
// The reference issues ~10 elementwise torch kernels per parameter tensor (mul_, addcmul_, sqrt, add_, addcdiv_, copy_
// ...: 18 tensors for the 51M plan).  Here one kernel reads p, g, m, v once, writes p, m, v once, and can emit the bf16
// copy of the new weights that the pointwise GEMMs consume (which saves the separate cast pass of the next step).
// The step-dependent scalars (N_sma, step_size: radam.py:65-84) are host arithmetic in float64 as in the reference and
// arrive folded into `decay` = -weight_decay * lr and `step` = -step_size * lr.
#include "kdcc_common.cuh"

namespace kdcc {

constexpr int OPT_THREADS = 256;

struct RadamScalars {
  float beta1, beta2, one_m_beta1, one_m_beta2, eps, decay, step;
  int mode;  // 0: N_sma >= 5 (adaptive, :87-92), 1: degenerated to SGD with momentum (:93-97), 2: moments only (step_size < 0)
};

__device__ __forceinline__ void radam_one(float &p, float g, float &m, float &v, const RadamScalars &s) {
  v = v * s.beta2 + s.one_m_beta2 * g * g;   // :60
  m = m * s.beta1 + s.one_m_beta1 * g;       // :61
  if (s.mode == 2) return;
  if (s.decay != 0.f) p += s.decay * p;      // :89 / :95
  p += s.mode == 0 ? s.step * (m / (sqrtf(v) + s.eps)) : s.step * m;   // :90-91 / :96
}

// NSRC > 1: data-parallel form.  g points at `NSRC` copies of the gradient, `src_stride` floats apart -- one per rank, pushed
// into this rank's memory by its peers over NVLink while the backward was still running (kdcc.PeerGradBucket) -- and the
// step uses their mean, summed in rank order so that every rank computes bit-identical parameters.  The all-reduce is
// thereby fused into the optimizer pass: no collective kernel, no extra pass over the gradient.
template <int NSRC>
__global__ void __launch_bounds__(OPT_THREADS)
radam_kernel(float *__restrict__ p, const float *__restrict__ g, long src_stride, int n_src, float *__restrict__ m, float *__restrict__ v,
             __nv_bfloat16 *__restrict__ p_lp, long n, const RadamScalars s) {
  pdl_prologue_done();
  const long n4 = n / 4;
  const float inv = 1.f / (float)n_src;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
    float4 gg = __ldcs(reinterpret_cast<const float4 *>(g) + i);
    if (NSRC > 1) {
      for (int r = 1; r < n_src; ++r) {
        const float4 o = __ldcs(reinterpret_cast<const float4 *>(g + (long)r * src_stride) + i);
        gg.x += o.x; gg.y += o.y; gg.z += o.z; gg.w += o.w;
      }
      gg.x *= inv; gg.y *= inv; gg.z *= inv; gg.w *= inv;
    }
    radam_one(pp.x, gg.x, mm.x, vv.x, s); radam_one(pp.y, gg.y, mm.y, vv.y, s);
    radam_one(pp.z, gg.z, mm.z, vv.z, s); radam_one(pp.w, gg.w, mm.w, vv.w, s);
    reinterpret_cast<float4 *>(m)[i] = mm; reinterpret_cast<float4 *>(v)[i] = vv;
    if (s.mode != 2) reinterpret_cast<float4 *>(p)[i] = pp;
    if (p_lp) reinterpret_cast<uint2 *>(p_lp)[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
  }
  // ragged tail (n % 4 elements)
  const long i = n4 * 4 + (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float pp = p[i], mm = m[i], vv = v[i], gg = g[i];
    if (NSRC > 1) {
      for (int r = 1; r < n_src; ++r) gg += g[(long)r * src_stride + i];
      gg *= inv;
    }
    radam_one(pp, gg, mm, vv, s);
    m[i] = mm; v[i] = vv;
    if (s.mode != 2) p[i] = pp;
    if (p_lp) p_lp[i] = __float2bfloat16(pp);
  }
}

}  // namespace kdcc

using namespace kdcc;

static int radam_launch(float *p, const float *g, long src_stride, int n_src, float *m, float *v, void *p_lp, long n, float beta1,
                        float beta2, float one_minus_beta1, float one_minus_beta2, float eps, float decay, float step, int mode,
                        kdcc_stream_t stream) {
  if (n < 0 || mode < 0 || mode > 2 || n_src < 1) return KDCC_EINVAL;
  if (n == 0) return KDCC_OK;
  if (!p || !g || !m || !v) return KDCC_EINVAL;
  if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v) || (p_lp && (reinterpret_cast<uintptr_t>(p_lp) & 7)) ||
      (n_src > 1 && src_stride % 4 != 0))
    return KDCC_EALIGN;
  RadamScalars s{beta1, beta2, one_minus_beta1, one_minus_beta2, eps, decay, step, mode};
  const int grid = (int)min((long)kNumSMs * 8, ceil_div<long>(max(n / 4, 1L), OPT_THREADS));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16 *lp = static_cast<__nv_bfloat16 *>(p_lp);
  if (n_src == 1) launch_pdl(radam_kernel<1>, dim3(grid), dim3(OPT_THREADS), 0, st, p, g, src_stride, n_src, m, v, lp, n, s);
  else launch_pdl(radam_kernel<2>, dim3(grid), dim3(OPT_THREADS), 0, st, p, g, src_stride, n_src, m, v, lp, n, s);
  return launch_status();
}

KDCC_API int kdcc_radam_step(float *p, const float *g, float *m, float *v, void *p_lp, long n, float beta1, float beta2,
                             float one_minus_beta1, float one_minus_beta2, float eps, float decay, float step, int mode,
                             kdcc_stream_t stream) {
  return radam_launch(p, g, 0, 1, m, v, p_lp, n, beta1, beta2, one_minus_beta1, one_minus_beta2, eps, decay, step, mode, stream);
}

KDCC_API int kdcc_radam_step_multi(float *p, const float *g, long src_stride, int n_src, float *m, float *v, void *p_lp, long n,
                                   float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps, float decay,
                                   float step, int mode, kdcc_stream_t stream) {
  return radam_launch(p, g, src_stride, n_src, m, v, p_lp, n, beta1, beta2, one_minus_beta1, one_minus_beta2, eps, decay, step,
                      mode, stream);
}

// Segmentation metric of the training loop, on the device: per-pixel argmax over the class logits + confusion
// matrix against the labels, one HBM pass.
//
// Reference semantics: utils/util.py:108-128 (CityscapesMetricTracker.update / confusion_for_batch) and
// models/metric.py:6-18,49-55: labels equal to ignore_index (or outside [0, C)) are dropped,
// pred = argmax over dim 1 (first maximum wins), conf[target][pred] += 1.  The reference copies both full logit
// tensors to the host every iteration and bincounts in numpy; here nothing leaves the GPU until the
// IoU is asked for.  Integer counts: bit-exact and order independent.
#include "kdcc_common.cuh"

namespace kdcc {

constexpr int CONF_THREADS = 256;
constexpr int CONF_MAX_C = 32;
constexpr int CONF_OCC = 6;  // CTAs per SM: the grid-stride grid is exactly one resident wave

// logits (n, c, q) at s + n*batch_stride + c*class_stride + q (pixel stride 1: NCHW); labels int64 [N][HW]
template <typename T, int VEC>
__global__ void __launch_bounds__(CONF_THREADS, CONF_OCC)
confusion_kernel(const T *__restrict__ s, const long long *__restrict__ labels, unsigned long long *__restrict__ conf,
                 int N, int C, long HW, long batch_stride, long class_stride, int ignore_index) {
  extern __shared__ unsigned int hist[];  // [warps][C*C] per-warp counts (<= 32 KB)
  const int warp = threadIdx.x >> 5;
  const int cells = C * C;
  unsigned int *mine = hist + (size_t)warp * cells;
  for (int i = threadIdx.x & 31; i < cells; i += 32) mine[i] = 0u;
  __syncwarp();
  const long groups = HW / VEC;  // VEC consecutive pixels per thread and step
  const long total = (long)N * groups;
  for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const long n = g / groups, q = (g % groups) * VEC;
    const T *base = s + n * batch_stride + q;
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    for (int c = 0; c < C; ++c) {
      float v[VEC];
      if constexpr (VEC == 4 && sizeof(T) == 4) {
        const float4 f = __ldcs(reinterpret_cast<const float4 *>(base + (long)c * class_stride));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = to_f32(base[(long)c * class_stride + e]);
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        // first maximum wins; a NaN beats everything (torch.argmax semantics)
        const bool take = (c == 0) || (v[e] > best[e]) || (v[e] != v[e] && best[e] == best[e]);
        if (take) { best[e] = v[e]; arg[e] = c; }
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const long long t = __ldcs(labels + n * HW + q + e);
      if (t >= 0 && t < C && t != ignore_index) atomicAdd(&mine[(int)t * C + arg[e]], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cells; i += blockDim.x) {
    unsigned long long sum = 0;
    for (int w = 0; w < CONF_THREADS / 32; ++w) sum += hist[(size_t)w * cells + i];
    if (sum) atomicAdd(conf + i, sum);
  }
}

}  // namespace kdcc

using namespace kdcc;

KDCC_API int kdcc_confusion_update(const void *logits, const long long *labels, long long *conf, int N, int C, long HW,
                                   long batch_stride, long class_stride, int ignore_index, int dtype,
                                   kdcc_stream_t stream) {
  if (N < 0 || C <= 0 || HW < 0 || (dtype != KDCC_F32 && dtype != KDCC_BF16)) return KDCC_EINVAL;
  if (C > CONF_MAX_C) return KDCC_ESHAPE;
  if (N == 0 || HW == 0) return KDCC_OK;
  if (!logits || !labels || !conf) return KDCC_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cells = C * C;
  const size_t dyn = (size_t)(CONF_THREADS / 32) * cells * sizeof(unsigned int);
  auto *out = reinterpret_cast<unsigned long long *>(conf);
  const bool vec4 = dtype == KDCC_F32 && HW % 4 == 0 && batch_stride % 4 == 0 && class_stride % 4 == 0 && aligned16(logits);
  const long groups = (long)N * (vec4 ? HW / 4 : HW);
  const int grid = (int)min((long)kNumSMs * CONF_OCC, ceil_div<long>(groups, CONF_THREADS));
  if (vec4)
    confusion_kernel<float, 4><<<grid, CONF_THREADS, dyn, st>>>(static_cast<const float *>(logits), labels, out, N, C, HW,
                                                                batch_stride, class_stride, ignore_index);
  else if (dtype == KDCC_F32)
    confusion_kernel<float, 1><<<grid, CONF_THREADS, dyn, st>>>(static_cast<const float *>(logits), labels, out, N, C, HW,
                                                                batch_stride, class_stride, ignore_index);
  else
    confusion_kernel<__nv_bfloat16, 1><<<grid, CONF_THREADS, dyn, st>>>(static_cast<const __nv_bfloat16 *>(logits), labels, out, N,
                                                                        C, HW, batch_stride, class_stride, ignore_index);
  return launch_status();
}

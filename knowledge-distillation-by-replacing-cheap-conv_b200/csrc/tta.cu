// Sliding-window test-time inference, on the device: overlap-add of the window outputs, un-mirroring of the flipped
// pass, bilinear resize to the original size and the average over passes and scales.
//
// Reference semantics: utils/tta_process.py:9-52 (reverse_mapping / resize_output / collect_windows_result), driven
// by models/students/depthwise_student.py:187-206 (inference_test), where the window outputs are copied to the host
// and stitched with numpy + cv2 in float64.  Here the windows stay where the student wrote them; both kernels are one
// pass over their output (HBM-bound: every output element gathers the 1-4 windows that cover it).
//
// The reference divides the overlap-add by a counter that it indexes [y1:y2, x1:x2] on a (classes, h, w) array
// (tta_process.py:46): classes are selected by the window's y range and ROWS by its x range.  count_mode 0 reproduces
// that counter exactly (results identical to the reference, including its inf / nan where the counter is 0);
// count_mode 1 divides by the per-pixel coverage, which is what the overlap-add needs.
#include "kdcc_common.cuh"

namespace kdcc {

constexpr int TTA_THREADS = 256;
constexpr int TTA_MAX_WINDOWS = 512;

// out[c][y][x] (+)= alpha * stitched[c][y][flip ? w-1-x : x]
__global__ void __launch_bounds__(TTA_THREADS)
tta_stitch_kernel(const float *__restrict__ win, const int4 *__restrict__ coords, int n, int C, int th, int tw, int h,
                  int w, int flip, int count_mode, float alpha, float *__restrict__ out, int accumulate) {
  __shared__ int4 box[TTA_MAX_WINDOWS];  // (x1, y1, x2, y2)
  for (int i = threadIdx.x; i < n; i += blockDim.x) box[i] = coords[i];
  __syncthreads();
  const long total = (long)C * h * w;
  const long plane = (long)th * tw;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int x = (int)(o % w);
    const int y = (int)((o / w) % h);
    const int c = (int)(o / ((long)w * h));
    const int xs = flip ? w - 1 - x : x;
    float sum = 0.f;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
      const int4 b = box[i];
      const bool inside = xs >= b.x && xs < b.z && y >= b.y && y < b.w;
      if (inside) sum += __ldg(win + ((long)i * C + c) * plane + (long)(y - b.y) * tw + (xs - b.x));
      // reference counter: dim 0 (classes) sliced by [y1, y2), dim 1 (rows) by [x1, x2), every column
      cnt += count_mode == 0 ? (c >= b.y && c < b.w && y >= b.x && y < b.z) : inside;
    }
    const float v = alpha * (sum / (float)cnt);  // cnt == 0: inf / nan exactly like the numpy division
    out[o] = accumulate ? out[o] + v : v;
  }
}

// dst[c][Y][X] (+)= alpha * bilinear(src[c])(Y, X), cv2.INTER_LINEAR conventions (tta_process.py:29-36): source
// coordinate (d + 0.5) * in / out - 0.5 in float32, taps floor / floor + 1 clamped to the plane
__global__ void __launch_bounds__(TTA_THREADS)
resize_bilinear_kernel(const float *__restrict__ src, int C, int h, int w, float *__restrict__ dst, int H, int W,
                       float alpha, int accumulate) {
  const long total = (long)C * H * W;
  const double sy = (double)h / H, sx = (double)w / W;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int X = (int)(o % W);
    const int Y = (int)((o / W) % H);
    const int c = (int)(o / ((long)W * H));
    float fx = (float)((X + 0.5) * sx - 0.5), fy = (float)((Y + 0.5) * sy - 0.5);
    int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
    fx -= (float)x0; fy -= (float)y0;
    if (x0 < 0 || x0 >= w - 1) fx = 0.f;
    if (y0 < 0 || y0 >= h - 1) fy = 0.f;
    x0 = min(max(x0, 0), w - 1); y0 = min(max(y0, 0), h - 1);
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const float *p = src + (long)c * h * w;
    // a tap of weight 0 is not used (an inf there must not become nan through 0 * inf; same size = copy, like cv2)
    auto lerp = [](float a, float b, float f) { return f > 0.f ? a * (1.f - f) + b * f : a; };
    const float top = lerp(__ldg(p + (long)y0 * w + x0), __ldg(p + (long)y0 * w + x1), fx);
    const float bot = lerp(__ldg(p + (long)y1 * w + x0), __ldg(p + (long)y1 * w + x1), fx);
    const float v = alpha * lerp(top, bot, fy);
    dst[o] = accumulate ? dst[o] + v : v;
  }
}

static int stream_grid(long total) { return (int)min((long)kNumSMs * 8, ceil_div<long>(total, TTA_THREADS)); }

}  // namespace kdcc

using namespace kdcc;

KDCC_API int kdcc_tta_stitch(const float *windows, const int *coords, int n, int C, int th, int tw, int h, int w, int flip,
                             int count_mode, float alpha, float *out, int accumulate, kdcc_stream_t stream) {
  if (n < 0 || C <= 0 || th <= 0 || tw <= 0 || h <= 0 || w <= 0 || (count_mode != 0 && count_mode != 1)) return KDCC_EINVAL;
  if (n > TTA_MAX_WINDOWS) return KDCC_ESHAPE;
  if (!out || (n > 0 && (!windows || !coords))) return KDCC_EINVAL;
  if (!aligned16(coords)) return KDCC_EALIGN;
  const long total = (long)C * h * w;
  tta_stitch_kernel<<<stream_grid(total), TTA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      windows, reinterpret_cast<const int4 *>(coords), n, C, th, tw, h, w, flip, count_mode, alpha, out, accumulate);
  return launch_status();
}

KDCC_API int kdcc_resize_bilinear(const float *src, int C, int h, int w, float *dst, int H, int W, float alpha,
                                  int accumulate, kdcc_stream_t stream) {
  if (C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return KDCC_EINVAL;
  if (!src || !dst) return KDCC_EINVAL;
  const long total = (long)C * H * W;
  resize_bilinear_kernel<<<stream_grid(total), TTA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(src, C, h, w, dst, H, W,
                                                                                                  alpha, accumulate);
  return launch_status();
}

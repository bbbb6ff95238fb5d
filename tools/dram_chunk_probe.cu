// Micro-benchmark (measurement tooling, not product): HBM bandwidth as a function of the contiguous run a warp
// touches.  The NHWC depthwise kernels read and write 512-byte runs (256 channels of one pixel) that are 8 KB apart,
// the KD kernel 256-byte runs 4 MB apart; the copy peak in MEASURED_PEAKS.json is for fully sequential streams.
// Each warp copies runs of `chunk` bytes (16 bytes per lane and instruction); consecutive warps of the grid take runs
// that are `stride` bytes apart (stride == chunk: sequential).  Prints read + write GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dram_chunk_probe tools/dram_chunk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// run r lives at (r % runs_per_row) * stride + (r / runs_per_row) * chunk  (a transposed walk over a [stride/chunk][...] grid)
__global__ void __launch_bounds__(256) copy_runs(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long runs, int chunk,
                                                 long stride, long rows) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const int vec_per_run = chunk / 16;
  for (long r = warp; r < runs; r += nwarps) {
    const long row = r % rows, col = r / rows;          // consecutive warps -> consecutive rows -> `stride` apart
    const long base = (row * stride + col * chunk) / 16;
    for (int v = lane; v < vec_per_run; v += 32) {
      const uint4 x = __ldcs(src + base + v);
      __stcs(dst + base + v, x);
    }
  }
}

int main() {
  const long bytes = 2L << 30;   // 2 GiB per buffer: far beyond the 126 MB L2
  uint4 *src, *dst;
  if (cudaMalloc(&src, bytes) != cudaSuccess || cudaMalloc(&dst, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(src, 1, bytes); cudaMemset(dst, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%8s %10s %10s\n", "chunk", "stride", "GB/s (rd+wr)");
  for (int chunk : {128, 256, 512, 1024, 2048, 4096, 16384}) {
    for (long stride : {(long)chunk, 8192L, 4L << 20}) {
      if (stride < chunk) continue;
      const long rows = bytes / stride;             // runs that are `stride` apart
      const long cols = stride / chunk;             // runs inside one stride
      const long runs = rows * cols;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        copy_runs<<<148 * 8, 256>>>(src, dst, runs, chunk, stride, rows);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%8d %10ld %10.0f\n", chunk, stride, 2.0 * bytes / ms / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}

// How long does a hand-off between two warps of a CTA take on a B200?  (round 2, dw_tc_wgrad3.cu: its per-phase pipeline
// is a chain of such hops.)  One CTA per SM, thread A = lane 0 of warp 0, thread B = lane 0 of warp 1; they play ping-pong:
//   mode 0: A mbarrier.arrive -> B wait, B mbarrier.arrive -> A wait                      (two plain hops per round)
//   mode 1: A tcgen05.commit (nothing outstanding) -> B wait, B arrive -> A wait            (one commit hop, one plain hop)
//   mode 2: A one M128 N32 K16 MMA + tcgen05.commit -> B wait, B arrive -> A wait
//   mode 3: as 2, and B (whole warp 1) reads 32 TMEM columns (tcgen05.ld + wait::ld) and fences before it arrives
//   mode 4: as 3, and B also writes 16 columns (tcgen05.st + wait::st)
//   tools/hop_probe <mode> [extra warps] [1 = only lane 0 of an extra warp polls]
// extra warps spin on a third mbarrier for the whole run (all 32 lanes, as the kernels' warp roles do): how much does their
// polling slow the hand-offs down?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../knowledge-distillation-by-replacing-cheap-conv_b200/csrc/sm100_ptx.cuh"

using namespace kdcc;

__global__ void __launch_bounds__(1024, 1) hop(int mode, int iters, int lane0_only, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[3];
  __shared__ uint32_t slot;
  const uint32_t barA = ptx::smem_u32(&bars[0]), barB = ptx::smem_u32(&bars[1]), barC = ptx::smem_u32(&bars[2]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem_raw + (base - ptx::smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(barA, 1); ptx::mbar_init(barB, 1); ptx::mbar_init(barC, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<128>(ptx::smem_u32(&slot));
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t hi = (128u >> 4) | (1u << 14);
  const uint32_t b_lo = ((base & 0x3FFFF) >> 4) | ((512u >> 4) << 16);
  if (warp == 0) {
    if (ptx::elect_one()) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (mode == 0) ptx::mbar_arrive(barA);
        else {
          if (mode >= 2) ptx::umma_f16_ts(tmem, tmem + 64, b_lo, hi, idesc, 0u);
          ptx::umma_commit(barA);
        }
        ptx::mbar_wait(barB, (uint32_t)(i & 1));
        ptx::tcgen05_fence_after();
      }
      out[blockIdx.x] = clock64() - t0;
      ptx::mbar_arrive(barC);
    }
  } else if (warp >= 2) {
    if (lane0_only) {
      if (lane == 0) ptx::mbar_wait(barC, 0);
      __syncwarp();
    } else {
      ptx::mbar_wait(barC, 0);
    }
  } else {
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
      ptx::mbar_wait(barA, (uint32_t)(i & 1));
      if (mode >= 3) {
        ptx::tcgen05_fence_after();
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)32 << 16), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) acc ^= r[q];
        if (mode >= 4) {
          uint32_t o[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) o[q] = r[2 * q] + r[2 * q + 1];
          ptx::tmem_st_32x32b_x16(tmem + 96 + ((uint32_t)32 << 16), o);
          ptx::tmem_st_wait();
        }
        ptx::tcgen05_fence_before();
        __syncwarp();
      }
      if (lane == 0) ptx::mbar_arrive(barB);
    }
    if (acc == 0x12345u) out[200] = acc;
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<128>(tmem);
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, iters = 2000, extra = argc > 2 ? atoi(argv[2]) : 0, lane0 = argc > 3 ? atoi(argv[3]) : 0;
  long long *d_out;
  cudaMalloc(&d_out, 4096 * sizeof(long long));
  cudaFuncSetAttribute(hop, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    hop<<<148, 64 + 32 * extra, 32 * 1024>>>(mode, iters, lane0, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
  }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int b = 0; b < 148; ++b) mx = h[b] > mx ? h[b] : mx;
  printf("mode %d, %d polling warps%s: %.1f clk per round trip\n", mode, extra, lane0 ? " (lane 0 only)" : "", (double)mx / iters);
  return 0;
}

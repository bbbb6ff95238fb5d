// Micro-benchmark (measurement tooling, not product): cycles per tcgen05.mma for the small-N shapes the
// depthwise kernels issue, by operand source/layout.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
#include "../knowledge-distillation-by-replacing-cheap-conv_b200/csrc/sm100_ptx.cuh"

using namespace kdcc;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}


__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
      "r"(idesc), "r"(1u)
      : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t tmem_dst, uint32_t lo, uint32_t hi) {
  asm volatile("{\n.reg .b64 d;\nmov.b64 d, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], d;\n}\n" ::"r"(tmem_dst), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc),
      "r"(1u)
      : "memory");
}

// mode 0: SS, A K-major SW128, B K-major no-swizzle     (dw_tc.cu today)
// mode 1: SS, A K-major no-swizzle chunk-major, B K-major no-swizzle   (phase layout)
// mode 2: TS, A TMEM, B K-major no-swizzle
// mode 3: TS, A TMEM, B MN-major SW128                  (dw_tc_wgrad.cu)
// mode 4: SS, A K-major SW128, B K-major SW128           (plain GEMM)
// mode 5: tcgen05.cp 128x256b (A slice smem -> TMEM) + TS MMA;  mode 6: the copies alone
__global__ void __launch_bounds__(128, 1) probe(int mode, int N, int iters, int nacc, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(gen)[i] = 0x3c003c00u + (i & 3);
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(ptx::smem_u32(&slot));
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32 && elect_one()) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (mode == 3) idesc |= 1u << 16;
    const uint32_t a_base = base, b_base = base + 96 * 1024;
    const uint32_t sw_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t ns_hi = (128u >> 4) | (1u << 14);
    uint32_t a_lo, a_hi, b_lo, b_hi;
    a_lo = ((a_base & 0x3FFFF) >> 4) | (1u << 16); a_hi = sw_hi;
    if (mode == 1 || mode >= 5) { a_lo = ((a_base & 0x3FFFF) >> 4) | ((uint32_t)((168 * 16) >> 4) << 16); a_hi = ns_hi; }
    b_lo = ((b_base & 0x3FFFF) >> 4) | ((uint32_t)((N * 16) >> 4) << 16); b_hi = ns_hi;
    if (mode == 3) { b_lo = ((b_base & 0x3FFFF) >> 4) | ((uint32_t)((16384) >> 4) << 16); b_hi = sw_hi; }
    if (mode == 4) { b_lo = ((b_base & 0x3FFFF) >> 4) | (1u << 16); b_hi = sw_hi; }
    const long long t0 = clock64();
    // straight-line groups of 8 MMAs over NACC accumulators; all operand variation is compile-time
    auto body = [&](auto nacc_c) {
      constexpr int NACC = decltype(nacc_c)::value;
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t d = tmem + (uint32_t)(j % NACC) * (uint32_t)N;
          if (mode == 5 || mode == 6) {
            // copy the next A slice into one of 8 TMEM slots, then multiply from it (mode 6: copies only)
            cp_128x256b(tmem + 448u + (uint32_t)j * 8u, a_lo + (uint32_t)j * 5u, a_hi);
            if (mode == 5) mma_ts(d, tmem + 448u + (uint32_t)j * 8u, b_lo + (uint32_t)(j & 1) * 2u, b_hi, idesc);
          } else if (mode == 2 || mode == 3) mma_ts(d, tmem + 480u + (uint32_t)(j & 1) * 8u, b_lo + (uint32_t)j * 8u, b_hi, idesc);
          else mma_ss(d, a_lo + (mode == 1 ? (uint32_t)j * 5u : (uint32_t)j * 8u), a_hi, b_lo + (uint32_t)(j & 1) * 2u, b_hi, idesc);
        }
      }
    };
    if (nacc == 1) body(std::integral_constant<int, 1>{});
    else if (nacc == 2) body(std::integral_constant<int, 2>{});
    else if (nacc == 4) body(std::integral_constant<int, 4>{});
    else body(std::integral_constant<int, 8>{});
    const long long t1 = clock64();
    ptx::umma_commit(ptx::smem_u32(&bar));
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 2 * 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  const int Ns[] = {32, 64, 128, 176, 256};
  const int accs[] = {1, 2, 4, 8};
  for (int mode = 0; mode < 7; ++mode)
    for (int N : Ns) for (int nacc : accs) {
      if (N == 176 && mode != 3 && mode != 2) continue;
      if (nacc * N > 448) continue;
      if (mode >= 5 && N != 32) continue;
      if (mode < 5 && !(N == 32 && nacc == 4)) continue;
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<148, 128, 200 * 1024>>>(mode, N, iters, nacc, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      }
      long long h[2 * 148];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0, mi = 0;
      for (int b = 0; b < 148; ++b) { if (h[2 * b + 1] > mx) mx = h[2 * b + 1]; if (h[2 * b] > mi) mi = h[2 * b]; }
      printf("mode %d N %3d nacc %2d: issue %.1f clk/mma, complete %.1f clk/mma (block0 %.1f)\n", mode, N, nacc, (double)mi / iters,
             (double)mx / iters, (double)h[1] / iters);
    }
  return 0;
}

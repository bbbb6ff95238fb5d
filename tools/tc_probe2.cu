// Micro-benchmark (measurement tooling, not product): what bounds the depthwise tensor-core kernels on a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe2 tools/tc_probe2.cu
//   tools/tc_probe2 <mode> [M] [N] [warps]
// modes (one CTA per SM, clocks from clock64, max over the 148 CTAs):
//   ss      tcgen05.mma SS, A and B un-swizzled K-major (the conv operand layouts), A address changes per MMA
//   ssuse   the same, groups of 3 MMAs on one A tile: collector::a::fill, ::use, ::lastuse
//   ts      tcgen05.mma TS (A in TMEM)
//   wsss    tcgen05.mma.ws SS (M = 32 / 64 / 128)
//   wsts    tcgen05.mma.ws TS
//   ldtm    tcgen05.ld 32x32b.x32 throughput with <warps> warps
//   sttm    tcgen05.st 32x32b.x32 throughput with <warps> warps
//   shfl    shfl.sync.up throughput with <warps> warps
// Each mode runs in its own process (tools/gpu_probe2.sh): an illegal shape poisons the context.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../knowledge-distillation-by-replacing-cheap-conv_b200/csrc/sm100_ptx.cuh"

using namespace kdcc;

enum { SS = 0, SSUSE, TS, WSSS, WSTS, LDTM, STTM, SHFL };

#define MMA_SS(QUAL)                                                                                                       \
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n" \
               "tcgen05.mma" QUAL " [%0], da, db, %5, p;\n}\n" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),         \
               "r"(idesc), "r"(1u)                                                                                         \
               : "memory")
#define MMA_TS(QUAL)                                                                                                  \
  asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"                        \
               "tcgen05.mma" QUAL " [%0], [%1], db, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), \
               "r"(1u)                                                                                                \
               : "memory")

__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_SS(".cta_group::1.kind::f16");
}
__device__ __forceinline__ void mma_ss_fill(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_SS(".cta_group::1.kind::f16.collector::a::fill");
}
__device__ __forceinline__ void mma_ss_use(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_SS(".cta_group::1.kind::f16.collector::a::use");
}
__device__ __forceinline__ void mma_ss_last(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_SS(".cta_group::1.kind::f16.collector::a::lastuse");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_TS(".cta_group::1.kind::f16");
}
__device__ __forceinline__ void mma_ws_ss(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_SS(".ws.cta_group::1.kind::f16");
}
__device__ __forceinline__ void mma_ws_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  MMA_TS(".ws.cta_group::1.kind::f16");
}

// g_nacc: accumulators the MMA loop rotates over (0 = as many as fit, up to 4; 1 = every MMA accumulates into the same tile)
// g_bmode: 0 = B K-major; 1 = B MN-major without swizzle, 8-column chunks 168 rows apart, start row 20 * (j % 3) (dw_tc_wgrad3.cu)
__constant__ int g_nacc, g_bmode, g_ldwarps;  // g_ldwarps: warps 4.. read TMEM (tcgen05.ld x32 + wait, or st x32 if negative) while the MMAs run

template <int mode>
__global__ void __launch_bounds__(512, 1) probe(int M, int N, int iters, long long *out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ long long wclk[16];
  __shared__ volatile int stop_flag;
  __shared__ unsigned int ld_ops[32];
  if (threadIdx.x == 0) stop_flag = 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(gen)[i] = 0x3c003c00u + (i & 3);
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(ptx::smem_u32(&slot));
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (mode <= WSTS) {
    if (threadIdx.x < 32 && ptx::elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | (g_bmode ? (1u << 16) : 0u);
      const uint32_t a_base = base, b_base = base + 96 * 1024;
      const uint32_t hi = (128u >> 4) | (1u << 14);  // 8-row groups 128 B apart, version 1, no swizzle
      const uint32_t bhi = g_bmode ? ((168u * 16u) >> 4) | (1u << 14) : hi;
      const uint32_t a_lo = ((a_base & 0x3FFFF) >> 4) | ((uint32_t)((168 * 16) >> 4) << 16);  // K chunks 168 rows apart
      const uint32_t b_lo = ((b_base & 0x3FFFF) >> 4) | ((uint32_t)(((g_bmode ? 128 : N * 16)) >> 4) << 16);
      const int nacc = g_nacc ? g_nacc : (448 / N < 4 ? (448 / N < 2 ? 1 : 2) : 4);
      const uint32_t bstep = g_bmode ? 20u : 0u;
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 12) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const uint32_t d = tmem + (uint32_t)(j % nacc) * (uint32_t)N;
          const uint32_t a_j = a_lo + (uint32_t)j * 5u, b_j = b_lo + (uint32_t)(j & 1) * 2u + (uint32_t)(j % 3) * bstep;
          const uint32_t a_t = tmem + 480u + (uint32_t)(j & 1) * 8u;
          if constexpr (mode == SS) mma_ss(d, a_j, hi, b_j, hi, idesc);
          else if constexpr (mode == SSUSE) {
            const uint32_t a_g = a_lo + (uint32_t)(j / 3) * 5u;
            if (j % 3 == 0) mma_ss_fill(d, a_g, hi, b_j, hi, idesc);
            else if (j % 3 == 1) mma_ss_use(d, a_g, hi, b_j, hi, idesc);
            else mma_ss_last(d, a_g, hi, b_j, hi, idesc);
          } else if constexpr (mode == TS) mma_ts(d, a_t, b_j, bhi, idesc);
          else if constexpr (mode == WSSS) mma_ws_ss(d, a_j, hi, b_j, hi, idesc);
          else mma_ws_ts(d, a_t, b_j, hi, idesc);
        }
      }
      const long long t1 = clock64();
      ptx::umma_commit(ptx::smem_u32(&bar));
      ptx::mbar_wait(ptx::smem_u32(&bar), 0);
      const long long t2 = clock64();
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t2 - t0;
      stop_flag = 1;
    }
    if (warp >= 4) {  // background TMEM traffic from other warps (their quadrant, columns 256..383: not the MMA's tiles)
      const uint32_t t_row = tmem + 256u + ((uint32_t)((warp & 3) * 32) << 16);
      unsigned int ops = 0;
      uint32_t acc = 0, r[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) r[q] = q;
      while (!stop_flag) {
        if (g_ldwarps > 0) {
          ptx::tmem_ld_32x32b_x32(t_row + (ops & 3) * 32u, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) acc ^= r[q];
        } else {
          ptx::tmem_st_32x32b_x32(t_row + (ops & 3) * 32u, r);
          ptx::tmem_st_wait();
        }
        ++ops;
      }
      if (lane == 0) ld_ops[warp] = ops;
      if (acc == 0x12345678u) out[500] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tot = 0;
      for (int w = 4; w < (int)(blockDim.x >> 5); ++w) tot += ld_ops[w];
      out[400 + blockIdx.x] = (long long)tot;
    }
  } else {
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if constexpr (mode == LDTM) {
      const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      for (int i = 0; i < iters; i += 4) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)((i * 32) & 255), r0);
        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)((i * 32 + 32) & 255), r1);
        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)((i * 32 + 64) & 255), r2);
        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)((i * 32 + 96) & 255), r3);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) acc ^= r0[q] ^ r1[q] ^ r2[q] ^ r3[q];
      }
    } else if constexpr (mode == STTM) {
      const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) r[q] = threadIdx.x * 33 + q;
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) ptx::tmem_st_32x32b_x32(t_row + (uint32_t)(((i + j) * 32) & 255), r);
        ptx::tmem_st_wait();
        r[i & 31] += 1;
      }
      acc = r[5];
    } else {
      uint32_t v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = threadIdx.x * 7 + q;
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __shfl_up_sync(0xffffffffu, v[q], 5) + 1;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc ^= v[q];
    }
    const long long t1 = clock64();
    if (lane == 0) wclk[warp] = t1 - t0;
    if (acc == 0x12345678u) out[400 + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long mx = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = wclk[w] > mx ? wclk[w] : mx;
      out[2 * blockIdx.x] = mx;
      out[2 * blockIdx.x + 1] = mx;
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

template <int MODE>
static void launch1(int threads, int M, int N, int iters, long long *d_out) {
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<MODE><<<148, threads, 200 * 1024>>>(M, N, iters, d_out);
}
static void launch(int mode, int threads, int M, int N, int iters, long long *d_out) {
  switch (mode) {
    case SS: launch1<SS>(threads, M, N, iters, d_out); break;
    case SSUSE: launch1<SSUSE>(threads, M, N, iters, d_out); break;
    case TS: launch1<TS>(threads, M, N, iters, d_out); break;
    case WSSS: launch1<WSSS>(threads, M, N, iters, d_out); break;
    case WSTS: launch1<WSTS>(threads, M, N, iters, d_out); break;
    case LDTM: launch1<LDTM>(threads, M, N, iters, d_out); break;
    case STTM: launch1<STTM>(threads, M, N, iters, d_out); break;
    default: launch1<SHFL>(threads, M, N, iters, d_out); break;
  }
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const char *names[] = {"ss", "ssuse", "ts", "wsss", "wsts", "ldtm", "sttm", "shfl"};
  int mode = -1;
  for (int i = 0; i < 8; ++i) if (!strcmp(argv[1], names[i])) mode = i;
  if (mode < 0) return 2;
  const int M = argc > 2 ? atoi(argv[2]) : 128, N = argc > 3 ? atoi(argv[3]) : 32, warps = argc > 4 ? atoi(argv[4]) : 4;
  const int nacc = mode <= WSTS && argc > 4 ? atoi(argv[4]) : 0, bmode = argc > 5 ? atoi(argv[5]) : 0, ldwarps = argc > 6 ? atoi(argv[6]) : 0;
  cudaMemcpyToSymbol(g_ldwarps, &ldwarps, sizeof(int));
  cudaMemcpyToSymbol(g_nacc, &nacc, sizeof(int));
  cudaMemcpyToSymbol(g_bmode, &bmode, sizeof(int));
  long long *d_out;
  cudaMalloc(&d_out, 4096 * sizeof(long long));
  const int iters = mode <= WSTS ? 4092 : 4096;
  const int threads = mode <= WSTS ? 128 + 32 * abs(ldwarps) : 32 * warps;
  for (int rep = 0; rep < 2; ++rep) {
    launch(mode, threads, M, N, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s M %d N %d: %s\n", argv[1], M, N, cudaGetErrorString(e)); return 1; }
  }
  long long h[2 * 148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mi = 0;
  for (int b = 0; b < 148; ++b) { if (h[2 * b + 1] > mx) mx = h[2 * b + 1]; if (h[2 * b] > mi) mi = h[2 * b]; }
  if (mode <= WSTS && ldwarps) {
    long long ops[148];
    cudaMemcpy(ops, d_out + 400, sizeof(ops), cudaMemcpyDeviceToHost);
    printf("  with %d warps doing tcgen05.%s x32 + wait: %.1f B/clk/SM of TMEM traffic beside the MMAs\n", abs(ldwarps), ldwarps > 0 ? "ld" : "st",
           4096.0 * ops[0] / (double)h[1]);
  }
  if (mode <= WSTS)
    printf("%-5s M %3d N %3d nacc %d bmode %d: issue %.1f clk/mma, complete %.1f clk/mma\n", argv[1], M, N, nacc, bmode, (double)mi / iters, (double)mx / iters);
  else if (mode == SHFL)
    printf("%-5s warps %2d: %.2f clk per warp-shfl per SM (%.2f per warp)\n", argv[1], warps, (double)mx / iters / warps, (double)mx / iters);
  else
    printf("%-5s warps %2d: %.1f clk per x32 op per warp, %.1f B/clk/SM\n", argv[1], warps, (double)mx / iters, 4096.0 * warps * iters / mx);
  return 0;
}

// Micro-benchmark (measurement tooling, not product): what the fp32 pipe of one SM sustains for the instruction
// mixes of the streaming NHWC 3x3 kernel (dw_nhwc3.cu), so that its arithmetic floor is a measured number:
//   mode 0  FFMA   d = a * b + d        three distinct register sources (weights and inputs in registers)
//   mode 1  FFMA   d = a * b + d        `a` shared by consecutive instructions (operand reuse cache)
//   mode 2  FFMA2  packed fma.rn.f32x2, three distinct 64-bit sources
//   mode 3  FFMA2  `b` shared by three consecutive instructions (the kernel's pattern: one input pair, three tap rows)
//   mode 4  the bf16 -> fp32 widening (shift / mask) alone
// Prints FMA lanes per clock per SM for 1..12 resident warps per scheduler-quarter layout (blockDim = 32 * warps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fma_probe tools/fma_probe.cu && tools/fma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;"
               : "=l"(*reinterpret_cast<unsigned long long *>(&d))
               : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
                 "l"(*reinterpret_cast<unsigned long long *>(&c)));
  return d;
}

constexpr int ACC = 12;   // independent accumulator chains per thread (the kernel has 3 rows x 4 pairs)

template <int MODE>
__global__ void probe(int iters, float seed, float *sink, long long *clk) {
  float a[ACC], b[ACC], d[ACC];
  float2 a2[ACC], b2[ACC], d2[ACC];
  uint32_t w[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) {
    a[i] = seed + i; b[i] = seed * 0.5f + i; d[i] = 0.f;
    a2[i] = make_float2(a[i], a[i] + 1); b2[i] = make_float2(b[i], b[i] + 1); d2[i] = make_float2(0.f, 0.f);
    w[i] = __float_as_uint(seed) + i * 0x10001u;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ACC; ++i) {
        if (MODE == 0) d[i] = fmaf(a[i], b[(i + r) % ACC], d[i]);
        if (MODE == 1) d[i] = fmaf(a[r], b[i], d[i]);
        if (MODE == 2) d2[i] = ffma2(a2[i], b2[(i + r) % ACC], d2[i]);
        if (MODE == 3) d2[i] = ffma2(a2[i], b2[(i / 3 + r) % ACC], d2[i]);
        if (MODE == 4) { d[i] += __uint_as_float(w[i] << 16); a[i] += __uint_as_float(w[(i + r) % ACC] & 0xffff0000u); }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i] + d2[i].x + d2[i].y + a[i];
  if (s == 12345.678f) sink[0] = s;   // keep the chains alive
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int lanes_per_instr) {
  float *sink; long long *clk;
  cudaMalloc(&sink, 4); cudaMalloc(&clk, 8 * 1024);
  const int iters = 2000;
  printf("%-44s", name);
  for (int warps : {1, 2, 4, 8, 12, 16}) {
    probe<MODE><<<148, 32 * warps>>>(iters, 1.0f, sink, clk);   // warm-up
    probe<MODE><<<148, 32 * warps>>>(iters, 1.0f, sink, clk);
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    long long worst = 0;
    for (long long v : h) worst = v > worst ? v : worst;
    const double instr = (double)iters * 4 * ACC * warps;              // warp instructions per SM
    printf("  w%-2d %6.1f", warps, instr * lanes_per_instr / (double)worst);   // fp32 FMA lanes per clock per SM
  }
  printf("   (FMA lanes / clk / SM; 128 = one FFMA per lane-slot per clock)\n");
  cudaFree(sink); cudaFree(clk);
}

int main() {
  run<0>("FFMA, 3 distinct sources", 32);
  run<1>("FFMA, shared multiplicand (reuse)", 32);
  run<2>("FFMA2, 3 distinct 64-bit sources", 64);
  run<3>("FFMA2, input pair shared by 3 instructions", 64);
  run<4>("bf16 widening pair (shift + mask + 2 FADD)", 64);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}

// Microbenchmark (dev tool): single-warp issue rates on sm_100 — independent FFMA, MUFU.EX2, and the softmax mix.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float x[16], y[16];
  for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 1e-3f + i * 0.01f - 1.0f; y[i] = x[i] * 0.5f; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(y[i])); }
      if (MODE == 1) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      if (MODE == 2) {  // 1 MUFU + 3 FMA-pipe ops, all independent across i
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(y[i]) : "f"(x[(i + 8) & 15]));
        asm volatile("add.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(x[(i + 4) & 15]));
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(y[i]) : "f"(x[(i + 2) & 15]));
      }
      if (MODE == 3) {  // 1 MUFU + 3 integer ALU ops
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        unsigned u = __float_as_uint(y[i]);
        asm volatile("add.u32 %0, %0, 32768;" : "+r"(u));
        asm volatile("shr.u32 %0, %0, 3;" : "+r"(u));
        asm volatile("xor.b32 %0, %0, 0x55;" : "+r"(u));
        y[i] = __uint_as_float(u);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const char* names[] = {"FFMA x16", "MUFU.EX2 x16", "16 x (MUFU + 3 FP)", "16 x (MUFU + 3 INT)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int threads : {128, 256}) {
      if (mode == 0) k<0><<<148, threads>>>(out, 1000, clk);
      if (mode == 1) k<1><<<148, threads>>>(out, 1000, clk);
      if (mode == 2) k<2><<<148, threads>>>(out, 1000, clk);
      if (mode == 3) k<3><<<148, threads>>>(out, 1000, clk);
      cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      printf("%-22s warps/SMSP %d: %.2f clk per group-of-16 element (per warp)\n", names[mode], threads / 128, (double)h / (1000.0 * 16));
    }
  return 0;
}

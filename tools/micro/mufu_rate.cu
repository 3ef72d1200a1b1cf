// Microbenchmark: issue interval of MUFU.EX2 per warp / per SMSP on this GPU (dev tool).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* clk) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i * 0.01f - 1.0f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  for (int threads : {32, 128, 256, 512}) {
    k<<<148, threads>>>(out, 1000, clk); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    int warps_per_smsp = (threads / 32 + 3) / 4;
    printf("threads/SM %4d: %.2f clk per MUFU.EX2 warp-instr per warp; per SMSP %.2f clk/instr\n", threads,
           (double)h / (1000.0 * 16), (double)h / (1000.0 * 16 * warps_per_smsp));
  }
  return 0;
}

// Microbenchmark (dev tool): what does a co-resident warp cost the attention kernel's exp2 pass on the same SMSP?
// Warp A (warps 2..5, one per SMSP = TMEM lane quarter) runs the kernel's exp2 pass (tcgen05.ld -> FFMA2 -> MUFU.EX2 ->
// integer pack -> tcgen05.st) over 208 columns in a loop; warp B (warps 6..9, same SMSPs) runs one of:
//   0 nothing, 1 tcgen05.ld stream (2 x32 loads per wait), 2 row-maximum pass, 3 FMNMX only (no TMEM), 4 LDS spin (the turn
//   wait), 5 mbarrier try_wait on a barrier that never completes, 9 mbarrier test_wait spin, 6 the same exp2 pass,
//   7 FFMA chain, 8 O drain + 8 STG.128
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -o tools/micro/_bin/smsp_interference tools/micro/smsp_interference.cu -lcuda
#include <cstdio>
#include <cstdlib>

#include "../../arp_b200/csrc/attention_tc.cuh"
using namespace arp;

__device__ __forceinline__ void exp_pass(uint32_t t_s) {
  const float scale_log2e = 0.18f, mo = 3.0f;
  const uint64_t scale2 = f32x2_pack(scale_log2e, scale_log2e), nmo2 = f32x2_pack(-mo, -mo);
  auto soft_chunk = [&](const uint32_t(&src)[32], int c) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float x0, x1;
      f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(src[j]), __uint_as_float(src[j + 1])), scale2, nmo2), x0, x1);
      pk[j >> 1] = pack_bf16_trunc(ex2_approx(x0), ex2_approx(x1));
    }
    tmem_st_32x16(t_s + c * 16, pk);
  };
  uint32_t b0[32], b1[32];
  tmem_ld_32x32(t_s, b0);
  tmem_ld_wait();
#pragma unroll 1
  for (int c = 0; c + 1 < 6; c += 2) {
    tmem_ld_32x32(t_s + (c + 1) * 32, b1);
    soft_chunk(b0, c);
    tmem_ld_wait();
    if (c + 2 < 6) tmem_ld_32x32(t_s + (c + 2) * 32, b0);
    soft_chunk(b1, c + 1);
    tmem_ld_wait();
  }
  {
    uint32_t r[16], pk[8];
    tmem_ld_32x16(t_s + 192, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 2)
      pk[j >> 1] = pack_bf16_trunc(ex2_approx(fmaf(__uint_as_float(r[j]), scale_log2e, -mo)),
                                   ex2_approx(fmaf(__uint_as_float(r[j + 1]), scale_log2e, -mo)));
    tmem_st_32x8(t_s + 96, pk);
  }
  tmem_st_wait();
}

template <int MODE>
__device__ __forceinline__ void b_work(uint32_t t_s, float& acc, volatile int* flag, uint64_t* never, uint4* gout) {
  if (MODE == 1 || MODE == 2) {
    uint32_t a[32], b[32];
    float m = acc;
#pragma unroll 1
    for (int c = 0; c + 1 < 6; c += 2) {
      tmem_ld_32x32(t_s + c * 32, a);
      tmem_ld_32x32(t_s + c * 32 + 32, b);
      tmem_ld_wait();
      if (MODE == 2) {
        float m0 = m, m1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(a[j]), __uint_as_float(a[j + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(b[j]), __uint_as_float(b[j + 1])));
        }
        m = fmaxf(m0, m1);
      } else {
        m += __uint_as_float(a[0] ^ b[31]);
      }
    }
    acc = m;
  } else if (MODE == 3) {
    float m0 = acc, m1 = -acc;
#pragma unroll
    for (int j = 0; j < 104; ++j) {
      m0 = fmaxf(m0, fmaxf(m1 + 1.0f, acc));
      m1 = fmaxf(m1, fmaxf(m0, -acc));
    }
    acc = m0 + m1;
  } else if (MODE == 4) {
    for (int i = 0; i < 16; ++i)
      if (lds_volatile(flag) == 12345 + i) acc += 1.f;
  } else if (MODE == 5) {
    mbar_try_wait(never, 0);
  } else if (MODE == 9) {
    for (int i = 0; i < 16; ++i)
      if (mbar_test_wait(never, 0)) acc += 1.f;
  } else if (MODE == 6) {
    exp_pass(t_s);
  } else if (MODE == 7) {
    float x = acc, y = acc + 1.f;
#pragma unroll
    for (int j = 0; j < 104; ++j) { x = fmaf(x, 1.0001f, y); y = fmaf(y, 0.9999f, x); }
    acc = x + y;
  } else if (MODE == 8) {
    uint32_t o0[32], o1[32], os[16];
    tmem_ld_32x32(t_s + 128, o0);
    tmem_ld_32x32(t_s + 160, o1);
    tmem_ld_32x16(t_s + 192, os);
    tmem_ld_wait();
    const float inv = 1.0f / (__uint_as_float(os[0]) + 2.f);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      gout[c] = make_uint4(pack_op(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv),
                           pack_op(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv),
                           pack_op(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv),
                           pack_op(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv));
#pragma unroll
    for (int c = 0; c < 4; ++c)
      gout[4 + c] = make_uint4(pack_op(__uint_as_float(o1[8 * c]) * inv, __uint_as_float(o1[8 * c + 1]) * inv),
                               pack_op(__uint_as_float(o1[8 * c + 2]) * inv, __uint_as_float(o1[8 * c + 3]) * inv),
                               pack_op(__uint_as_float(o1[8 * c + 4]) * inv, __uint_as_float(o1[8 * c + 5]) * inv),
                               pack_op(__uint_as_float(o1[8 * c + 6]) * inv, __uint_as_float(o1[8 * c + 7]) * inv));
  }
}

template <int MODE>
__global__ void __launch_bounds__(320, 1) k(int iters, long long* out, float* sink, uint4* gbuf) {
  __shared__ uint64_t never;
  __shared__ uint32_t tmem_slot;
  __shared__ int done;
  __shared__ volatile int flag;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&never, 1); fence_mbar_init(); done = 0; flag = 0; }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int quarter = warp & 3;
  if (warp >= 2 && warp < 6) {
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    long long t0 = clock64();
    for (int n = 0; n < iters; ++n) exp_pass(t_s);
    long long t1 = clock64();
    if (lane == 0) {
      if (blockIdx.x == 0) out[quarter] = (t1 - t0) / iters;
      atomicAdd(&done, 1);
    }
  } else if (warp >= 6 && MODE != 0) {
    const uint32_t t_s = tmem_base + 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    float acc = lane * 0.01f;
    long long t0 = clock64();
    int n = 0;
    while (lds_volatile(&done) < 4) {
      b_work<MODE>(t_s, acc, &flag, &never, gbuf + (blockIdx.x * 128 + (warp - 6) * 32 + lane) * 8);
      ++n;
    }
    long long t1 = clock64();
    sink[blockIdx.x * 128 + (warp - 6) * 32 + lane] = acc;
    if (lane == 0 && blockIdx.x == 0) out[4 + quarter] = n ? (t1 - t0) / n : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

template <int MODE>
void run(const char* name, long long* out, float* sink, uint4* gbuf) {
  cudaMemset(out, 0, 64);
  k<MODE><<<148, 320>>>(2000, out, sink, gbuf);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[8];
  cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
  printf("B = %-36s | exp2 pass of A, clk q0..q3 %5lld %5lld %5lld %5lld | one B iteration %5lld %5lld %5lld %5lld\n", name,
         h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
}

int main() {
  long long* out; float* sink; uint4* gbuf;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 148 * 128 * 4); cudaMalloc(&gbuf, 148 * 128 * 128);
  run<0>("nothing", out, sink, gbuf);
  run<1>("tcgen05.ld stream", out, sink, gbuf);
  run<2>("row-maximum pass", out, sink, gbuf);
  run<3>("FMNMX only", out, sink, gbuf);
  run<4>("LDS spin", out, sink, gbuf);
  run<5>("mbarrier.try_wait (never completes)", out, sink, gbuf);
  run<9>("mbarrier.test_wait spin", out, sink, gbuf);
  run<6>("the same exp2 pass", out, sink, gbuf);
  run<7>("FFMA chain", out, sink, gbuf);
  run<8>("O drain + 8 STG.128", out, sink, gbuf);
  return 0;
}

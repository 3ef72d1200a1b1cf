// Microbenchmark (dev tool): how long do the attention kernel's two MMA batches take while softmax-style warps of the
// OTHER query-tile slot stream tcgen05.ld / tcgen05.st / MUFU on the same SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tmem_contention tools/micro/tmem_contention.cu -lcuda
// MMA batches (warp 0, one elected thread, commit + mbarrier wait per batch):
//   1 = Q K^T of one query tile: 4 x (M128 N208 K16), both operands from smem, D = TMEM cols [0, 208)
//   2 = P [V|1]: 13 x (M128 N80 K16), A from TMEM cols [0, 104), D = cols [128, 208)
//   3 = P V without the ones block: 13 x (M128 N64 K16)
//   4 = 1 then 2 back to back (one commit)
// Softmax-style load (warps 2..5 = TMEM lane quarters 0..3, region cols [256, 464)):
//   0 idle, 1 LDTM only, 2 row-maximum pass, 3 full exp2 pass (LDTM + FFMA2 + MUFU + pack + STTM), 4 STTM only,
//   5 MUFU only, 6 exp2 pass without STTM, 7 exp2 pass without LDTM
#include <cstdio>
#include <cstdlib>

#include "../../arp_b200/csrc/attention_tc.cuh"

using namespace arp;

constexpr int NK = 208;

__device__ __forceinline__ void soft_pass(int mode, uint32_t t_s, float& acc) {
  const float scale = 0.18f, mo = 3.0f;
  const uint64_t scale2 = f32x2_pack(scale, scale), nmo2 = f32x2_pack(-mo, -mo);
  if (mode == 1 || mode == 2) {
    uint32_t a[32], b[32];
    float m = acc;
#pragma unroll 1
    for (int c = 0; c + 1 < 6; c += 2) {
      tmem_ld_32x32(t_s + c * 32, a);
      tmem_ld_32x32(t_s + c * 32 + 32, b);
      tmem_ld_wait();
      if (mode == 2) {
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
    uint32_t r[16];
    tmem_ld_32x16(t_s + 192, r);
    tmem_ld_wait();
    m = fmaxf(m, __uint_as_float(r[3]));
    acc = m;
    return;
  }
  if (mode == 5) {
    float x = acc;
#pragma unroll 1
    for (int c = 0; c < 13; ++c) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ex2_approx(x + j);
#pragma unroll
      for (int j = 0; j < 16; ++j) x += v[j];
    }
    acc = x;
    return;
  }
  if (mode == 4) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = __float_as_uint(acc) + j;
#pragma unroll 1
    for (int c = 0; c < 6; ++c) tmem_st_32x16(t_s + c * 16, pk);
    tmem_st_wait();
    return;
  }
  // modes 3, 6, 7: the exp2 pass
  uint32_t b0[32], b1[32];
  auto chunk = [&](const uint32_t(&src)[32], int c) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float x0, x1;
      f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(src[j]), __uint_as_float(src[j + 1])), scale2, nmo2), x0, x1);
      pk[j >> 1] = pack_bf16_trunc(ex2_approx(x0), ex2_approx(x1));
    }
    if (mode != 6) tmem_st_32x16(t_s + c * 16, pk);
    else acc += __uint_as_float(pk[0] ^ pk[15]);
  };
  if (mode != 7) {
    tmem_ld_32x32(t_s, b0);
    tmem_ld_wait();
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) { b0[j] = __float_as_uint(acc + j); b1[j] = __float_as_uint(acc - j); }
  }
#pragma unroll 1
  for (int c = 0; c + 1 < 6; c += 2) {
    if (mode != 7) tmem_ld_32x32(t_s + (c + 1) * 32, b1);
    chunk(b0, c);
    if (mode != 7) tmem_ld_wait();
    if (mode != 7 && c + 2 < 6) tmem_ld_32x32(t_s + (c + 2) * 32, b0);
    chunk(b1, c + 1);
    if (mode != 7) tmem_ld_wait();
  }
  if (mode != 6) tmem_st_wait();
}

__global__ void __launch_bounds__(320, 1) k(int mma_mode, int soft_mode, int iters, long long* out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 3 * 27648) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, NK);
  constexpr uint32_t idesc_o80 = umma_idesc_bf16(128, 80, 0, 1), idesc_o64 = umma_idesc_bf16(128, 64, 0, 1);
  const uint32_t sb = smem_u32(smem);
  if (warp == 0) {
    long long t0 = clock64();
    if (mma_mode) {
      for (int it = 0; it < iters; ++it) {
        if (elect_one()) {
          if (mma_mode == 1 || mma_mode == 4) {
            const uint64_t dq = umma_desc_kmajor_sw128(sb), dk = umma_desc_kmajor_sw128(sb + 16384);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
          }
          if (mma_mode >= 2) {
            const uint64_t dv = umma_desc_mnmajor_sw128(sb + 16384 + 27648, 27648);
#pragma unroll
            for (int kk = 0; kk < 13; ++kk)
              umma_bf16_ts(tmem_base + 128, tmem_base + kk * 8, dv + kk * (2048 >> 4), mma_mode == 3 ? idesc_o64 : idesc_o80,
                           kk != 0);
          }
          umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, it & 1);
        tc_fence_after();
      }
    }
    long long t1 = clock64();
    if (lane == 0) {
      if (blockIdx.x == 0) out[0] = mma_mode ? (t1 - t0) / iters : 0;
      done = 1;
    }
  } else if (warp >= 2 && warp < 6 && soft_mode) {
    const int quarter = warp & 3;
    const uint32_t t_s = tmem_base + 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    float acc = lane * 0.01f;
    long long t0 = clock64();
    int n = 0;
    if (mma_mode) {
      while (!done) { soft_pass(soft_mode, t_s, acc); ++n; }
    } else {
      for (; n < iters; ++n) soft_pass(soft_mode, t_s, acc);
    }
    long long t1 = clock64();
    sink[blockIdx.x * 128 + (warp - 2) * 32 + lane] = acc;
    if (lane == 0 && blockIdx.x == 0) out[1 + quarter] = n ? (t1 - t0) / n : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 148 * 128 * 4);
  const int smem = 16384 + 3 * 27648 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* mma_names[] = {"none", "QK^T 4x(128x208x16) SS", "PV 13x(128x80x16) TS", "PV 13x(128x64x16) TS", "QK^T + PV(80)"};
  const char* soft_names[] = {"idle", "LDTM only", "row-max pass", "exp2 pass (full)", "STTM only", "MUFU only",
                              "exp2 pass, no STTM", "exp2 pass, no LDTM"};
  for (int mm = 0; mm <= 4; ++mm)
    for (int sm = 0; sm <= 7; ++sm) {
      if (mm == 0 && sm == 0) continue;
      cudaMemset(out, 0, 64);
      k<<<148, 320, smem>>>(mm, sm, 2000, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[5];
      cudaMemcpy(h, out, 40, cudaMemcpyDeviceToHost);
      printf("mma %-26s | soft %-20s | clk per MMA batch %5lld | clk per soft pass q0..q3 %5lld %5lld %5lld %5lld\n", mma_names[mm],
             soft_names[sm], h[0], h[1], h[2], h[3], h[4]);
    }
  return 0;
}

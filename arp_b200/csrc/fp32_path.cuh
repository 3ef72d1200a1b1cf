// fp32 verification path (ArpConfig.precision = ARP_PREC_F32): the same pipeline as the bf16 tensor-core
// path with every activation and weight in fp32 and every contraction on the FP32 FMA pipe.
//
// Why it exists: BASELINE's parity statement has two bars — 1e-3 for the bf16 pipeline and 1e-5 for fp32.
// The reference computes in fp32 on CPU (clip.load(...).float(), label_reward.py:126-141), so an fp32 run of
// OUR pipeline isolates "is the restated algorithm right" from "how much does bf16 cost": decode, patch order,
// token layout, LayerNorm, attention scaling, QuickGELU, heads and the scan are shared host logic / kernels.
// It is a checker-grade path (~50x slower than the tensor-core path), not a product path.
#pragma once

#include "common.cuh"

namespace arp {

enum F32Act : int { F32_ACT_NONE = 0, F32_ACT_QUICKGELU = 1, F32_ACT_RELU = 2 };

// C[M,N] = act(A[M,K] W[N,K]^T + bias[N] + rowtab[row % period, N]) (+ resid[M,N]); all fp32, row-major,
// out may alias resid. 128x128 tile per CTA, 16-deep k slices, 8x8 outputs per thread. N % 128 == 0, K % 16 == 0.
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS)
sgemm_nt_f32_kernel(const float* __restrict__ A, const float* __restrict__ Wt, float* out, int M, int N, int K,
                    int lda, int ldo, const float* __restrict__ bias, const float* resid, int ldr,
                    const float* __restrict__ rowtab, int period, int act) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int ty = tid >> 4, tx = tid & 15;   // 16 x 16 threads, each 8 rows x 8 cols (strided by 16 for coalescing)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader mapping: 128 rows x 16 k = 512 float4; 256 threads x 2
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) a = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + r) * lda + k0 + lk);
      As[lk][r] = a.x; As[lk + 1][r] = a.y; As[lk + 2][r] = a.z; As[lk + 3][r] = a.w;
      const float4 b = *reinterpret_cast<const float4*>(Wt + static_cast<size_t>(n0 + r) * K + k0 + lk);
      Bs[lk][r] = b.x; Bs[lk + 1][r] = b.y; Bs[lk + 2][r] = b.z; Bs[lk + 3][r] = b.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty + 16 * i;
    if (row >= M) continue;
    const float* tab = rowtab ? rowtab + static_cast<size_t>(row % period) * N : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + tx + 16 * j;
      float v = acc[i][j];
      if (bias) v += bias[col];
      if (tab) v += tab[col];
      if (act == F32_ACT_QUICKGELU) v = v * (1.0f / (1.0f + expf(-1.702f * v)));   // x * sigmoid(1.702 x)
      else if (act == F32_ACT_RELU) v = fmaxf(v, 0.f);
      if (resid) v = resid[static_cast<size_t>(row) * ldr + col] + v;
      out[static_cast<size_t>(row) * ldo + col] = v;
    }
  }
}

// fp32 [T,3,224,224] (what the decode kernel emits for the reference's transform) -> fp32 patch rows
// A[b*tokens + 1 + py*G + px, c*P*P + ph*P + pw]; row b*tokens (class-token slot) = 0.
__global__ void __launch_bounds__(256)
im2col_f32_kernel(const float* __restrict__ chw, float* __restrict__ A, int T, int P, int G, int tokens) {
  const int K = 3 * P * P;
  const size_t total = static_cast<size_t>(T) * tokens * K;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % K);
    const size_t row = i / K;
    const int tok = static_cast<int>(row % tokens);
    const size_t b = row / tokens;
    float v = 0.f;
    if (tok > 0) {
      const int py = (tok - 1) / G, px = (tok - 1) % G;
      const int c = col / (P * P), ph = (col / P) % P, pw = col % P;
      v = chw[((b * 3 + c) * 224 + py * P + ph) * 224 + px * P + pw];
    }
    A[i] = v;
  }
}

// y fp32 [M,W] = LN(x fp32 [M,W]); y may alias x. One warp per row, two-pass moments.
template <int W>
__global__ void __launch_bounds__(256)
layernorm_f32_f32_kernel(const float* x, const float* __restrict__ gamma, const float* __restrict__ beta, float* y,
                         int M, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + static_cast<size_t>(row) * W;
  float v[W / 32];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < W / 32; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
  const float mean = warp_sum(s) * (1.0f / W);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < W / 32; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / W) + eps);
  float* yr = y + static_cast<size_t>(row) * W;
#pragma unroll
  for (int i = 0; i < W / 32; ++i) {
    const int c = lane + 32 * i;
    yr[c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

// softmax(Q K^T / 8) V per (head, frame); qkv fp32 [B*L, 3*width], out fp32 [B*L, width]. One CTA per
// (head, frame): K and V of the head in shared memory, one warp per query row, lanes over keys.
constexpr int A32_THREADS = 256;
__host__ __device__ inline int attn_f32_smem_bytes(int L) { return (2 * L * 65 + (A32_THREADS / 32) * (L + 64)) * 4; }

__global__ void __launch_bounds__(A32_THREADS)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int L, int width) {
  extern __shared__ float a32_smem[];
  float* sK = a32_smem;               // [L][65]
  float* sV = sK + L * 65;            // [L][65]
  float* sW = sV + L * 65;            // per warp: p[L] + q[64]
  const int head = blockIdx.x, frame = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + static_cast<size_t>(frame) * L * 3 * width + head * 64;
  for (int i = threadIdx.x; i < L * 64; i += A32_THREADS) {
    const int r = i >> 6, d = i & 63;
    sK[r * 65 + d] = base[static_cast<size_t>(r) * 3 * width + width + d];
    sV[r * 65 + d] = base[static_cast<size_t>(r) * 3 * width + 2 * width + d];
  }
  __syncthreads();
  float* p = sW + warp * (L + 64);
  float* q = p + L;
  for (int r = warp; r < L; r += A32_THREADS / 32) {
    q[lane] = base[static_cast<size_t>(r) * 3 * width + lane] * 0.125f;        // torch scales q before q k^T
    q[lane + 32] = base[static_cast<size_t>(r) * 3 * width + lane + 32] * 0.125f;
    __syncwarp();
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) {
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) s = fmaf(q[d], sK[j * 65 + d], s);
      p[j] = s;
      m = fmaxf(m, s);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = expf(p[j] - m);
      p[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, sV[j * 65 + lane], o0);
      o1 = fmaf(pj, sV[j * 65 + lane + 32], o1);
    }
    float* o = out + (static_cast<size_t>(frame) * L + r) * width + head * 64;
    o[lane] = o0 / sum;
    o[lane + 32] = o1 / sum;
    __syncwarp();
  }
}

// taps fp32 [B, ld_taps] columns [col0, col0+W) = x[b*tokens, :]   (fp32 twin of gather_cls_bf16_kernel)
template <int W>
__global__ void __launch_bounds__(256)
gather_cls_f32_kernel(const float* __restrict__ x, float* __restrict__ taps, int B, int tokens, int ld_taps, int col0) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const float* src = x + static_cast<size_t>(b) * tokens * W;
  float* dst = taps + static_cast<size_t>(b) * ld_taps + col0;
  for (int i = lane; i < W; i += 32) dst[i] = src[i];
}

}  // namespace arp

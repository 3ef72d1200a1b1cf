// LayerNorm and the small row-wise kernels around the ViT residual stream
// (CLIP computes LayerNorm in fp32, eps = 1e-5 — in-tree mirror
// arp_dt/models/openai/layers.py:9,246-248,322,331).
//
// All are HBM-bound: one warp per 768-wide row, the row lives in registers
// (24 floats / lane, float4 loads), two-pass mean/variance, warp-shuffle reductions.
#pragma once

#include "common.cuh"

namespace arp {

template <int W>
struct RowRegs {
  static_assert(W % 128 == 0, "row width must be a multiple of 128");
  static constexpr int kVec = W / 128;  // float4 per lane
  float4 v[kVec];
};

// CG = true: read the row with ld.global.cg (L2 only) — for rows another SM has just updated inside the same kernel.
template <int W, bool CG = false>
__device__ __forceinline__ void ln_row(const float* __restrict__ x, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, RowRegs<W>& r) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < RowRegs<W>::kVec; ++i) {
    r.v[i] = CG ? __ldcg(reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4))
                : *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
    s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / W);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < RowRegs<W>::kVec; ++i) {
    const float a = r.v[i].x - mean, b = r.v[i].y - mean, c = r.v[i].z - mean, d = r.v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / W) + eps);
#pragma unroll
  for (int i = 0; i < RowRegs<W>::kVec; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + (i * 32 + lane) * 4));
    r.v[i].x = (r.v[i].x - mean) * rstd * g.x + b.x;
    r.v[i].y = (r.v[i].y - mean) * rstd * g.y + b.y;
    r.v[i].z = (r.v[i].z - mean) * rstd * g.z + b.z;
    r.v[i].w = (r.v[i].w - mean) * rstd * g.w + b.w;
  }
}

// y(bf16)[M,W] = LN(x fp32 [M,W])                 (ln_1 / ln_2 feeding the QKV and c_fc GEMMs)
template <int W>
__global__ void __launch_bounds__(256)
layernorm_f32_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                          const float* __restrict__ beta, op_t* __restrict__ y, int M, float eps,
                          int reverse) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  if (reverse) row = M - 1 - row;   // snake order across kernels: start on the rows the previous kernel wrote last
  const int lane = threadIdx.x & 31;
  RowRegs<W> r;
  ln_row<W>(x + static_cast<size_t>(row) * W, gamma, beta, eps, r);
  uint2* out = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * W);
#pragma unroll
  for (int i = 0; i < RowRegs<W>::kVec; ++i)
    out[i * 32 + lane] = make_uint2(pack_op(r.v[i].x, r.v[i].y), pack_op(r.v[i].z, r.v[i].w));
}

// x fp32 [M,W] = LN(x) in place                    (ln_pre)
template <int W>
__global__ void __launch_bounds__(256)
layernorm_f32_inplace_kernel(float* __restrict__ x, const float* __restrict__ gamma,
                             const float* __restrict__ beta, int M, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  RowRegs<W> r;
  float* xr = x + static_cast<size_t>(row) * W;
  ln_row<W>(xr, gamma, beta, eps, r);
#pragma unroll
  for (int i = 0; i < RowRegs<W>::kVec; ++i) reinterpret_cast<float4*>(xr)[i * 32 + lane] = r.v[i];
}

// ---- 16-bit residual stream (the default path): LayerNorm is never materialised. The consumer GEMM multiplies the RAW
// stream by gamma-folded weights and applies (rstd, -mean*rstd) per row in its epilogue (gemm_tcgen05.cuh G2_LNFOLD);
// these kernels produce that float2 per row. Moments are taken of the ROUNDED 16-bit values — exactly what the tensor
// core multiplies — two-pass (mean, then centred squares) with the row in registers.
template <int W>
__device__ __forceinline__ float2 row_stats_from_regs(const float (&v)[W / 32], float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < W / 32; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / W);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < W / 32; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / W) + eps);
  return make_float2(rstd, -mean * rstd);
}

// stats[row] = (rstd, -mean*rstd) of x[row, :]   (x = 16-bit residual stream [M, W]); one warp per row, 16-byte loads
template <int W>
__global__ void __launch_bounds__(256)
row_moments_kernel(const op_t* __restrict__ x, float2* __restrict__ stats, int M, float eps, int reverse) {
  static_assert(W % 256 == 0, "row width must be a multiple of 256");
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  if (reverse) row = M - 1 - row;
  const int lane = threadIdx.x & 31;
  const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * W);
  float v[W / 32];
#pragma unroll
  for (int i = 0; i < W / 256; ++i) {
    const uint4 u = src[i * 32 + lane];
    const op2_t* h2 = reinterpret_cast<const op2_t*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[i * 8 + 2 * e] = op_to_float(h2[e].x);
      v[i * 8 + 2 * e + 1] = op_to_float(h2[e].y);
    }
  }
  const float2 st = row_stats_from_regs<W>(v, eps);
  if (lane == 0) stats[row] = st;
}

// stats[row] = (rstd, -mean*rstd) from the row's `nparts` partial (mean, M2), each over 128 columns — the merge the
// LayerNorm-folding GEMM does in its epilogue (gemm_tcgen05.cuh), as a kernel for the test seam
__global__ void __launch_bounds__(256)
merge_row_stats_kernel(const float2* __restrict__ parts, int nparts, float2* __restrict__ stats, int M, float eps) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  const float2* pp = parts + static_cast<size_t>(row) * nparts;
  float mean = 0.f, m2 = 0.f;
  for (int i = 0; i < nparts; ++i) mean += pp[i].x;
  mean *= 1.0f / static_cast<float>(nparts);
  for (int i = 0; i < nparts; ++i) {
    const float d = pp[i].x - mean;
    m2 += fmaf(128.0f * d, d, pp[i].y);
  }
  const float rstd = rsqrtf(m2 / (128.0f * static_cast<float>(nparts)) + eps);
  stats[row] = make_float2(rstd, -mean * rstd);
}

// ln_pre of the 16-bit path: x16 = fl16(LN(x0 fp32)) and the first block's ln_1 statistics of x16.
template <int W>
__global__ void __launch_bounds__(256)
layernorm_pre_r16_kernel(const float* __restrict__ x0, const float* __restrict__ gamma, const float* __restrict__ beta,
                         op_t* __restrict__ x16, float2* __restrict__ stats, int M, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  RowRegs<W> r;
  ln_row<W>(x0 + static_cast<size_t>(row) * W, gamma, beta, eps, r);
  uint2* out = reinterpret_cast<uint2*>(x16 + static_cast<size_t>(row) * W);
  float v[W / 32];
#pragma unroll
  for (int i = 0; i < RowRegs<W>::kVec; ++i) {
    const op2_t a = floats_to_op2(r.v[i].x, r.v[i].y), b = floats_to_op2(r.v[i].z, r.v[i].w);
    out[i * 32 + lane] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    v[4 * i] = op_to_float(a.x); v[4 * i + 1] = op_to_float(a.y);
    v[4 * i + 2] = op_to_float(b.x); v[4 * i + 3] = op_to_float(b.y);
  }
  const float2 st = row_stats_from_regs<W>(v, eps);
  if (lane == 0) stats[row] = st;
}

// Weight preparation for the folded LayerNorm (run once per weight load), one warp per output row n:
//   Wf[n,k] = bf16(W[n,k] * gamma[k]);  svec[n] = sum_k float(Wf[n,k]);  cvec[n] = sum_k W[n,k]*beta[k] + bias[n]
// svec uses the ROUNDED weights so that acc - mean*svec cancels the mean exactly as the tensor core saw it.
__global__ void __launch_bounds__(256)
fold_ln_weights_kernel(const float* __restrict__ Wt, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ bias, op_t* __restrict__ Wf, float* __restrict__ svec,
                       float* __restrict__ cvec, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f, c = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = Wt[static_cast<size_t>(n) * K + k];
    const op_t wf = float_to_op(w * gamma[k]);
    Wf[static_cast<size_t>(n) * K + k] = wf;
    s += op_to_float(wf);
    c = fmaf(w, beta[k], c);
  }
  s = warp_sum(s);
  c = warp_sum(c);
  if (lane == 0) {
    svec[n] = s;
    cvec[n] = c + bias[n];
  }
}

// taps(bf16)[B, ld_taps] columns [layer*W, (layer+1)*W) = x[b*tokens + 0, :]
// The CLS row of every resblock output; replaces the reference's forward hooks
// (finetune_module/utils.py:6-18, clip_multiscale_adapter.py:138-143).
// four consecutive elements of a residual-stream row (fp32 or 16-bit) as floats
__device__ __forceinline__ float4 load4_as_float(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4_as_float(const op_t* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const op2_t a = *reinterpret_cast<const op2_t*>(&u.x), b = *reinterpret_cast<const op2_t*>(&u.y);
  return make_float4(op_to_float(a.x), op_to_float(a.y), op_to_float(b.x), op_to_float(b.y));
}

template <int W, typename XT>
__global__ void __launch_bounds__(256)
gather_cls_bf16_kernel(const XT* __restrict__ x, op_t* __restrict__ taps, int B, int tokens,
                       int ld_taps, int col0) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const XT* src = x + static_cast<size_t>(b) * tokens * W;
  uint2* dst = reinterpret_cast<uint2*>(taps + static_cast<size_t>(b) * ld_taps + col0);
#pragma unroll
  for (int i = 0; i < W / 128; ++i) {
    const float4 v = load4_as_float(src + (i * 32 + lane) * 4);
    dst[i * 32 + lane] = make_uint2(pack_op(v.x, v.y), pack_op(v.z, v.w));
  }
}

// dst fp32 [B, ld_dst] columns [col0, col0+W) = x[b*tokens, :]  — the class-token rows of the residual stream (fp32 or
// 16-bit), compacted: the pruned last block and the heads (ld_dst = W), the fine-tuning taps (ld_dst = layers*W)
template <int W, typename XT>
__global__ void __launch_bounds__(256)
gather_cls_rows_f32_kernel(const XT* __restrict__ x, float* __restrict__ dst, int B, int tokens, int ld_dst, int col0) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const XT* src = x + static_cast<size_t>(b) * tokens * W;
  float4* out = reinterpret_cast<float4*>(dst + static_cast<size_t>(b) * ld_dst + col0);
#pragma unroll
  for (int i = 0; i < W / 128; ++i) out[i * 32 + lane] = load4_as_float(src + (i * 32 + lane) * 4);
}

// Attention for the CLASS-TOKEN query only (last resblock): one warp per (frame, head).
//   q     bf16 [B, W]           projected class-token queries (head h at columns h*64)
//   qkv   bf16 [B*L, 3W]        K at columns W + h*64, V at 2W + h*64 (the Q third is not read)
//   out   bf16 [B, W]
// Lanes own keys for the scores (fp32 dot products, fp32 softmax with the true max), then own 2 head dims for P V.
template <int L>
__global__ void __launch_bounds__(256)
cls_attention_kernel(const op_t* __restrict__ q, const op_t* __restrict__ qkv,
                     op_t* __restrict__ out, int B, int heads, int width, float scale) {
  __shared__ float s_p[8][L + 3];
  __shared__ float s_q[8][64];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 8 + w;
  if (item >= B * heads) return;
  const int frame = item / heads, head = item - frame * heads;
  {
    const op2_t v = reinterpret_cast<const op2_t*>(q + static_cast<size_t>(frame) * width + head * 64)[lane];
    s_q[w][2 * lane] = op_to_float(v.x) * scale;      // torch scales q by 1/sqrt(d) before q k^T
    s_q[w][2 * lane + 1] = op_to_float(v.y) * scale;
  }
  __syncwarp();
  const op_t* kbase = qkv + static_cast<size_t>(frame) * L * 3 * width + width + head * 64;
  const op_t* vbase = kbase + width;
  // scores: lane -> (key lane/8 of a group of 4, 16-byte chunk lane%8): every load instruction covers four whole 128-byte
  // K rows; the 8 partial dot products of a key are combined with 3 shuffles
  float m = -INFINITY;
  const int kq = lane >> 3, kc = lane & 7;
  float qv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qv[e] = s_q[w][kc * 8 + e];
#pragma unroll 4
  for (int j0 = 0; j0 < L; j0 += 4) {
    const int j = j0 + kq;
    float s = 0.f;
    if (j < L) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(j) * 3 * width) + kc);
      const op2_t* h2 = reinterpret_cast<const op2_t*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s = fmaf(qv[2 * e], op_to_float(h2[e].x), s);
        s = fmaf(qv[2 * e + 1], op_to_float(h2[e].y), s);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (kc == 0 && j < L) s_p[w][j] = s;
  }
  __syncwarp();
  for (int j = lane; j < L; j += 32) m = fmaxf(m, s_p[w][j]);
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = __expf(s_p[w][j] - m);
    s_p[w][j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
#pragma unroll 16      // 16 independent 128-byte row loads in flight per warp: the loop is pure load latency otherwise
  for (int j = 0; j < L; ++j) {
    const op2_t v = __ldg(reinterpret_cast<const op2_t*>(vbase + static_cast<size_t>(j) * 3 * width) + lane);
    const float pj = s_p[w][j];
    o0 = fmaf(pj, op_to_float(v.x), o0);
    o1 = fmaf(pj, op_to_float(v.y), o1);
  }
  const float inv = 1.0f / sum;
  reinterpret_cast<op2_t*>(out + static_cast<size_t>(frame) * width + head * 64)[lane] =
      floats_to_op2(o0 * inv, o1 * inv);
}

// fp32 [T,3,224,224] (already resized / augmented / normalised by the caller) -> bf16 patch rows of the patch-embed GEMM:
// A[b*tokens + 1 + py*G + px, c*P*P + ph*P + pw]; row b*tokens (class-token slot) = 0. One thread per 8 outputs.
__global__ void __launch_bounds__(256)
patchify_chw_bf16_kernel(const float* __restrict__ chw, op_t* __restrict__ A, int T, int P, int G, int tokens) {
  const int K = 3 * P * P;
  const size_t total8 = static_cast<size_t>(T) * tokens * K / 8;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t e = i * 8;
    const int col = static_cast<int>(e % K);          // 8 consecutive pw of one (c, ph): P is a multiple of 8
    const size_t row = e / K;
    const int tok = static_cast<int>(row % tokens);
    const size_t b = row / tokens;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (tok > 0) {
      const int py = (tok - 1) / G, px = (tok - 1) % G;
      const int c = col / (P * P), ph = (col / P) % P, pw = col % P;
      const float* src = chw + ((b * 3 + c) * 224 + py * P + ph) * 224 + px * P + pw;
      const float4 lo = *reinterpret_cast<const float4*>(src), hi = *reinterpret_cast<const float4*>(src + 4);
      o = make_uint4(pack_op(lo.x, lo.y), pack_op(lo.z, lo.w), pack_op(hi.x, hi.y), pack_op(hi.z, hi.w));
    }
    *reinterpret_cast<uint4*>(A + e) = o;
  }
}

// fp32 -> bf16 conversion of a contiguous buffer (weights at load time, adapter features).
__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, op_t* __restrict__ dst, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_op(v.x, v.y), pack_op(v.z, v.w));
  }
  if (i < n) {  // ragged tail (n % 4 != 0): at most one thread lands here
    for (size_t j = i; j < n; ++j) dst[j] = float_to_op(src[j]);
  }
}

}  // namespace arp

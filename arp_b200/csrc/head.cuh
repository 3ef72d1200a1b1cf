// Reward heads (SURVEY.md §2 K9 / K10 epilogues): everything after the last resblock, fused per frame.
//
//  clip head      f = ln_post(x[:,0]) @ proj ; reward = exp(logit_scale) * <f/|f|, t/|t|>
//                 (openai/CLIP VisionTransformer tail + CLIP.forward, called at label_reward.py:141-145)
//  adapter head   a = sigma(w) * feat + (1 - sigma(w)) * mlp(feat) ; L2-normalise ; same cosine
//                 (finetune_module/clip_multiscale_adapter.py:145-151, label_reward.py:213-228);
//                 "ensemble" variant normalises each of the 13 512-wide scales separately and
//                 averages the 13 cosines (label_reward.py:217-222).
//  Text embeddings arrive already L2-normalised (cached once per run; the reference re-runs the
//  text tower every episode, label_reward.py:135-138). All n_text cosines are produced; `reduce`
//  picks row 0 (what the reference actually does — SURVEY.md Q1) or the mean.
// fp32 throughout: these are a few hundred KFLOP per frame, L2-resident weights.
#pragma once

#include "common.cuh"

namespace arp {

constexpr int HEAD_MAX_TEXT = 16;
enum HeadReduce : int { REDUCE_FIRST = 0, REDUCE_MEAN = 1 };

template <int NT>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // all NT threads call; returns the total to every thread
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NT / 32; ++i) t += scratch[i];
  return t;
}

// One CTA (256 threads) per FR consecutive frames: the 768x512 projection is streamed once per CTA and applied to all
// FR class-token rows (with one frame per CTA the 1.5 MB matrix was re-read from L2 by every CTA).
//   x        fp32 [B*tokens, W]   residual stream after the last block (row b*tokens = class token)
//   proj     fp32 [W, E]
//   text     fp32 [n_text, E] unit rows (may be null when only features are wanted)
//   feat_out fp32 [B, ld_feat] : un-normalised f written at column feat_col0 (adapter / goal modes) or null
//   logits   fp32 [B, n_text] or null ; reward fp32 [B] or null
constexpr int HEAD_FR = 4;
constexpr int HEAD_SMEM_BYTES = 8 * HEAD_FR * 512 * 4;   // dynamic: the per-warp partial outputs of clip_head_kernel

template <int W, int E>
__global__ void __launch_bounds__(256)
clip_head_kernel(const float* __restrict__ x, int tokens, const float* __restrict__ ln_g,
                 const float* __restrict__ ln_b, float eps, const float* __restrict__ proj,
                 const float* __restrict__ text, int n_text, float scale, int reduce,
                 float* __restrict__ feat_out, int ld_feat, int feat_col0, float* __restrict__ logits,
                 float* __restrict__ reward, int n_frames) {
  static_assert(W % 256 == 0 && E % 256 == 0, "head widths must be multiples of the CTA size");
  __shared__ float s_f[HEAD_FR][W];
  __shared__ float s_y[HEAD_FR][E];
  __shared__ float s_red[8];
  __shared__ float s_inv[HEAD_FR];
  __shared__ float s_logit[HEAD_FR][HEAD_MAX_TEXT];
  const int b0 = blockIdx.x * HEAD_FR, tid = threadIdx.x;
  const int nf = min(HEAD_FR, n_frames - b0);

  for (int f = 0; f < HEAD_FR; ++f) {        // ln_post of each class-token row (zeros for frames past the end)
    float v[W / 256];
    float s = 0.f;
    if (f < nf) {
      const float* xr = x + static_cast<size_t>(b0 + f) * tokens * W;
#pragma unroll
      for (int i = 0; i < W / 256; ++i) { v[i] = xr[tid + i * 256]; s += v[i]; }
    } else {
#pragma unroll
      for (int i = 0; i < W / 256; ++i) v[i] = 0.f;
    }
    const float mean = block_sum<256>(s, s_red) * (1.0f / W);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < W / 256; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum<256>(q, s_red) * (1.0f / W) + eps);
#pragma unroll
    for (int i = 0; i < W / 256; ++i) {
      const int c = tid + i * 256;
      s_f[f][c] = f < nf ? (v[i] - mean) * rstd * ln_g[c] + ln_b[c] : 0.f;
    }
  }
  __syncthreads();

  // projection, split over k across the 8 warps: a warp streams whole 2 KB rows of proj (4 x LDG.128 per lane, 16 loads
  // in flight) for its 96 values of k and keeps FR x 16 partial outputs per lane; the partials meet in shared memory.
  // (One output pair per thread over all 768 k was a chain of 192 dependent L2 round trips: ~100 us per CTA.)
  static_assert(E == 512 && W % 8 == 0, "lane -> column mapping below assumes 512 outputs");
  extern __shared__ float s_part[];            // [8 warps][HEAD_FR][E]
  {
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int KW = W / 8;
    float acc[HEAD_FR][16];
#pragma unroll
    for (int f = 0; f < HEAD_FR; ++f)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[f][j] = 0.f;
    const float4* prow = reinterpret_cast<const float4*>(proj) + static_cast<size_t>(warp) * KW * (E / 4) + lane;
#pragma unroll 4
    for (int kk = 0; kk < KW; ++kk) {
      float4 pv[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) pv[g] = __ldg(prow + static_cast<size_t>(kk) * (E / 4) + g * 32);
#pragma unroll
      for (int f = 0; f < HEAD_FR; ++f) {
        const float fk = s_f[f][warp * KW + kk];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          acc[f][4 * g] = fmaf(fk, pv[g].x, acc[f][4 * g]);
          acc[f][4 * g + 1] = fmaf(fk, pv[g].y, acc[f][4 * g + 1]);
          acc[f][4 * g + 2] = fmaf(fk, pv[g].z, acc[f][4 * g + 2]);
          acc[f][4 * g + 3] = fmaf(fk, pv[g].w, acc[f][4 * g + 3]);
        }
      }
    }
#pragma unroll
    for (int f = 0; f < HEAD_FR; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g)
        reinterpret_cast<float4*>(s_part + (static_cast<size_t>(warp) * HEAD_FR + f) * E)[g * 32 + lane] =
            make_float4(acc[f][4 * g], acc[f][4 * g + 1], acc[f][4 * g + 2], acc[f][4 * g + 3]);
  }
  __syncthreads();
  for (int f = 0; f < HEAD_FR; ++f) {
    float n2 = 0.f;
#pragma unroll
    for (int j = 0; j < E / 256; ++j) {
      const int col = tid + j * 256;
      float yv = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) yv += s_part[(static_cast<size_t>(w8) * HEAD_FR + f) * E + col];
      s_y[f][col] = yv;
      n2 += yv * yv;
      if (feat_out && f < nf) feat_out[static_cast<size_t>(b0 + f) * ld_feat + feat_col0 + col] = yv;
    }
    const float tot = block_sum<256>(n2, s_red);
    if (tid == 0) s_inv[f] = 1.0f / sqrtf(tot);
  }
  if (text == nullptr) return;
  __syncthreads();

  // cosines: warp w handles (frame, text) pairs w, w+8, ...
  const int warp = tid >> 5, lane = tid & 31;
  for (int p = warp; p < nf * n_text; p += 8) {
    const int f = p / n_text, t = p - f * n_text;
    float d = 0.f;
    for (int k = lane; k < E; k += 32) d = fmaf(s_y[f][k], __ldg(text + static_cast<size_t>(t) * E + k), d);
    d = warp_sum(d);
    if (lane == 0) {
      const float lg = scale * (d * s_inv[f]);
      s_logit[f][t] = lg;
      if (logits) logits[static_cast<size_t>(b0 + f) * n_text + t] = lg;
    }
  }
  __syncthreads();
  if (tid < nf && reward) {
    float r = s_logit[tid][0];
    if (reduce == REDUCE_MEAN) {
      for (int t = 1; t < n_text; ++t) r += s_logit[tid][t];
      r /= static_cast<float>(n_text);
    }
    reward[b0 + tid] = r;
  }
}

// Adapter gate + normalise + cosine. One CTA per frame, one warp per 512-wide scale (S warps).
//   feat, mlp fp32 [B, S*E]; text fp32 [n_text, S*E] (unit over S*E, or per-scale unit when ensemble)
template <int S, int E>
__global__ void __launch_bounds__(S * 32)
adapter_head_kernel(const float* __restrict__ feat, const float* __restrict__ mlp, float res,
                    const float* __restrict__ text, int n_text, float scale, int ensemble, int reduce,
                    float* __restrict__ logits, float* __restrict__ reward, float* __restrict__ adapted_out) {
  __shared__ float s_n2[S];
  __shared__ float s_dot[S][HEAD_MAX_TEXT];
  const int b = blockIdx.x, sidx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = static_cast<size_t>(b) * S * E + static_cast<size_t>(sidx) * E;
  float a[E / 32];
  float n2 = 0.f;
#pragma unroll
  for (int i = 0; i < E / 32; ++i) {
    const int c = lane + i * 32;
    a[i] = res * feat[base + c] + (1.0f - res) * mlp[base + c];
    n2 += a[i] * a[i];
    if (adapted_out) adapted_out[base + c] = a[i];
  }
  n2 = warp_sum(n2);
  if (lane == 0) s_n2[sidx] = n2;
  for (int t = 0; t < n_text; ++t) {
    const float* tr = text + static_cast<size_t>(t) * S * E + static_cast<size_t>(sidx) * E;
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < E / 32; ++i) d = fmaf(a[i], __ldg(tr + lane + i * 32), d);
    d = warp_sum(d);
    if (lane == 0) s_dot[sidx][t] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float racc = 0.f;
    const int n_use = reduce == REDUCE_MEAN ? n_text : 1;
    for (int t = 0; t < n_text; ++t) {
      float lg;
      if (ensemble) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += s_dot[s][t] / fmaxf(sqrtf(s_n2[s]), 1e-12f);
        lg = scale * acc / static_cast<float>(S);
      } else {
        float tot = 0.f, d = 0.f;
        for (int s = 0; s < S; ++s) { tot += s_n2[s]; d += s_dot[s][t]; }
        lg = scale * (d / fmaxf(sqrtf(tot), 1e-12f));
      }
      if (logits) logits[static_cast<size_t>(b) * n_text + t] = lg;
      if (t < n_use) racc += lg;
    }
    if (reward) reward[b] = racc / static_cast<float>(n_use);
  }
}

// Goal-conditioned reward (label_reward.py:148-163,180-196): r_t = sign * || f_t - f_goal ||_2 where the
// goal is the LAST frame of the episode. One warp per frame. sign = -1 for "clip_goal_conditioned",
// +1 for the adapter "_goal_conditioned" variant (the reference omits the minus there, :193-195).
__global__ void __launch_bounds__(256)
goal_distance_kernel(const float* __restrict__ feat, int dim, const long long* __restrict__ frame_goal,
                     long long T, float sign, float* __restrict__ reward) {
  const long long t = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const int lane = threadIdx.x & 31;
  const float* a = feat + t * dim;
  const float* g = feat + frame_goal[t] * dim;
  float s = 0.f;
  for (int k = lane; k < dim; k += 32) { const float d = a[k] - g[k]; s = fmaf(d, d, s); }
  s = warp_sum(s);
  if (lane == 0) reward[t] = sign * sqrtf(s);
}

}  // namespace arp

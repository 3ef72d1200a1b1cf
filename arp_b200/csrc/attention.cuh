// Multi-head self-attention core for the CLIP vision tower:
//   O = softmax(Q K^T / sqrt(64)) V  per (frame, head), L = 197 (ViT-B/16) or 50 (ViT-B/32) tokens, no mask.
// Reference semantics: nn.MultiheadAttention inside openai/CLIP ResidualAttentionBlock
// (in-tree mirror arp_dt/models/openai/layers.py:235-250); called via arp_dt/label_reward.py:141.
//
// v1 kernel: one CTA per (frame, head); K and V of the head are staged in shared memory
// (padded rows, conflict-free ldmatrix), each warp owns 16-query tiles, the whole
// 16 x L score strip stays in registers (no online-softmax rescaling needed at L <= 208),
// P is re-used as the A fragment of the PV product (bf16 mma.sync m16n8k16, fp32 accumulate).
// Attention is 4.1 % of the path's FLOPs; the GEMMs run on tcgen05.
#pragma once

#include "common.cuh"

namespace arp {

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
#if ARP_OP_FP16
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
#else
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
#endif
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int ATT_DH = 64;
constexpr int ATT_ROW = ATT_DH + 8;  // padded smem row (144 B): 8 consecutive rows hit 32 distinct banks
constexpr int ATT_THREADS = 128;

template <int L>
struct AttnCfg {
  static constexpr int LP = (L + 15) / 16 * 16;  // padded token count (208 / 64)
  static constexpr int NT = LP / 8;              // key tiles of 8
  static constexpr int MT = LP / 16;             // query tiles of 16
  static constexpr int SMEM = 2 * LP * ATT_ROW * 2;
};

// qkv: bf16 [B*L, 3*width], row = frame*L + token, columns [q | k | v], head h at h*64 inside each.
// out: bf16 [B*L, width].
template <int L>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const op_t* __restrict__ qkv, op_t* __restrict__ out, int width,
                 float scale_log2e) {
  using C = AttnCfg<L>;
  extern __shared__ __align__(16) uint8_t att_smem[];
  op_t* sK = reinterpret_cast<op_t*>(att_smem);
  op_t* sV = sK + C::LP * ATT_ROW;

  const int head = blockIdx.x;
  const int frame = blockIdx.y;
  const int ld = 3 * width;
  const op_t* base = qkv + static_cast<size_t>(frame) * L * ld + head * ATT_DH;

  // ---- stage K and V (16-byte loads; padded tail rows zeroed so 0 * garbage never makes a NaN) ----
  for (int i = threadIdx.x; i < C::LP * 8; i += ATT_THREADS) {
    const int row = i >> 3, ch = i & 7;
    uint4 k = make_uint4(0, 0, 0, 0), v = make_uint4(0, 0, 0, 0);
    if (row < L) {
      const op_t* p = base + static_cast<size_t>(row) * ld + ch * 8;
      k = *reinterpret_cast<const uint4*>(p + width);
      v = *reinterpret_cast<const uint4*>(p + 2 * width);
    }
    *reinterpret_cast<uint4*>(sK + row * ATT_ROW + ch * 8) = k;
    *reinterpret_cast<uint4*>(sV + row * ATT_ROW + ch * 8) = v;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  // ldmatrix row providers: lanes 0-7 -> matrix 0, 8-15 -> 1, 16-23 -> 2, 24-31 -> 3
  const int lm_r = lane & 7, lm_m = lane >> 3;

  for (int mt = warp; mt < C::MT; mt += ATT_THREADS / 32) {
    const int q0 = mt * 16;
    // ---- Q fragments straight from global (A operand, row-major 16x64) ----
    uint32_t qa[4][4];
    {
      const int r0 = q0 + g, r1 = q0 + g + 8;
      const op_t* p0 = base + static_cast<size_t>(r0) * ld;
      const op_t* p1 = base + static_cast<size_t>(r1) * ld;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c = ks * 16 + 2 * t;
        qa[ks][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + c) : 0u;
        qa[ks][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + c) : 0u;
        qa[ks][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + c + 8) : 0u;
        qa[ks][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + c + 8) : 0u;
      }
    }
    // ---- S = Q K^T ----
    float s[C::NT][4];
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        // matrices: (keys nt*8.., dh kp*32 + {0,8,16,24})
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = sK_u + ((nt * 8 + lm_r) * ATT_ROW + kp * 32 + lm_m * 8) * 2;
        ldmatrix_x4(addr, b0, b1, b2, b3);
        mma_bf16_16816(s[nt], qa[kp * 2], b0, b1);
        mma_bf16_16816(s[nt], qa[kp * 2 + 1], b2, b3);
      }
    }
    // ---- softmax over keys (rows g and g+8; 4 lanes share a row) ----
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      const int key = nt * 8 + 2 * t;
      if (key >= L) s[nt][0] = s[nt][2] = -INFINITY;
      if (key + 1 >= L) s[nt][1] = s[nt][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float o0 = m0 * scale_log2e, o1 = m1 * scale_log2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * scale_log2e - o0);
      s[nt][1] = exp2f(s[nt][1] * scale_log2e - o0);
      s[nt][2] = exp2f(s[nt][2] * scale_log2e - o1);
      s[nt][3] = exp2f(s[nt][3] * scale_log2e - o1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- O = P V ----
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < C::MT; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_op(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_op(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_op(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_op(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        // matrices: keys kk*16 + {0..7, 8..15} x dh (np*16 + {0, 8}); transposed on load
        uint32_t b0, b1, b2, b3;
        const uint32_t addr =
            sV_u + ((kk * 16 + (lm_m & 1) * 8 + lm_r) * ATT_ROW + np * 16 + (lm_m >> 1) * 8) * 2;
        ldmatrix_x4_trans(addr, b0, b1, b2, b3);
        mma_bf16_16816(o[np * 2], pa, b0, b1);
        mma_bf16_16816(o[np * 2 + 1], pa, b2, b3);
      }
    }
    // ---- normalise and store ----
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    op_t* ob = out + static_cast<size_t>(frame) * L * width + head * ATT_DH + 2 * t;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      if (r0 < L)
        *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * width + nd * 8) =
            pack_op(o[nd][0] * i0, o[nd][1] * i0);
      if (r1 < L)
        *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * width + nd * 8) =
            pack_op(o[nd][2] * i1, o[nd][3] * i1);
    }
  }
}

}  // namespace arp

// Attention core on the 5th-gen tensor cores:  O = softmax(Q K^T / 8) V  per (frame, head).
// Reference semantics: nn.MultiheadAttention inside openai/CLIP's ResidualAttentionBlock (no mask for the
// vision tower), reached from arp_dt/label_reward.py:141.
//
// Persistent CTAs loop over (frame, head) items. For one item (L = 197 tokens, padded to 208 keys / 2x128 queries):
//   TMA      Q (two 128x64 tiles), K and V (208x64) of the head -> 128B-swizzled smem, double buffered across items
//   MMA #1   S[128 x 208] = Q K^T     tcgen05.mma kind::f16, A and B from smem (K-major), fp32 accumulators in TMEM
//   softmax  one thread per query row: tcgen05.ld the row, max / exp2 / sum in registers — no cross-lane traffic —
//            and write P back INTO the same TMEM columns as packed bf16 (tcgen05.st)
//   MMA #2   O[128 x 64] = P V        A operand straight from TMEM, B = V from smem as an MN-major operand
//   epilogue tcgen05.ld O, scale by 1/rowsum, bf16, one full 128-byte line per thread to HBM
// TMEM map per query tile t (base = 256 t): S fp32 [base, base+208) -> P bf16x2 [base, base+104); O fp32 [base+128, base+192).
// Warp roles (320 threads): 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..9 = softmax/epilogue
// (warps 2-5 own query tile 0, warps 6-9 query tile 1; a warp may only touch TMEM lanes 32*(warp%4)..+31).
#pragma once

#include "common.cuh"

namespace arp {

constexpr int ATC_THREADS = 320;
constexpr int ATC_DH = 64;

template <int L>
struct AtcCfg {
  static constexpr int NK = (L + 15) / 16 * 16;        // padded keys: 208 (L=197), 64 (L=50)
  static constexpr int QT = (L + 127) / 128;           // query tiles: 2 / 1
  static constexpr int Q_BYTES = 128 * ATC_DH * 2;     // 16 KB per query tile
  static constexpr int KV_BYTES = NK * ATC_DH * 2;     // 26 KB
  static constexpr int KV_PAD = (KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int BUF_BYTES = QT * Q_BYTES + 2 * KV_PAD;
  static constexpr int TX_BYTES = QT * Q_BYTES + 2 * KV_BYTES;
  static constexpr int SMEM_BYTES = 2 * BUF_BYTES + 1024 + 256;
  static constexpr int O_COL = 128;                    // O accumulator column offset inside a tile's TMEM region
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major shared-memory operand, 128B swizzle: rows (the K index of the MMA) are 128 B = 64 elements of the
// MN index; groups of 8 rows are 1024 B apart (SBO). One 64-wide MN span -> LBO unused.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// qkv: bf16 [rows, 3*width] (tensor maps: box 64x128 for Q, 64xNK for K/V); out: bf16 [B*L, width]
//
// Each query tile of an item is an independent "job" with its own TMEM slot (256 columns) and its own
// S/P/O barriers, so the MMA warp can run tile 1's Q K^T while tile 0's rows are in softmax, and tile 0's
// P V while tile 1's rows are in softmax.
template <int L>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    __nv_bfloat16* __restrict__ out, int n_frames, int heads, int width, float scale_log2e,
                    int reverse) {
  using C = AtcCfg<L>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * C::BUF_BYTES);
  uint64_t* smem_full = bars;        // [2] TMA -> MMA
  uint64_t* smem_empty = bars + 2;   // [2] MMA (all P V of the item retired) -> TMA
  uint64_t* s_full = bars + 4;       // [2 slots] MMA -> softmax: S ready
  uint64_t* p_full = bars + 6;       // [2] softmax -> MMA: P written
  uint64_t* o_full = bars + 8;       // [2] MMA -> epilogue: O ready
  uint64_t* tmem_free = bars + 10;   // [2] epilogue -> MMA: O drained, slot reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_frames * heads;
  constexpr int SM_WARPS = 4 * C::QT;   // softmax warps that actually own rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&smem_full[i], 1);
      mbar_init(&smem_empty[i], C::QT);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&tmem_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&smem_empty[b], ph ^ 1);
      if (lane == 0) {
        const int item_o = reverse ? n_items - 1 - item : item;   // snake order across kernels (L2 reuse)
        const int frame = item_o / heads, head = item_o - frame * heads;
        uint8_t* buf = smem + b * C::BUF_BYTES;
        const int row = frame * L;
        mbar_arrive_expect_tx(&smem_full[b], C::TX_BYTES);
#pragma unroll
        for (int t = 0; t < C::QT; ++t)
          tma_load_2d(buf + t * C::Q_BYTES, &tmap_q, &smem_full[b], head * ATC_DH, row + t * 128);
        tma_load_2d(buf + C::QT * C::Q_BYTES, &tmap_kv, &smem_full[b], width + head * ATC_DH, row);
        tma_load_2d(buf + C::QT * C::Q_BYTES + C::KV_PAD, &tmap_kv, &smem_full[b], 2 * width + head * ATC_DH, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, C::NK);          // Q K^T: both operands K-major
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, ATC_DH, 0, 1);   // P V: A from TMEM, B = V MN-major
    auto issue_s = [&](uint32_t sbuf, int t) {
      const uint64_t dk = umma_desc_kmajor_sw128(sbuf + C::QT * C::Q_BYTES);
      const uint64_t dq = umma_desc_kmajor_sw128(sbuf + t * C::Q_BYTES);
#pragma unroll
      for (int k = 0; k < ATC_DH / 16; ++k)
        umma_bf16_ss(tmem_base + t * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      umma_commit(&s_full[t]);
    };
    auto issue_pv = [&](uint32_t sbuf, int t) {
      const uint64_t dv = umma_desc_mnmajor_sw128(sbuf + C::QT * C::Q_BYTES + C::KV_PAD);
#pragma unroll
      for (int k = 0; k < C::NK / 16; ++k)
        // 16 keys per MMA: A advances 8 packed columns, V advances 16 rows x 128 B = 2048 B
        umma_bf16_ts(tmem_base + t * 256 + C::O_COL, tmem_base + t * 256 + k * 8, dv + k * (2048 >> 4), idesc_o, k != 0);
      umma_commit(&o_full[t]);
    };
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph_buf = (it >> 1) & 1, ph = it & 1;
      const uint32_t sbuf = smem_u32(smem + b * C::BUF_BYTES);
      mbar_wait(&smem_full[b], ph_buf);
#pragma unroll
      for (int t = 0; t < C::QT; ++t) {
        mbar_wait(&tmem_free[t], ph ^ 1);     // previous item's O of this slot has been drained
        tc_fence_after();
        if (lane == 0) issue_s(sbuf, t);
        __syncwarp();
      }
#pragma unroll
      for (int t = 0; t < C::QT; ++t) {
        mbar_wait(&p_full[t], ph);
        tc_fence_after();
        if (lane == 0) {
          issue_pv(sbuf, t);
          umma_commit(&smem_empty[b]);        // the item's smem is free once BOTH slots' P V have retired (count = QT)
        }
        __syncwarp();
      }
    }
  } else if (warp - 2 < SM_WARPS) {
    // ===================== softmax + epilogue: one thread per query row =====================
    const int qt = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int qrow = qt * 128 + quarter * 32 + lane;     // query index inside the frame
    const uint32_t t_s = tmem_base + qt * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    constexpr int NFULL = C::NK / 32;                    // full 32-column chunks
    constexpr int TAIL = C::NK % 32;                     // 16 or 0
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int item_o = reverse ? n_items - 1 - item : item;
      const int frame = item_o / heads, head = item_o - frame * heads;
      mbar_wait(&s_full[qt], ph);
      tc_fence_after();
      // Only the last (partial or padded) chunk can contain key columns >= L; the others need no masking.
      constexpr int NCLEAN = (L / 32 < NFULL) ? L / 32 : NFULL;   // full chunks whose 32 columns are all real keys
      // ---- pass 1: row maximum over the L real keys (two 32-column loads in flight per wait) ----
      float m = -INFINITY;
      {
        uint32_t a[32], bq[32];
#pragma unroll 1
        for (int c = 0; c + 1 < NCLEAN; c += 2) {
          tmem_ld_32x32(t_s + c * 32, a);
          tmem_ld_32x32(t_s + c * 32 + 32, bq);
          tmem_ld_wait();
          float m0 = m, m1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(a[j]), __uint_as_float(a[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(bq[j]), __uint_as_float(bq[j + 1])));
          }
          m = fmaxf(m0, m1);
        }
        if (NCLEAN & 1) {
          tmem_ld_32x32(t_s + (NCLEAN - 1) * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(a[j]), __uint_as_float(a[j + 1])));
        }
#pragma unroll
        for (int c = NCLEAN; c < NFULL; ++c) {        // masked full chunks (only when L is not past them, e.g. L = 50)
          tmem_ld_32x32(t_s + c * 32, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < L) m = fmaxf(m, __uint_as_float(a[j]));
        }
        if (TAIL) {
          uint32_t r[16];
          tmem_ld_32x16(t_s + NFULL * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (NFULL * 32 + j < L) m = fmaxf(m, __uint_as_float(r[j]));
        }
      }
      const float mo = m * scale_log2e;
      // ---- pass 2: p = exp2(s*scale - max*scale); P (bf16 pairs) overwrites the S columns it came from.
      //      Two register buffers ping-pong: the load of the next chunk is in flight while this one is processed. ----
      float sum0 = 0.f, sum1 = 0.f;
      auto soft_chunk = [&](const uint32_t (&src)[32], int c, bool masked) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(src[j]), scale_log2e, -mo));
          float p1 = ex2_approx(fmaf(__uint_as_float(src[j + 1]), scale_log2e, -mo));
          if (masked) {
            if (c * 32 + j >= L) p0 = 0.f;
            if (c * 32 + j + 1 >= L) p1 = 0.f;
          }
          sum0 += p0;
          sum1 += p1;
          pk[j >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_32x16(t_s + c * 16, pk);
      };
      {
        uint32_t b0[32], b1[32];
        tmem_ld_32x32(t_s, b0);
        tmem_ld_wait();
#pragma unroll 1
        for (int c = 0; c + 1 < NFULL; c += 2) {
          tmem_ld_32x32(t_s + (c + 1) * 32, b1);            // chunk c+1 in flight
          soft_chunk(b0, c, c >= NCLEAN);
          tmem_ld_wait();
          if (c + 2 < NFULL) tmem_ld_32x32(t_s + (c + 2) * 32, b0);   // chunk c+2 in flight
          soft_chunk(b1, c + 1, c + 1 >= NCLEAN);
          tmem_ld_wait();
        }
        if (NFULL & 1) soft_chunk(b0, NFULL - 1, NFULL - 1 >= NCLEAN);
        if (TAIL) {
          uint32_t r[16], pk[8];
          tmem_ld_32x16(t_s + NFULL * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float p0 = (NFULL * 32 + j < L) ? ex2_approx(fmaf(__uint_as_float(r[j]), scale_log2e, -mo)) : 0.f;
            const float p1 = (NFULL * 32 + j + 1 < L) ? ex2_approx(fmaf(__uint_as_float(r[j + 1]), scale_log2e, -mo)) : 0.f;
            sum0 += p0;
            sum1 += p1;
            pk[j >> 1] = pack_bf16(p0, p1);
          }
          tmem_st_32x8(t_s + NFULL * 16, pk);
        }
      }
      const float sum = sum0 + sum1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[qt]);
      // ---- epilogue: O / rowsum -> bf16 -> HBM (each thread owns one 128-byte output line) ----
      mbar_wait(&o_full[qt], ph);
      tc_fence_after();
      const float inv = 1.0f / sum;
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_s + C::O_COL, o0);
      tmem_ld_32x32(t_s + C::O_COL + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[qt]);
      if (qrow < L) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(frame) * L + qrow) * width + head * ATC_DH);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16(__uint_as_float(o0[8 * c]) * inv, __uint_as_float(o0[8 * c + 1]) * inv),
                              pack_bf16(__uint_as_float(o0[8 * c + 2]) * inv, __uint_as_float(o0[8 * c + 3]) * inv),
                              pack_bf16(__uint_as_float(o0[8 * c + 4]) * inv, __uint_as_float(o0[8 * c + 5]) * inv),
                              pack_bf16(__uint_as_float(o0[8 * c + 6]) * inv, __uint_as_float(o0[8 * c + 7]) * inv));
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[4 + c] = make_uint4(pack_bf16(__uint_as_float(o1[8 * c]) * inv, __uint_as_float(o1[8 * c + 1]) * inv),
                                  pack_bf16(__uint_as_float(o1[8 * c + 2]) * inv, __uint_as_float(o1[8 * c + 3]) * inv),
                                  pack_bf16(__uint_as_float(o1[8 * c + 4]) * inv, __uint_as_float(o1[8 * c + 5]) * inv),
                                  pack_bf16(__uint_as_float(o1[8 * c + 6]) * inv, __uint_as_float(o1[8 * c + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace arp

// Persistent warp-specialised bf16 GEMM on the 5th-gen tensor cores:
//   C[M,N] = epilogue( A[M,K] · W[N,K]^T )        (A, W bf16 K-major; fp32 accumulate in TMEM)
//
// This is the one dense-contraction kernel behind every linear layer on the
// reward-labeling path (SURVEY.md §2 K2,K4,K6,K7,K8,K10; reference call sites:
// openai/CLIP VisionTransformer / ResidualAttentionBlock as invoked from
// arp_dt/label_reward.py:141 and finetune_module/clip_multiscale_adapter.py:135-151).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer     cp.async.bulk.tensor A[128x64], W[256x64] tiles, SW128, 4-stage mbarrier ring
//   warp 1      MMA issuer       tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, one elected lane
//   warps 2..5  epilogue         tcgen05.ld 32x32b.x32 -> bias / periodic row table / residual / activation -> HBM
// Two 256-column accumulator stages in TMEM (all 512 columns) let the epilogue
// of tile i drain while tile i+1 is being multiplied.
#pragma once

#include "common.cuh"

namespace arp {

enum GemmAct : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_RELU = 2 };

struct GemmArgs {
  int M, N, K;
  void* out;            // [M, ldo] bf16 or fp32
  int ldo;              // elements
  const float* bias;    // [N] or nullptr
  const float* resid;   // fp32 [M, ldr] or nullptr; may alias out (in-place residual stream)
  int ldr;
  const float* rowtab;  // fp32 [period, N] added to row (row % period), or nullptr (pos-emb + cls)
  int period;
  // ---- LayerNorm fused behind the residual GEMM (gemm2 MODE 1 only; all null = off) ----
  // After a CTA's TMA reduce-adds of a tile have completed it bumps ln_cnt[row block of 128]; the CTA that brings the
  // count to N/256 (every column tile of those rows is in) normalises the 128 rows right there — they are still in
  // L2 — and writes ln_out (bf16 [M, N]) = LN(out rows; ln_gamma, ln_beta). Replaces a standalone LayerNorm kernel
  // that re-read the whole fp32 residual stream from HBM.
  const float* ln_gamma;
  const float* ln_beta;
  op_t* ln_out;
  int* ln_cnt;          // [ceil(M/128)] zeroed before the launch
  int reverse;          // gemm2: walk the tiles from the last row block to the first (snake order across kernels,
                        // so a kernel starts on the rows its predecessor wrote last — still in L2)
  // ---- LayerNorm fold (gemm2 MODE 2 / 3) ----
  // MODE 2 (residual + LN statistics): x = resid + acc + bias is written to `out` (fp32, may alias resid), a bf16
  //   copy to `xb` [M, N], and per-row partial moments (sum, sum of squares over each 128-column span) to
  //   `stats_out` [M, 2*N/128]. Every partial is written by exactly one warp: no atomics, deterministic.
  // MODE 3 (LN applied algebraically): A was the RAW bf16 residual stream and W' = W*diag(gamma), so
  //   LN(x) W^T + bias = rstd_r * (acc - mean_r * svec_n) + cvec_n with svec = W' 1, cvec = W beta + bias;
  //   mean_r / rstd_r come from `stats_in` [M, 2*stats_nh] (moments over K = 128*stats_nh columns).
  op_t* xb;
  float* stats_out;
  const float* stats_in;
  int stats_nh;
  const float* svec;
  const float* cvec;
  float eps;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;  // 64 bf16 = one 128-byte swizzle atom
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;  // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

template <typename OutT, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + GEMM_STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * GEMM_STAGES;    // [2]       MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (args.M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = args.N / GEMM_BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = args.K / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < GEMM_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + stage * GEMM_STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], GEMM_STAGE_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BK, n_blk * GEMM_BN);
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, GEMM_BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * GEMM_STAGE_BYTES);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address field
            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + quarter * 32 + lane;
      const bool row_ok = row < args.M;
      const uint32_t taddr = tmem_base + acc * GEMM_BN + (static_cast<uint32_t>(quarter * 32) << 16);
      const float* tab_row =
          args.rowtab ? args.rowtab + static_cast<size_t>(row % args.period) * args.N : nullptr;
      const float* res_row = args.resid ? args.resid + static_cast<size_t>(row) * args.ldr : nullptr;
      OutT* out_row = reinterpret_cast<OutT*>(args.out) + static_cast<size_t>(row) * args.ldo;
#pragma unroll 1
      for (int c = 0; c < GEMM_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * GEMM_BN + c * 32;
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (args.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (tab_row) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(tab_row + n0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (ACT == ACT_QUICKGELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.0f + __expf(-1.702f * v[j]));
          } else if (ACT == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (res_row) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(res_row + n0 + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (sizeof(OutT) == 4) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out_row) + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<op_t*>(out_row) + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = make_uint4(pack_op(v[8 * j], v[8 * j + 1]), pack_op(v[8 * j + 2], v[8 * j + 3]),
                                pack_op(v[8 * j + 4], v[8 * j + 5]), pack_op(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace arp

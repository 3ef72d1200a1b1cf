// Persistent warp-specialised 16-bit GEMM on the 5th-gen tensor cores:
//   C[M,N] = epilogue( A[M,K] · W[N,K]^T )        (A, W in the operand format op_t, K-major; fp32 accumulate in TMEM)
//
// The one dense-contraction kernel behind every linear layer on the reward-labeling path (SURVEY.md §2
// K2,K4,K6,K7,K8,K10; reference call sites: openai/CLIP VisionTransformer / ResidualAttentionBlock as invoked from
// arp_dt/label_reward.py:141 and finetune_module/clip_multiscale_adapter.py:135-151).
//
//   * CTA pairs (CG = 2): tcgen05.mma.cta_group::2 with M = 256 across the two SMs of a cluster, each CTA staging
//     its own 128 A rows and half of the W tile — halves shared-memory and L2 traffic for the B operand.
//   * warp roles (320 threads): 0 = TMA producer (6-stage mbarrier ring), 1 = MMA issuer + TMEM owner,
//     2..9 = epilogue (warp w reads TMEM lane quarter w%4; warps 2-5 take accumulator columns 0-127, 6-9 columns
//     128-255). Two 256-column accumulator stages in TMEM: tile i drains while tile i+1 is multiplied.
//   * the epilogue never issues an uncoalesced global access: 32x128-byte sub-tiles go TMEM -> registers (bias /
//     periodic row table / LayerNorm fold / activation) -> 128B-swizzled shared memory -> TMA store, or
//   * G2_REDUCE: the residual stream is updated by TMA reduce-add (cp.reduce.async.bulk.tensor .add, performed at
//     L2): x += tile without ever reading x into the SM. Works on the fp32 stream (precision mode 2, and the
//     class-token rows of the pruned last block) and on the 16-bit stream of the default path
//     (x = fl16(x + fl16(acc + bias)) — the same two roundings the reference's own fp16 CUDA route performs).
//   * G2_RESID_STATS: the 16-bit residual update of the default path, x = fl16(x + acc + bias) with ONE rounding, done in
//     registers: the warp requests its 32x128-byte tile of x with fully coalesced 16-byte loads BEFORE it waits for the
//     accumulator, turns it into the row-per-lane layout through its staging buffer, adds, and stores through the same
//     staging / TMA path — accumulating the mean and centred second moment of the ROUNDED values of its 128 columns on the
//     way (two-pass per 32 columns, Chan's merge). Each warp writes one partial (mean, M2) per row and tile; the
//     LayerNorm-folding GEMM that consumes x merges the 2·N/256 partials of a row in its epilogue. No statistics kernel, no
//     second pass over x, deterministic (every partial has exactly one writer).
//   * G2_LNFOLD: LayerNorm applied algebraically. A is the RAW residual stream, W' = W·diag(gamma), and
//       LN(x) W^T + b = rstd_r · acc − rstd_r·mean_r · svec_n + cvec_n,   svec = W'·1,  cvec = W·beta + b,
//     with (rstd_r, −mean_r·rstd_r) read per row from `ln_stats` (row_moments kernels, layernorm.cuh). No
//     normalised copy of x is ever written: LayerNorm costs one float2 per row.
#pragma once

#include "common.cuh"

namespace arp {

enum GemmAct : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_RELU = 2 };
enum G2Mode : int { G2_STORE = 0, G2_REDUCE = 1, G2_LNFOLD = 2, G2_RESID_STATS = 3 };

struct GemmArgs {
  int M, N, K;
  void* out;               // [M, ldo] op_t or fp32 (addressed through tmap_out; kept for diagnostics)
  int ldo;                 // elements
  const float* bias;       // [N] or nullptr
  const float* rowtab;     // fp32 [period, N] added to row (row % period), or nullptr (pos-emb + cls)
  int period;
  int reverse;             // walk the tiles from the last row block to the first (snake order across kernels, so a
                           // kernel starts on the rows its predecessor wrote last — still in L2)
  const float2* ln_stats;  // G2_LNFOLD: ln_nparts == 1: [M] (rstd, -mean*rstd) of the rows of A (row_moments kernels);
                           //            ln_nparts  > 1: [M, ln_nparts] partial (mean, M2), each over 128 columns (G2_RESID_STATS)
  int ln_nparts;
  const float* svec;       // G2_LNFOLD: [N]
  const float* cvec;       // G2_LNFOLD: [N]
  const op_t* resid;       // G2_RESID_STATS: the 16-bit residual stream [M, ldr] (aliases out: updated in place)
  int ldr;
  float2* stats_out;       // G2_RESID_STATS: [M, 2*N/256] partial (mean, M2) of the updated rows, one per 128 columns
  float eps;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;  // 64 x 16-bit = one 128-byte swizzle atom
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB

// QuickGELU x*sigmoid(1.702x) with ONE MUFU op: sigmoid(z) = 0.5 + 0.5*tanh(z/2)  ->  h + h*tanh(0.851x), h = x/2.
// tanh.approx.f32 has ~2^-11 relative error, at the rounding of the 16-bit store that follows.
__device__ __forceinline__ float quick_gelu(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

constexpr int G2_THREADS = 320;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_STAGE_UNIT = 32 * 128;  // one staging buffer: 32 rows x 128 B
constexpr int G2_STAGING_BYTES = G2_EPI_WARPS * G2_STAGE_UNIT;

template <int CG>
struct G2Cfg {
  static constexpr int B_ROWS = GEMM_BN / CG;                       // W rows staged per CTA
  static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_ROWS * GEMM_BK * 2;
  static constexpr int STAGES = CG == 1 ? 3 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + G2_STAGING_BYTES + 1024 + 256;
};

template <typename OutT, int ACT, int CG, int MODE>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const GemmArgs args) {
  using Cfg = G2Cfg<CG>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + G2_STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  static_assert(MODE != G2_RESID_STATS || sizeof(OutT) == 2, "G2_RESID_STATS updates the 16-bit residual stream");

  // warp index through a shuffle: provably warp-uniform, so the MMA / TMA issue paths keep their descriptors in
  // uniform registers (back-to-back UTCHMMA instead of an ELECT / R2UR.BROADCAST loop before every instruction)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  const int num_m = (args.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int num_n = args.N / GEMM_BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = args.K / GEMM_BK;
  // Iteration -> tile: tiles go round-robin over the clusters in n-inner order, so the clusters that share an A row block
  // run in the same wave and A comes out of L2 for all but the first of them.
  const int my_tiles = cluster_id < num_tiles ? (num_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  auto tile_coords = [&](int it, int& m_blk, int& n_blk) {
    const int tile = cluster_id + it * num_clusters;
    const int tile_o = args.reverse ? num_tiles - 1 - tile : tile;
    m_blk = tile_o / num_n;
    n_blk = tile_o % num_n;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], G2_EPI_WARPS * CG);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 1) tmem_alloc<512>(tmem_slot); else tmem_alloc_cg2<512>(tmem_slot);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of W) =====================
    // One elected thread runs the whole loop (no per-iteration ELECT / __syncwarp; the other lanes park at the final
    // barrier and cost no issue slots).
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        int m_blk, n_blk;
        tile_coords(it, m_blk, n_blk);
        const int a_row = (m_blk * CG + rank) * GEMM_BM;
        const int b_row = n_blk * GEMM_BN + rank * Cfg::B_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
          if (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * GEMM_BK, a_row);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * GEMM_BK, b_row);
          } else {
            // both CTAs' bytes are credited to the leader's barrier, which the leader arms for 2x the bytes
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t leader_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_cg2(sa, &tmap_a, leader_bar, kb * GEMM_BK, a_row);
            tma_load_2d_cg2(sb, &tmap_b, leader_bar, kb * GEMM_BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of the pair only) =====================
    if (rank == 0 && elect_one()) {   // a single thread waits, issues and commits
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, GEMM_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            if (CG == 1) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if (CG == 1) {
            umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
          } else {
            umma_commit_cg2(&empty_bar[stage], 0b11);
            if (kb == num_kb - 1) umma_commit_cg2(&tfull_bar[acc], 0b11);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may touch
    const int half = ew >> 2;       // accumulator column half
    uint8_t* sbuf = staging + ew * G2_STAGE_UNIT;
    constexpr int UNIT_COLS = sizeof(OutT) == 4 ? 32 : 64;   // 128 B per row
    constexpr int UNITS = 128 / UNIT_COLS;
    const uint32_t tempty_leader0 = CG == 2 ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0;
    const uint32_t tempty_leader1 = CG == 2 ? mapa_u32(smem_u32(&tempty_bar[1]), 0) : 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int sw = lane & 7;
    for (int it = 0; it < my_tiles; ++it) {
      int m_blk, n_blk;
      tile_coords(it, m_blk, n_blk);
      const int row0 = (m_blk * CG + rank) * GEMM_BM + quarter * 32;
      const int row = row0 + lane;
      // G2_RESID_STATS: the warp's 32 x 128 residual values of the tile (two 32-row x 128-byte units), requested before the
      // accumulator is waited for. Coalesced: load i of a unit covers rows 4i..4i+3, lane -> (row 4i + lane/8, chunk lane%8).
      uint4 xo[MODE == G2_RESID_STATS ? 16 : 1];
      float st_n = 0.f, st_mean = 0.f, st_m2 = 0.f;   // count / mean / centred 2nd moment of this lane's row over the tile's columns
      if (MODE == G2_RESID_STATS) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int r = row0 + 4 * (c & 7) + (lane >> 3);
          const op_t* xp = args.resid + static_cast<size_t>(r) * args.ldr + n_blk * GEMM_BN + half * 128 + (c >> 3) * 64 + (lane & 7) * 8;
          xo[c] = r < args.M ? *reinterpret_cast<const uint4*>(xp) : make_uint4(0u, 0u, 0u, 0u);
        }
        // the NEXT tile's residual values are pulled into L2 now, a whole tile period ahead: under a saturated HBM the
        // loaded DRAM latency is several microseconds, more than this warp can cover between two of its own tiles
        if (lane == 0 && it + 1 < my_tiles) {
          int m2, n2;
          tile_coords(it + 1, m2, n2);
          const int r2 = (m2 * CG + rank) * GEMM_BM + quarter * 32, c2 = n2 * GEMM_BN + half * 128;
          tma_prefetch_l2_2d(&tmap_out, c2, r2);
          tma_prefetch_l2_2d(&tmap_out, c2 + 64, r2);
        }
      }
      float ln_rstd = 1.f, ln_rm = 0.f;        // G2_LNFOLD: the row's 1/std and -mean/std (in flight under the MMAs)
      if (MODE == G2_LNFOLD && row < args.M) {
        if (args.ln_nparts == 1) {
          const float2 p = __ldg(args.ln_stats + row);
          ln_rstd = p.x;
          ln_rm = p.y;
        } else {
          // merge the row's partial (mean, M2) — equal counts of 128 columns each (Chan et al.). All loads are issued
          // before the first use: a dependent load-add loop would serialise one L2 round trip per partial.
          const float2* pp = args.ln_stats + static_cast<size_t>(row) * args.ln_nparts;
          float mean = 0.f, m2 = 0.f;
          if (args.ln_nparts == 6) {            // width 768: 48 contiguous, 16-byte aligned bytes
            const float4 a = __ldg(reinterpret_cast<const float4*>(pp));
            const float4 b = __ldg(reinterpret_cast<const float4*>(pp) + 1);
            const float4 c = __ldg(reinterpret_cast<const float4*>(pp) + 2);
            mean = ((a.x + a.z) + (b.x + b.z) + (c.x + c.z)) * (1.0f / 6.0f);
            const float d0 = a.x - mean, d1 = a.z - mean, d2 = b.x - mean, d3 = b.z - mean, d4 = c.x - mean, d5 = c.z - mean;
            m2 = ((a.y + a.w) + (b.y + b.w) + (c.y + c.w)) +
                 128.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) + (d4 * d4 + d5 * d5));
          } else {
            for (int i = 0; i < args.ln_nparts; ++i) mean += __ldg(pp + i).x;
            mean *= 1.0f / static_cast<float>(args.ln_nparts);
            for (int i = 0; i < args.ln_nparts; ++i) {
              const float2 p = __ldg(pp + i);
              const float d = p.x - mean;
              m2 += fmaf(128.0f * d, d, p.y);
            }
          }
          ln_rstd = rsqrtf(m2 / (128.0f * static_cast<float>(args.ln_nparts)) + args.eps);
          ln_rm = -mean * ln_rstd;
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * GEMM_BN + half * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
      const float* tab_row =
          args.rowtab ? args.rowtab + static_cast<size_t>(row % args.period) * args.N : nullptr;
#pragma unroll(MODE == G2_RESID_STATS ? UNITS : 1)
      for (int u = 0; u < UNITS; ++u) {
        const int n0 = n_blk * GEMM_BN + half * 128 + u * UNIT_COLS;
        if (MODE == G2_RESID_STATS) {
          // 64-column unit in two 32-column halves (register budget: 168 with ten warps per CTA):
          //   x_new = fl16(x_old + acc + bias); statistics of the ROUNDED values — what is stored and what the next GEMM
          //   multiplies — two-pass per half (they are all in registers), Chan's merge into the running (n, mean, M2)
          if (lane == 0) tma_store_wait_read<0>();     // the staging buffer was last read by the previous unit's store
          __syncwarp();
          // coalesced layout -> row-per-lane through the staging buffer (same 128B swizzle as the output: conflict-free)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);
            *reinterpret_cast<uint4*>(sbuf + r * 128 + (((lane & 7) ^ (r & 7)) << 4)) = xo[u * 8 + i];
          }
          __syncwarp();
          uint4* srow = reinterpret_cast<uint4*>(sbuf + lane * 128);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + u * 64 + hh * 32, r);
            tmem_ld_wait();
            if (u == UNITS - 1 && hh == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 1) mbar_arrive_relaxed(&tempty_bar[acc]);
                else mbar_arrive_cluster_relaxed(acc ? tempty_leader1 : tempty_leader0);
              }
            }
            float y[32];
            float s1 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 q = srow[(hh * 4 + c) ^ sw];
              const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + hh * 32 + 8 * c));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + hh * 32 + 8 * c + 4));
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const op2_t xh = *reinterpret_cast<const op2_t*>(&w4[e]);
                const op2_t yh = floats_to_op2(__uint_as_float(r[8 * c + 2 * e]) + bb[2 * e] + op_to_float(xh.x),
                                               __uint_as_float(r[8 * c + 2 * e + 1]) + bb[2 * e + 1] + op_to_float(xh.y));
                pk[e] = *reinterpret_cast<const uint32_t*>(&yh);
                y[8 * c + 2 * e] = op_to_float(yh.x);
                y[8 * c + 2 * e + 1] = op_to_float(yh.y);
                s1 += y[8 * c + 2 * e] + y[8 * c + 2 * e + 1];
              }
              srow[(hh * 4 + c) ^ sw] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            const float mu = s1 * (1.0f / 32.0f);
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = y[j] - mu;
              m2 = fmaf(d, d, m2);
            }
            const float nt = st_n + 32.0f;
            const float d = mu - st_mean, f = __fdividef(32.0f, nt);
            st_mean = fmaf(d, f, st_mean);
            st_m2 += fmaf(d * d, st_n * f, m2);
            st_n = nt;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, sbuf, n0, row0);
            tma_store_commit();
          }
          continue;
        }
        float v[UNIT_COLS];
        {
          uint32_t r[32];
          tmem_ld_32x32(taddr + u * UNIT_COLS, r);
          if (UNIT_COLS == 64) {
            uint32_t r2[32];
            tmem_ld_32x32(taddr + u * UNIT_COLS + 32, r2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[(UNIT_COLS == 64 ? 32 : 0) + j] = __uint_as_float(r2[j]);
          } else {
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (u == UNITS - 1) {
          // all of this warp's accumulator columns are in registers: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 1) mbar_arrive_relaxed(&tempty_bar[acc]);
            else mbar_arrive_cluster_relaxed(acc ? tempty_leader1 : tempty_leader0);
          }
        }
        if (MODE == G2_LNFOLD) {
#pragma unroll
          for (int j = 0; j < UNIT_COLS; j += 4) {
            const float4 sv = __ldg(reinterpret_cast<const float4*>(args.svec + n0 + j));
            const float4 cv = __ldg(reinterpret_cast<const float4*>(args.cvec + n0 + j));
            v[j] = fmaf(ln_rstd, v[j], fmaf(ln_rm, sv.x, cv.x));
            v[j + 1] = fmaf(ln_rstd, v[j + 1], fmaf(ln_rm, sv.y, cv.y));
            v[j + 2] = fmaf(ln_rstd, v[j + 2], fmaf(ln_rm, sv.z, cv.z));
            v[j + 3] = fmaf(ln_rstd, v[j + 3], fmaf(ln_rm, sv.w, cv.w));
          }
        } else if (args.bias) {
#pragma unroll
          for (int j = 0; j < UNIT_COLS; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (tab_row) {
#pragma unroll
          for (int j = 0; j < UNIT_COLS; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(tab_row + n0 + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (ACT == ACT_QUICKGELU) {
#pragma unroll
          for (int j = 0; j < UNIT_COLS; ++j) v[j] = quick_gelu(v[j]);
        } else if (ACT == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < UNIT_COLS; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        // the staging buffer was last read by the TMA store of the previous unit
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        uint4* srow = reinterpret_cast<uint4*>(sbuf + lane * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 q;
          if (sizeof(OutT) == 4) {
            q = make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]),
                           __float_as_uint(v[4 * c + 3]));
          } else {
            q = make_uint4(pack_op(v[8 * c], v[8 * c + 1]), pack_op(v[8 * c + 2], v[8 * c + 3]),
                           pack_op(v[8 * c + 4], v[8 * c + 5]), pack_op(v[8 * c + 6], v[8 * c + 7]));
          }
          srow[c ^ sw] = q;   // 128B swizzle: 16-byte chunk index XOR (row & 7)
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (MODE == G2_REDUCE) tma_reduce_add_2d(&tmap_out, sbuf, n0, row0);
          else tma_store_2d(&tmap_out, sbuf, n0, row0);
          tma_store_commit();
        }
      }
      if (MODE == G2_RESID_STATS && row < args.M)
        args.stats_out[static_cast<size_t>(row) * (2 * num_n) + n_blk * 2 + half] = make_float2(st_mean, st_m2);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1) tmem_dealloc<512>(tmem_base); else tmem_dealloc_cg2<512>(tmem_base);
  }
}

}  // namespace arp

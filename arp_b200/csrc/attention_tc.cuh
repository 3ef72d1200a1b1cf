// Attention core on the 5th-gen tensor cores:  O = softmax(Q K^T / 8) V  per (frame, head).
// Reference semantics: nn.MultiheadAttention inside openai/CLIP's ResidualAttentionBlock (no mask for the
// vision tower), reached from arp_dt/label_reward.py:141.
//
// Persistent CTAs loop over (frame, head) items. For one item (L = 197 tokens, padded to 208 keys / 2x128 queries):
//   TMA      Q (two 128x64 tiles), K and V (208x64) of the head -> 128B-swizzled smem, double buffered across items
//   MMA #1   S[128 x 208] = Q K^T     tcgen05.mma kind::f16, A and B from smem (K-major), fp32 accumulators in TMEM
//   softmax  one thread per query row: tcgen05.ld the row, max / exp2 in registers — no cross-lane traffic —
//            and write P back INTO the same TMEM columns as packed 16-bit pairs (tcgen05.st)
//   MMA #2   O[128 x 80] = P [V | 1]  A operand straight from TMEM, B = V from smem as an MN-major operand plus an
//            all-ones N block: columns 64..79 of O are the row sums of exactly the weights the tensor core used
//   epilogue tcgen05.ld O, scale by 1/rowsum, 16-bit, staged in smem in the TMA's swizzled layout, ONE bulk tensor
//            store per warp (rows past the frame's last token are out of bounds in the 3-D map and are not written)
// TMEM map: query tile t owns S fp32 [NK t, NK t + NK) -> P pairs overwrite [NK t, NK t + NK/2); ONE O accumulator
// fp32 [NK QT, NK QT + 80) is shared by the two tiles, whose P V / drain windows are half a period apart (o_drained
// barriers). Because O does not overlay S, the thread that issues a tile's P V issues the tile's NEXT Q K^T right behind
// it (tcgen05.mma executes in issue order: the new S cannot overtake the P it overwrites) — the next S is computed
// while the warps drain and store O instead of after it.
// Query tile 1 (rows 128..196) is SPREAD over the four TMEM lane quarters — five 16-row TMA boxes put rows 128 + 16q.. in
// lanes 0..15 of quarter q and rows 176..196 in lanes 0..20 of quarter 3 — and its softmax reads S with the 16-lane
// tcgen05.ld.16x256b shape, which hands every row to four threads (row maximum = 4-thread shuffle; P goes back with the
// matching 16x128b store): a quarter's pass is 104 MUFU instructions per thread instead of 208, and the busiest SMSP
// carries 364 per item instead of 416 (197 rows are 24.6 eight-row units; 7 on one SMSP, 6 on the others is the optimum).
// Warp roles (448 threads): 0 = TMA producer, 1 = TMEM owner (alloc / dealloc only), 2..9 = softmax (warps 2-5 own query
// tile 0, warps 6-9 query tile 1; the quarter-3 warp of a tile also issues its MMAs), 10..13 = epilogue, one per TMEM
// lane quarter (a warp may only touch TMEM lanes 32*(warp%4)..+31); the quarter-3 epilogue warp also does the softmax
// of tile 1's last five rows (lanes 16..20 of quarter 3), in tile 1's turn, beside the softmax warp of that quarter.
//
// Scheduling (measured: ARP_ATTN_TRACE timeline with an observer warp, tools/attn_trace.py; ncu warp-state samples;
// microbenchmarks tools/micro/tmem_contention.cu, smsp_interference.cu; numbers in profiles/r02_attn_*):
//   * every softmax warp is a serial chain per item — exp2 pass, P V, next S, row maximum — and until that chain is
//     shorter than the other slot's exp2 pass its length, not the MUFU time, is the slot's period. Three things cut it
//     from ~3800 to ~2100 clk: the next S right behind P V (above), the epilogue on warps of its own, and mbarrier
//     hand-offs instead of MEMBAR.SC + ATOMS counters / LDS polling. Now the period is two exp2 passes per SMSP;
//   * a fixed MMA issuer per slot, the quarter-3 warp: tile 1's has no live row (rows 224..255), tile 0's shares its SMSP
//     with that idle warp and finishes its pass first. It parks on the slot's p_done barrier (4 warp arrivals);
//   * everything the MMA issue needs is derived from a shuffled (provably warp-uniform) warp index, so the
//     descriptors live in uniform registers: back-to-back UTCHMMA instead of an ELECT / R2UR.BROADCAST loop;
//   * the exp2 pass is taken in turns per SMSP (xu_turn barriers): the two slots' MUFU-bound passes never overlap, which
//     keeps the slots half a period apart (free-running: +2 % now, +12 % before the chain was shortened);
//   * the tensor pipe is NOT slowed by the other slot's tcgen05.ld / st traffic (Q K^T 655 vs 670 clk), but a warp's
//     exp2 pass is slowed by whatever its SMSP neighbours issue (a tcgen05.ld stream +27 %, the epilogue +13 %, an LDS
//     polling loop +8 %, a parked try_wait +2 %);
//   * thirty-two lanes storing 16 bytes to 32 different lines, eight times per warp, held the warps in the LSU: the
//     output goes through smem and the TMA instead;
//   * a refill of the item's 84 KB takes ~3000 clk with every SM loading: Q / K (dead after the item's S) and V (dead
//     after its P V) are released and refilled separately, so the S issued behind P V never waits for its operands;
//   * the exp2 pass uses packed fp32x2 FMA and an integer truncate-and-merge for the 16-bit pairs (F2FP would
//     issue on the XU pipe, the one MUFU.EX2 needs).
#pragma once

#include "common.cuh"

namespace arp {

constexpr int ATC_THREADS = 448;   // 14 warps: TMA, TMEM owner, 8 softmax, 4 epilogue
constexpr int ATC_DH = 64;

template <int L>
struct AtcCfg {
  static constexpr int NK = (L + 15) / 16 * 16;        // padded keys: 208 (L=197), 64 (L=50)
  static constexpr int QT = (L + 127) / 128;           // query tiles: 2 / 1
  static constexpr int Q_BYTES = 128 * ATC_DH * 2;     // 16 KB per query tile
  static constexpr int KV_BYTES = NK * ATC_DH * 2;     // 26 KB
  static constexpr int KV_PAD = (KV_BYTES + 1023) / 1024 * 1024;
  static constexpr int BUF_BYTES = QT * Q_BYTES + 2 * KV_PAD;
  // the all-ones N block of the P V operand: ONE 16-key step (16 rows x 128 B); every step's descriptor points its
  // second N block (LBO) at this same tile
  static constexpr int ONES_BYTES = 16 * 128;
  static constexpr int STAGE_BYTES = QT * 128 * ATC_DH * 2;   // O rows on their way to the bulk store: 4 KB per warp
  static constexpr int SMEM_BYTES = 2 * BUF_BYTES + ONES_BYTES + STAGE_BYTES + 1024 + 256;
  static constexpr int O_COL = QT * NK;                // the shared O accumulator's first TMEM column
  static constexpr int O_N = ATC_DH + 16;              // P V output width: 64 head dims + 16 copies of the row sum
  // Query tile 1 (rows 128..L-1) is SPREAD over the four TMEM lane quarters — 16 rows in lanes 0..15 of quarters 0..2, the
  // rest (21 for L = 197) in lanes 0..20 of quarter 3 — and its softmax reads S with the 16-lane tcgen05.ld shape, which
  // hands every row to FOUR threads: a quarter's pass costs 104 MUFU instructions (156 in quarter 3) instead of 208, and
  // no SMSP carries two full exp2 passes per item any more (416 -> 364 MUFU instructions on the busiest one).
  static constexpr bool SPREAD = QT == 2;
  static constexpr int Q1_LOADS = 5;                   // 16-row boxes of tile 1: quarters 0..2 one each, quarter 3 two
  static constexpr int QK_TX_BYTES = SPREAD ? Q_BYTES + Q1_LOADS * 16 * 128 + KV_BYTES : QT * Q_BYTES + KV_BYTES;
  static_assert(!SPREAD || (L - 176 > 16 && L - 176 <= 24 && NK % 64 == 16), "tile-1 spread is laid out for 193..200 tokens");
  static_assert(O_COL + O_N <= 512, "TMEM columns");
  static_assert(KV_PAD >= (NK / 16 - 1) * 2048, "the ones tile must sit past every step's V rows (positive LBO)");
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
// lbo_bytes: distance to the next 64-element block along MN (only read when the instruction's N exceeds 64).
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Two fp32 -> packed bf16 pair WITHOUT F2FP (which issues on the XU pipe, the same pipe as MUFU.EX2): adding 0x8000
// to the bit pattern rounds the magnitude to nearest (ties away from zero; finite inputs), then the two high halves
// are merged with a shift and a LOP3 on the ALU pipe.
__device__ __forceinline__ uint32_t pack_bf16_int(float lo, float hi) {
  const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
  return (a >> 16) | (b & 0xffff0000u);
}

// bf16 pair by TRUNCATION (two ALU ops, no rounding adds): the weights P are normalised by the tensor core's own sum of
// the truncated values (ones block), so the one-sided error cancels in the mean and what remains has the spread of
// round-to-nearest.
//
// fp16 operands: the same shortcut exists if the exponential is produced PRE-SCALED. For an fp32 value v·2^-112 the bit
// pattern shifted right by 13 is exactly the fp16 pattern of v truncated (5 exponent + 10 mantissa bits line up once the
// bias difference 127-15 is taken out; fp32 denormals map onto fp16 denormals). The softmax therefore computes
// p' = exp2(s·scale − max·scale + 15 − 112): the largest weight of a row is 2^15 instead of 1 (fp16 tops out at 65504),
// the smallest one kept is 2^-29 of it (ex2.approx.ftz flushes below 2^-126), and the common factor 2^15 cancels in
// O / rowsum because the row sum comes from the same operand. No F2FP on the XU pipe, which MUFU.EX2 needs.
constexpr float kPExpOffset = ARP_OP_FP16 ? 97.0f : 0.0f;   // subtracted from the exp2 argument (112 - 15)
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
#if ARP_OP_FP16
  return (__float_as_uint(lo) >> 13) | ((__float_as_uint(hi) << 3) & 0xffff0000u);
#else
  return (__float_as_uint(lo) >> 16) | (__float_as_uint(hi) & 0xffff0000u);
#endif
}
constexpr uint32_t kOnes2 = ARP_OP_FP16 ? 0x3c003c00u : 0x3f803f80u;   // two 1.0 in the operand format

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

// 16-lane shapes (lane / column of every register verified on the part: tools/micro/tmem_shapes.cu). With lane base
// L0 in the address, thread t of tcgen05.ld.16x256b.xN holds, for step i < N: lane L0 + t/4, columns 8i + 2(t%4), +1 in
// registers 4i, 4i+1 and lane L0 + 8 + t/4, same columns, in 4i+2, 4i+3. tcgen05.st.16x128b.xN writes register 2i to
// lane L0 + t/4, column 4i + t%4 and 2i+1 to lane L0 + 8 + t/4: exactly where the packed pair of the loaded columns goes.
__device__ __forceinline__ void tmem_ld_16x256_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x128_x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x128_x2(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Test-only build (-DARP_ATTN_JITTER, tests/test_gpu_edges.py::test_attention_protocol_survives_jitter): every warp
// sleeps a pseudo-random 0..8 us at each hand-off point, so that a protocol which only holds for the usual relative
// speeds of the warps (the epilogue outrunning the softmax, a barrier never more than a phase behind its waiter) fails
// in the test instead of under a profiler or a sanitizer.
#ifdef ARP_ATTN_JITTER
__device__ __forceinline__ void atc_jitter(uint32_t salt) {
  uint32_t h = (static_cast<uint32_t>(clock64()) ^ (salt * 2654435761u) ^ (threadIdx.x >> 5) * 40503u ^ blockIdx.x * 9176u);
  h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
  if (h & 1) __nanosleep((h >> 1) & 8191);
}
#define ATC_JITTER(salt) atc_jitter(salt)
#else
#define ATC_JITTER(salt)
#endif

#ifdef ARP_ATTN_TRACE
// dev-only timeline of block 0: clock64 stamps kept in shared memory (a plain st.shared per event), dumped at exit.
// events: 0 S_issue 1 PV_issue 2 S_ready 3 pass1_done 4 turn_acquired 5 P_arrive 6 O_ready 7 drained (quarter-0
// warps), then per lane quarter q: 8+q pass2_end, 12+q O_seen, 16+q stored, 20+q P stores retired
constexpr int ATC_TR_ITEMS = 10, ATC_TR_EVENTS = 36;
__device__ long long g_attn_trace[2 * ATC_TR_ITEMS * ATC_TR_EVENTS];
#define ATC_TRACE(ev, slot, item)                                                                              \
  do {                                                                                                         \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (item) < ATC_TR_ITEMS)                                   \
      atc_tr[((slot) * ATC_TR_ITEMS + (item)) * ATC_TR_EVENTS + (ev)] = clock64();                             \
  } while (0)
// the same from whichever lane executes it (inside an elect_one() block) — events 24.. : S issue begin / first MMA out /
// last MMA out / committed, P V issue begin / last MMA out / committed; 31, 32: an idle observer warp sees S / O ready
#define ATC_TRACE_T(ev, slot, item)                                                                            \
  do {                                                                                                         \
    if (blockIdx.x == 0 && (item) < ATC_TR_ITEMS)                                                              \
      atc_tr[((slot) * ATC_TR_ITEMS + (item)) * ATC_TR_EVENTS + (ev)] = clock64();                             \
  } while (0)
#else
#define ATC_TRACE(ev, slot, item)
#define ATC_TRACE_T(ev, slot, item)
#endif

// qkv: 16-bit [rows, 3*width] (tensor maps: box 64x128 for Q, 64xNK for K/V); out: 16-bit [B*L, width] through tmap_o,
// a 3-D view [B, L, width] with a 1 x 32 x 64 box
//
// Each query tile of an item is an independent "job" with its own S / P TMEM columns and its own S / O barriers; the
// tiles share the O accumulator, the MUFU pipe (xu_turn) and the item's K / V.
template <int L>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_q16,
                    const __grid_constant__ CUtensorMap tmap_o16, int n_frames, int heads, int width, float scale_log2e,
                    int reverse) {
  using C = AtcCfg<L>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones = smem + 2 * C::BUF_BYTES;     // [16 rows x 128 B] of 1.0: N block 1 of every P V step
  uint8_t* stage = ones + C::ONES_BYTES;       // [QT][4 quarters][32 rows x 128 B] O rows for the bulk store
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + C::STAGE_BYTES);
  // Two ring states per smem buffer. Q and K of an item are dead as soon as the item's Q K^T have run — almost a full
  // period before its V is (after both slots' P V) — and the NEXT item's Q K^T is issued right behind this item's P V,
  // so Q / K are refilled early, on their own barrier pair; a 84 KB refill takes ~3000 clk when every SM is loading.
  uint64_t* qk_full = bars;          // [2] TMA -> Q K^T issuer: the item's Q tiles and K have landed
  uint64_t* qk_empty = bars + 2;     // [2] tensor core (the item's Q K^T of every slot retired) -> TMA
  uint64_t* v_full = bars + 4;       // [2] TMA -> P V issuer: the item's V has landed
  uint64_t* v_empty = bars + 6;      // [2] tensor core (the item's P V of every slot retired) -> TMA
  uint64_t* s_full = bars + 8;       // [2 slots] tensor core -> softmax warps: S ready
  uint64_t* o_full = bars + 10;      // [2 slots] tensor core -> the slot's warps: O ready
  // [2 slots] the four epilogue warps have drained the slot's O -> the OTHER slot's P V issuer. One barrier per slot: a
  // parity wait is only sound if the waiter is never more than one phase ahead of the barrier, which holds when each
  // issuer walks the other slot's barrier phase by phase — and does not on a single barrier whose phases alternate
  // between two waiters (compute-sanitizer's timing found that one).
  uint64_t* o_drained = bars + 12;   // bars + 12, + 13
  // exp2 turn per TMEM lane quarter (= per SMSP): the two softmax warps that share an SMSP take their MUFU-bound
  // pass strictly in (item, slot) order, never both at once. Turn n of a quarter may start when phase n - 1 of its
  // barrier has completed (the warp that finished turn n - 1 arrives); a parked try_wait costs the SMSP next to nothing,
  // an LDS polling loop cost the neighbour's exp2 pass 8 %.
  uint64_t* xu_turn = bars + 14;     // [4 quarters]
  // P rows of a slot written (4 warp arrivals, 5 for the spread tile 1) -> the slot's MMA issuer: the quarter-3 warp of the
  // slot. Its SMSP carries the most exp2 work (364 MUFU instructions per item), so it is usually the LAST of the slot's
  // warps to arrive and finds the phase complete without parking; and a parked try_wait notices the last arrival sooner
  // than the two MEMBAR.SC + ATOMS of a "last one in issues" counter took.
  uint64_t* p_done = bars + 18;      // [2 slots]
  // spread layout: tile 0's quarter-3 warp has finished its exp2 pass (= tile 1's turn on that SMSP begins) -> the quarter-3
  // epilogue warp, which does the softmax of tile 1's last rows (the second 16-lane group of quarter 3)
  // (two barriers, used by alternate items: tile 0 may finish the NEXT item's pass before that warp has looked — its
  // pass after that needs the warp's own drain — so with one barrier the waiter could fall two phases behind, where a
  // parity wait blocks for good; the jitter test found it)
  uint64_t* rem_go = bars + 20;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  // warp index through a shuffle: provably warp-uniform for the compiler, so everything derived from it (slot,
  // TMEM addresses, descriptors) lives in uniform registers and an MMA issue is not an ELECT / R2UR.BROADCAST loop
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_items = n_frames * heads;
#ifdef ARP_ATTN_TRACE
  __shared__ long long atc_tr[2 * ATC_TR_ITEMS * ATC_TR_EVENTS];
  for (int i = threadIdx.x; i < 2 * ATC_TR_ITEMS * ATC_TR_EVENTS; i += ATC_THREADS) atc_tr[i] = 0;
#endif
  constexpr int SM_WARPS = 4 * C::QT;   // softmax warps that actually own rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    tma_prefetch_desc(&tmap_o);
    tma_prefetch_desc(&tmap_q16);
    tma_prefetch_desc(&tmap_o16);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], C::QT);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], C::QT);
      mbar_init(&s_full[i], 1);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&xu_turn[i], 1);
    mbar_init(&p_done[0], 4);
    mbar_init(&p_done[1], C::SPREAD ? 5 : 4);   // + the quarter-3 epilogue warp (rows 192.. of tile 1)
    mbar_init(&rem_go[0], 1);
    mbar_init(&rem_go[1], 1);
    mbar_init(&o_drained[0], 4);
    mbar_init(&o_drained[1], 4);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < C::ONES_BYTES / 16; i += ATC_THREADS)      // swizzle-invariant: every element is 1.0
    reinterpret_cast<uint4*>(ones)[i] = make_uint4(kOnes2, kOnes2, kOnes2, kOnes2);
  if (C::SPREAD) {
    // lanes 16..31 of quarters 0..2 of tile 1 are never loaded: keep their Q rows finite (zero) for the tensor core
    for (int i = threadIdx.x; i < 2 * 3 * 2048 / 16; i += ATC_THREADS) {
      const int b = i / (3 * 128), q = (i / 128) % 3, w = i % 128;
      reinterpret_cast<uint4*>(smem + b * C::BUF_BYTES + C::Q_BYTES + q * 4096 + 2048)[w] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async_smem();                                                // generic-proxy writes -> tensor-core reads
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // ---- MMA issue (one elected thread of a softmax warp) ----
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, C::NK);          // Q K^T: both operands K-major
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, C::O_N, 0, 1);    // P [V | 1]: A from TMEM, B MN-major, N = 80
  auto issue_s = [&](uint32_t sbuf, int t, [[maybe_unused]] int tr_item) {
    const uint64_t dk = umma_desc_kmajor_sw128(sbuf + C::QT * C::Q_BYTES);
    const uint64_t dq = umma_desc_kmajor_sw128(sbuf + t * C::Q_BYTES);
    ATC_TRACE_T(24, t, tr_item);
#pragma unroll
    for (int k = 0; k < ATC_DH / 16; ++k) umma_bf16_ss(tmem_base + t * C::NK, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
    umma_commit(&s_full[t]);
    umma_commit(&qk_empty[(sbuf - smem_u32(smem)) / C::BUF_BYTES]);   // Q / K of the buffer: free once every slot's S ran
    ATC_TRACE_T(27, t, tr_item);
  };
  auto issue_pv = [&](uint32_t sbuf, int t, [[maybe_unused]] int tr_item) {
    // N block 0 = 16 keys of the item's V tile, N block 1 (LBO away) = the all-ones tile: O[:, 64..79] = sum_k P[:, k],
    // the softmax denominator of exactly the 16-bit weights the tensor core used. A step advances V by 16 rows x 128 B
    // and shortens LBO by as much, so that block 1 stays on the one ones tile.
    const uint32_t v_addr = sbuf + C::QT * C::Q_BYTES + C::KV_PAD;
    const uint64_t dv = umma_desc_mnmajor_sw128(v_addr, smem_u32(ones) - v_addr);
    constexpr uint64_t kStep = (2048ull >> 4) - ((2048ull >> 4) << 16);   // start += 2048 B, LBO -= 2048 B
    ATC_TRACE_T(28, t, tr_item);
#pragma unroll
    for (int k = 0; k < C::NK / 16; ++k)   // 16 keys per MMA: A advances 8 packed columns
      umma_bf16_ts(tmem_base + C::O_COL, tmem_base + t * C::NK + k * 8, dv + k * kStep, idesc_o, k != 0);
    umma_commit(&o_full[t]);
    ATC_TRACE_T(30, t, tr_item);
  };

  // ---- query tile 1, spread over the lane quarters (AtcCfg::SPREAD): softmax of one 16-lane group ----
  // A thread serves rows t/4 ("A": registers 4i, 4i+1 of a load) and 8 + t/4 ("B": 4i+2, 4i+3) of the group, columns
  // 8i + 2(t%4), +1 of every 8-column step: a row's maximum is a 4-thread shuffle reduction, its P words go back with the
  // matching 16x128b store shape.
  const bool sp_c0 = 192 + 2 * (lane & 3) < L, sp_c1 = 193 + 2 * (lane & 3) < L;   // this thread's real columns of the last load
  auto spread_rowmax = [&](uint32_t tg, float& row_a, float& row_b) {
    auto max16 = [](const uint32_t (&r)[32], float& ma, float& mb) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ma = fmaxf(ma, fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])));
        mb = fmaxf(mb, fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
      }
    };
    uint32_t a[32], bq[32], tl[8];
    float a0 = -INFINITY, a1 = -INFINITY, b0m = -INFINITY, b1m = -INFINITY;
    tmem_ld_16x256_x8(tg, a);
    tmem_ld_16x256_x8(tg + 64, bq);
    tmem_ld_wait();
    max16(a, a0, b0m);
    max16(bq, a1, b1m);
    tmem_ld_16x256_x8(tg + 128, a);
    tmem_ld_16x256_x2(tg + 192, tl);
    tmem_ld_wait();
    max16(a, a0, b0m);
    if (sp_c0) { a1 = fmaxf(a1, __uint_as_float(tl[0])); b1m = fmaxf(b1m, __uint_as_float(tl[2])); }
    if (sp_c1) { a1 = fmaxf(a1, __uint_as_float(tl[1])); b1m = fmaxf(b1m, __uint_as_float(tl[3])); }
    float ma = fmaxf(a0, a1), mb = fmaxf(b0m, b1m);
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    row_a = ma;
    row_b = mb;
  };
  // p = exp2(s*scale - max*scale) of the group, P written in place; BOTH = rows A and B, otherwise row A only (row B's words
  // are written as zero). The next 64 columns are in flight while these are exponentiated.
  auto spread_exp = [&](uint32_t tg, float max_a, float max_b, bool both) {
    const float mo_a = max_a * scale_log2e + kPExpOffset, mo_b = max_b * scale_log2e + kPExpOffset;
    const uint64_t sc2 = f32x2_pack(scale_log2e, scale_log2e), na2 = f32x2_pack(-mo_a, -mo_a), nb2 = f32x2_pack(-mo_b, -mo_b);
    auto chunk16 = [&](const uint32_t (&src)[32], int blk) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float x0, x1;
        f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(src[4 * i]), __uint_as_float(src[4 * i + 1])), sc2, na2), x0, x1);
        pk[2 * i] = pack_bf16_trunc(ex2_approx(x0), ex2_approx(x1));
        if (both) {
          f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(src[4 * i + 2]), __uint_as_float(src[4 * i + 3])), sc2, nb2), x0, x1);
          pk[2 * i + 1] = pack_bf16_trunc(ex2_approx(x0), ex2_approx(x1));
        } else {
          pk[2 * i + 1] = 0u;
        }
      }
      tmem_st_16x128_x8(tg + blk * 32, pk);      // P words 32 blk .. + 31 <- S columns 64 blk .. + 63
    };
    uint32_t b0[32], b1[32], tl[8];
    tmem_ld_16x256_x8(tg, b0);
    tmem_ld_wait();
    tmem_ld_16x256_x8(tg + 64, b1);
    chunk16(b0, 0);
    tmem_ld_wait();
    tmem_ld_16x256_x8(tg + 128, b0);
    chunk16(b1, 1);
    tmem_ld_wait();
    tmem_ld_16x256_x2(tg + 192, tl);
    chunk16(b0, 2);
    tmem_ld_wait();
    uint32_t pk4[4] = {0u, 0u, 0u, 0u};
    const float pa0 = sp_c0 ? ex2_approx(fmaf(__uint_as_float(tl[0]), scale_log2e, -mo_a)) : 0.f;
    const float pa1 = sp_c1 ? ex2_approx(fmaf(__uint_as_float(tl[1]), scale_log2e, -mo_a)) : 0.f;
    pk4[0] = pack_bf16_trunc(pa0, pa1);
    if (both) {
      const float pb0 = sp_c0 ? ex2_approx(fmaf(__uint_as_float(tl[2]), scale_log2e, -mo_b)) : 0.f;
      const float pb1 = sp_c1 ? ex2_approx(fmaf(__uint_as_float(tl[3]), scale_log2e, -mo_b)) : 0.f;
      pk4[1] = pack_bf16_trunc(pb0, pb1);
    }
    tmem_st_16x128_x2(tg + 96, pk4);             // P words 96..103 <- S columns 192..207 (padding -> 0)
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int item_o = reverse ? n_items - 1 - item : item;   // snake order across kernels (L2 reuse)
      const int frame = item_o / heads, head = item_o - frame * heads;
      uint8_t* buf = smem + b * C::BUF_BYTES;
      const int row = frame * L;
      ATC_JITTER(8 + it);
      mbar_wait(&qk_empty[b], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&qk_full[b], C::QK_TX_BYTES);
        if (C::SPREAD) {
          tma_load_2d(buf, &tmap_q, &qk_full[b], head * ATC_DH, row);
          uint8_t* q1 = buf + C::Q_BYTES;
#pragma unroll
          for (int q = 0; q < 3; ++q)      // rows 128 + 16 q .. + 15 -> lanes 0..15 of quarter q
            tma_load_2d(q1 + q * 4096, &tmap_q16, &qk_full[b], head * ATC_DH, row + 128 + 16 * q);
          tma_load_2d(q1 + 3 * 4096, &tmap_q16, &qk_full[b], head * ATC_DH, row + 176);          // quarter 3: rows 176..207
          tma_load_2d(q1 + 3 * 4096 + 2048, &tmap_q16, &qk_full[b], head * ATC_DH, row + 192);
        } else {
#pragma unroll
          for (int t = 0; t < C::QT; ++t)
            tma_load_2d(buf + t * C::Q_BYTES, &tmap_q, &qk_full[b], head * ATC_DH, row + t * 128);
        }
        tma_load_2d(buf + C::QT * C::Q_BYTES, &tmap_kv, &qk_full[b], width + head * ATC_DH, row);
      }
      __syncwarp();
      mbar_wait(&v_empty[b], ph ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&v_full[b], C::KV_BYTES);
        tma_load_2d(buf + C::QT * C::Q_BYTES + C::KV_PAD, &tmap_kv, &v_full[b], 2 * width + head * ATC_DH, row);
      }
      __syncwarp();
    }
#ifdef ARP_ATTN_TRACE
  } else if (warp == 1) {
    // dev-only observer: an otherwise idle warp polls the S-ready / O-ready barriers and stamps when each phase completes
    if (lane == 0 && blockIdx.x == 0) {
      int cnt[4] = {0, 0, 0, 0}, total = 0;
      const int n_mine_o = n_items > 0 ? (n_items - 1) / static_cast<int>(gridDim.x) + 1 : 0;
      const int lim = n_mine_o < ATC_TR_ITEMS ? n_mine_o : ATC_TR_ITEMS;
      const long long t0 = clock64();
      while (total < 4 * lim && clock64() - t0 < 4000000) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          uint64_t* bar = b < 2 ? &s_full[b] : &o_full[b - 2];
          if (cnt[b] < lim && mbar_test_wait(bar, cnt[b] & 1)) {
            atc_tr[((b & 1) * ATC_TR_ITEMS + cnt[b]) * ATC_TR_EVENTS + (b < 2 ? 31 : 32)] = clock64();
            ++cnt[b];
            ++total;
          }
        }
      }
    }
#endif
  } else if (warp >= 2 && warp - 2 < SM_WARPS) {
    // ===================== softmax: tile 0 one thread per query row, tile 1 (spread) four threads per row =====================
    const int qt = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + qt * C::NK + lane_off;
    constexpr int NFULL = C::NK / 32;                    // full 32-column chunks
    constexpr int TAIL = C::NK % 32;                     // 16 or 0
    const int n_mine = n_items > static_cast<int>(blockIdx.x) ? (n_items - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    // a warp whose 32 rows are all padding (L = 50: rows 64..127 of the only tile) keeps the barrier protocol but does no
    // softmax: whatever sits in its P rows only reaches O rows that are never stored
    const bool spread = C::SPREAD && qt == 1;          // warp-uniform: the 16-lane path of query tile 1
    const bool warp_live = spread || qt * 128 + quarter * 32 < L;
    if (quarter == 0 && n_mine > 0) {      // the slot's first Q K^T
      mbar_wait(&qk_full[0], 0);
      tc_fence_after();
      if (elect_one()) issue_s(smem_u32(smem), qt, 0);
      ATC_TRACE(0, qt, 0);
      __syncwarp();
    }
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      ATC_JITTER(1 + it);
      mbar_wait(&s_full[qt], ph);
      tc_fence_after();
      if (quarter == 0) ATC_TRACE(2, qt, it);
      // Only the last (partial or padded) chunk can contain key columns >= L; the others need no masking.
      constexpr int NCLEAN = (L / 32 < NFULL) ? L / 32 : NFULL;   // full chunks whose 32 columns are all real keys
      // ---- pass 1: row maximum over the L real keys ----
      float m = -INFINITY;
      // spread path: row maxima of this quarter's 16-lane group; a thread serves rows t/4 ("A") and 8 + t/4 ("B"), a
      // quarter of their columns each
      float mA = -INFINITY, mB = -INFINITY;
      if (spread) {
        spread_rowmax(t_s, mA, mB);
      } else if (warp_live) {
        uint32_t a[32], bq[32], cq[32];
        // three 32-column loads in flight per wait: the pass is bound by tcgen05.ld round trips, not by its 104 FMNMX3
#pragma unroll 1
        for (int c = 0; c + 2 < NCLEAN; c += 3) {
          tmem_ld_32x32(t_s + c * 32, a);
          tmem_ld_32x32(t_s + c * 32 + 32, bq);
          tmem_ld_32x32(t_s + c * 32 + 64, cq);
          tmem_ld_wait();
          float m0 = m, m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(a[j]), __uint_as_float(a[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(bq[j]), __uint_as_float(bq[j + 1])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(cq[j]), __uint_as_float(cq[j + 1])));
          }
          m = fmaxf(m0, fmaxf(m1, m2));
        }
#pragma unroll
        for (int c = NCLEAN / 3 * 3; c < NCLEAN; ++c) {
          tmem_ld_32x32(t_s + c * 32, a);
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
      ATC_JITTER(2 + it);
      const float mo = m * scale_log2e + kPExpOffset;   // exp2 argument offset: row maximum (+ the fp16 pre-scale)
      if (quarter == 0) ATC_TRACE(3, qt, it);
      if (C::QT == 2) {     // wait for this warp's exp2 turn
        const int my_turn = 2 * static_cast<int>(it) + qt;
        if (my_turn > 0) mbar_wait(&xu_turn[quarter], (my_turn - 1) & 1);
      }
      // ---- pass 2: p = exp2(s*scale - max*scale); P (16-bit pairs) overwrites the S columns it came from.
      //      Two register buffers ping-pong: the load of the next chunk is in flight while this one is processed. ----
      if (quarter == 0) ATC_TRACE(4, qt, it);
      const uint64_t scale2 = f32x2_pack(scale_log2e, scale_log2e), nmo2 = f32x2_pack(-mo, -mo);
      auto soft_chunk = [&](const uint32_t (&src)[32], int c, bool masked) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float x0, x1;
          f32x2_unpack(f32x2_fma(f32x2_pack(__uint_as_float(src[j]), __uint_as_float(src[j + 1])), scale2, nmo2), x0, x1);
          float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          if (masked) {
            if (c * 32 + j >= L) p0 = 0.f;
            if (c * 32 + j + 1 >= L) p1 = 0.f;
          }
          pk[j >> 1] = pack_bf16_trunc(p0, p1);
        }
        tmem_st_32x16(t_s + c * 16, pk);
      };
      if (spread) {
        spread_exp(t_s, mA, mB, true);
      } else if (warp_live) {
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
            pk[j >> 1] = pack_bf16_trunc(p0, p1);
          }
          tmem_st_32x8(t_s + NFULL * 16, pk);
        }
      }
      if (C::QT == 2) {     // hand the exp2 turn to the other slot's warp of this quarter (earlier, mid-pass: 4-8 % slower)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&xu_turn[quarter]);
          if (C::SPREAD && qt == 0 && quarter == 3) mbar_arrive(&rem_go[it & 1]);
        }
      }
      ATC_TRACE(8 + quarter, qt, it);
      ATC_JITTER(3 + it);
      tmem_st_wait();
      ATC_TRACE(20 + quarter, qt, it);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_done[qt]);   // release: this warp's P rows are in TMEM
      if (quarter == 3) {
        // the slot's MMA issuer: P V as soon as all four warps' P rows are written and, right behind it, the slot's NEXT
        // Q K^T
        mbar_wait(&p_done[qt], ph);   // parked: a busy-polling issuer was no faster (137 vs 136 us)
        ATC_JITTER(4 + it);
        // the shared O accumulator must have been drained by its previous user (slots alternate, the turn protocol orders
        // them): slot 1 follows slot 0 of the same item, slot 0 follows the last slot of the previous item
        if (qt > 0) mbar_wait(&o_drained[qt - 1], ph);
        else if (it > 0) mbar_wait(&o_drained[C::QT - 1], ph ^ 1);
        mbar_wait(&v_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t nx = it + 1;
        const bool has_next = static_cast<int>(nx) < n_mine;
        if (elect_one()) {
          issue_pv(smem_u32(smem + (it & 1) * C::BUF_BYTES), qt, it);
          umma_commit(&v_empty[it & 1]);   // the item's V is free once BOTH slots' P V have retired (count = QT)
          if (has_next) {
            mbar_wait(&qk_full[nx & 1], (nx >> 1) & 1);
            issue_s(smem_u32(smem + (nx & 1) * C::BUF_BYTES), qt, nx);
          }
        }
        ATC_TRACE(1, qt, it);
        __syncwarp();
      }
      if (quarter == 0) ATC_TRACE(5, qt, it);
    }
  } else if (warp >= 2 + 8) {
    // ===================== epilogue warps: O / rowsum -> 16 bits -> smem -> bulk tensor store =====================
    // One warp per TMEM lane quarter serves BOTH slots (their O windows alternate in the shared accumulator). Taking the
    // drain, the scaling and the store out of the softmax warps shortens THEIR per-item chain to exp2 pass -> P V ->
    // next S -> row maximum, which is what sets the kernel's period.
    const int quarter = warp & 3;
    const uint32_t t_o = tmem_base + C::O_COL + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int item_o = reverse ? n_items - 1 - item : item;
      const int frame = item_o / heads, head = item_o - frame * heads;
      if (C::SPREAD && quarter == 3) {
        // Tile 1's rows 192.. sit in lanes 16..23 of quarter 3, a second 16-lane group that would double that quarter's
        // pass (it was the kernel's longest: 2480 against 1350 clk). This warp is idle now — tile 0's O comes ~1000 clk
        // later — so it takes the group: both warps feed the SMSP's MUFU pipe during tile 1's turn.
        const uint32_t tg2 = tmem_base + C::NK + (static_cast<uint32_t>(96 + 16) << 16);
        ATC_JITTER(9 + it);
        mbar_wait(&s_full[1], it & 1);
        mbar_wait(&rem_go[it & 1], (it >> 1) & 1);
        tc_fence_after();
        float mg, unused;
        spread_rowmax(tg2, mg, unused);
        spread_exp(tg2, mg, mg, false);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_done[1]);
      }
#pragma unroll
      for (int qt = 0; qt < C::QT; ++qt) {
        const bool spread = C::SPREAD && qt == 1;               // tile 1: 16 rows per quarter (21 in quarter 3)
        const bool warp_live = spread || qt * 128 + quarter * 32 < L;
        uint8_t* my_stage = stage + (qt * 4 + quarter) * 4096;
        ATC_JITTER(5 + it);
        mbar_wait(&o_full[qt], it & 1);
        tc_fence_after();
        ATC_TRACE(12 + quarter, qt, it);
        ATC_JITTER(6 + it);
        uint32_t o0[32], o1[32], osum = 0;
        if (warp_live) {
          tmem_ld_32x32(t_o, o0);
          tmem_ld_32x32(t_o + 32, o1);
          tmem_ld_32x1(t_o + ATC_DH, osum);       // a copy of the row sum (the ones block of the operand)
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&o_drained[qt]);            // O may be overwritten by the next P V (count = 4 warps)
          // the bulk store that last read this staging buffer (this slot, an item ago) is done; the other slot's may fly
          if (C::QT == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        }
        __syncwarp();
        ATC_JITTER(7 + it);
        if (warp_live) {
          // one 128-byte output row per thread in the TMA's 128B-swizzled layout — chunk c of row r sits at chunk
          // c ^ (r & 7): conflict-free 16-byte shared stores
          const float inv = rcp_approx(__uint_as_float(osum));
          const uint64_t inv2 = f32x2_pack(inv, inv);
          uint4* srow = reinterpret_cast<uint4*>(my_stage + lane * 128);
          const int sw = lane & 7;
          auto pack2 = [&](uint32_t lo, uint32_t hi) {
            float a, b;
            f32x2_unpack(f32x2_mul(f32x2_pack(__uint_as_float(lo), __uint_as_float(hi)), inv2), a, b);
            return pack_op(a, b);
          };
#pragma unroll
          for (int c = 0; c < 4; ++c)
            srow[c ^ sw] = make_uint4(pack2(o0[8 * c], o0[8 * c + 1]), pack2(o0[8 * c + 2], o0[8 * c + 3]),
                                      pack2(o0[8 * c + 4], o0[8 * c + 5]), pack2(o0[8 * c + 6], o0[8 * c + 7]));
#pragma unroll
          for (int c = 0; c < 4; ++c)
            srow[(4 + c) ^ sw] = make_uint4(pack2(o1[8 * c], o1[8 * c + 1]), pack2(o1[8 * c + 2], o1[8 * c + 3]),
                                            pack2(o1[8 * c + 4], o1[8 * c + 5]), pack2(o1[8 * c + 6], o1[8 * c + 7]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (spread) {      // lanes 0..15 -> rows 128 + 16 q ..; quarter 3: lanes 0..31 -> rows 176..207, clipped at L
              tma_store_3d(&tmap_o16, my_stage, head * ATC_DH, 128 + 16 * quarter, frame);
              if (quarter == 3) tma_store_3d(&tmap_o16, my_stage + 2048, head * ATC_DH, 192, frame);
            } else {
              tma_store_3d(&tmap_o, my_stage, head * ATC_DH, qt * 128 + quarter * 32, frame);
            }
            tma_store_commit();
          }
        }
        ATC_TRACE(16 + quarter, qt, it);
      }
    }
    if (lane == 0) tma_store_wait_read<0>();   // the staged rows must outlive the last bulk store's reads
  }

  tc_fence_before();
  __syncthreads();
#ifdef ARP_ATTN_TRACE
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < 2 * ATC_TR_ITEMS * ATC_TR_EVENTS; i += blockDim.x) g_attn_trace[i] = atc_tr[i];
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace arp

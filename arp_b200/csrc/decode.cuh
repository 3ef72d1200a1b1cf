// Fused frame decode (SURVEY.md §2 K1 / K1'): take the LAST stacked frame of each dataset row,
// optionally centre-crop, resize to 224x224, normalise, and emit either
//   (a) bf16 patch rows of the ViT patch-embed GEMM  A[b*tokens + 1 + py*G + px, c*P*P + ph*P + pw]
//       (token row 0 of every frame is written as zeros: the class-token slot), or
//   (b) fp32 [T,3,224,224] — exactly what the reference hands to the model (test / parity hook).
//
// Mode PIL_BICUBIC restates label_reward.py:109-121 (ToPILImage -> Resize(224, BICUBIC) -> CenterCrop
// -> ToTensor -> Normalize; crop variant :92-102): Pillow's two-pass fixed-point resample —
// horizontal pass to a uint8 intermediate, vertical pass to uint8, 22-bit coefficients
// (SURVEY.md Appendix B). Bit-exact by construction: integer arithmetic, and the u8 -> float
// normalisation is a 3x256 table built on the host with the same IEEE fp32 ops torch uses.
// Mode BILINEAR restates clip_multiscale_adapter.py:121-133 (float bilinear, align_corners=False,
// no antialias, /255, normalise).
//
// One CTA per (16-output-row band, frame). The input rows a band needs are staged in shared
// memory with 16-byte coalesced loads, both passes run out of shared memory, and the finished
// band is written back with 16-byte stores (a patch-row band is one contiguous 21.5 KB span
// of the GEMM A matrix when P = 16).
#pragma once

#include "common.cuh"

namespace arp {

constexpr int DEC_OUT = 224;
constexpr int DEC_BAND = 16;
constexpr int DEC_THREADS = 256;
constexpr int DEC_MAX_K = 16;  // max taps per output pixel (ksize) supported by the tables

enum DecodeMode : int { DEC_PIL_BICUBIC = 0, DEC_BILINEAR = 1 };
enum DecodeOut : int { DEC_OUT_PATCH_BF16 = 0, DEC_OUT_CHW_F32 = 1 };

struct DecodeArgs {
  const uint8_t* ob;        // first byte of the first frame's LAST stacked image
  long long frame_stride;   // bytes between consecutive dataset rows (F*H*W*3)
  int in_w;                 // full image width in pixels (row pitch = in_w*3)
  int crop_top, crop_left;  // crop origin inside the image
  int src_h, src_w;         // size after crop = resize source size
  int T;
  int mode;                 // DecodeMode
  int out_kind;             // DecodeOut
  // bicubic tables (device): per output index, first tap and tap count, and int32 coefficients [224][ksize]
  const int* h_min; const int* h_cnt; const int* h_k; int h_ksize;
  const int* v_min; const int* v_cnt; const int* v_k; int v_ksize;
  const float* lut;         // [3][256] normalised value of each byte per channel (bicubic mode)
  float mean[3], stdv[3];
  // output
  void* out;
  int patch, grid, tokens;  // P, 224/P, grid*grid+1
  int max_rows;             // max input rows any band needs (sizes the smem carve-up)
};

__host__ __device__ inline int dec_out_bytes(int out_kind) {
  return DEC_BAND * DEC_OUT * 3 * (out_kind == DEC_OUT_PATCH_BF16 ? 2 : 4);
}
// Staged input rows and the horizontally resampled rows. Both carry slack that is never written: the fast path applies
// all KS taps of every output pixel (coefficients past a pixel's tap count are 0), so it reads a few bytes past the
// last staged row / up to KS-1 rows past the last resampled row; the slack keeps those reads off memory other threads write.
__host__ __device__ inline int dec_in_bytes(int src_w, int max_rows) { return (max_rows * src_w * 3 + 32 + 15) / 16 * 16; }
__host__ __device__ inline int dec_tmp_bytes(int max_rows) { return ((max_rows + 8) * DEC_OUT * 3 + 15) / 16 * 16; }
__host__ __device__ inline int dec_smem_bytes(int src_w, int max_rows, int h_ksize, int v_ksize, int out_kind) {
  int in_b = dec_in_bytes(src_w, max_rows);
  int tmp_b = dec_tmp_bytes(max_rows);
  int out_b = dec_out_bytes(out_kind);
  int tab_b = (DEC_OUT * 2 + DEC_OUT * h_ksize + DEC_BAND * 2 + DEC_BAND * v_ksize) * 4;
  return in_b + tmp_b + out_b + tab_b + 64;
}

__device__ __forceinline__ uint8_t clip8(int v) { return static_cast<uint8_t>(min(255, max(0, v))); }

// Both Pillow passes of one band with the tap count known at compile time (KS = the largest window that occurs: 4 when
// upscaling, 5 for 256 -> 224; the host re-packs Pillow's wider, zero-padded coefficient rows):
// one thread per output COLUMN. Its KS horizontal coefficients live in registers for every staged row, the tap loops are
// fully unrolled (entries past a pixel's tap count are 0 in the tables, so all KS taps are always applied; the few
// bytes read past a row's end are multiplied by 0), and the vertical coefficients of an output row are a shared-memory
// broadcast. Same integer arithmetic as the generic loops below — 3x fewer instructions (no per-tap coefficient
// load / bounds test / index division).
template <int KS, typename Put>
__device__ __forceinline__ void bicubic_band_fast(const DecodeArgs& a, const uint8_t* s_in, uint8_t* s_tmp, int nrows,
                                                  int row_bytes, int r_lo, int y0, const int* t_vmin, const int* t_vk,
                                                  Put put) {
  const int x = threadIdx.x;
  if (x < DEC_OUT) {
    int k[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) k[j] = __ldg(a.h_k + x * KS + j);
    const uint8_t* p0 = s_in + __ldg(a.h_min + x) * 3;
    uint8_t* q = s_tmp + x * 3;
#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
      const uint8_t* p = p0 + r * row_bytes;
      int acc0 = 1 << 21, acc1 = 1 << 21, acc2 = 1 << 21;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        acc0 += p[3 * j] * k[j]; acc1 += p[3 * j + 1] * k[j]; acc2 += p[3 * j + 2] * k[j];
      }
      q[0] = clip8(acc0 >> 22); q[1] = clip8(acc1 >> 22); q[2] = clip8(acc2 >> 22);
      q += DEC_OUT * 3;
    }
  }
  __syncthreads();
  if (x < DEC_OUT) {
#pragma unroll 2
    for (int y = 0; y < DEC_BAND; ++y) {
      const uint8_t* p = s_tmp + ((t_vmin[y] - r_lo) * DEC_OUT + x) * 3;
      const int* k = t_vk + y * KS;
      int acc0 = 1 << 21, acc1 = 1 << 21, acc2 = 1 << 21;
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const int kj = k[j];
        acc0 += p[j * DEC_OUT * 3] * kj; acc1 += p[j * DEC_OUT * 3 + 1] * kj; acc2 += p[j * DEC_OUT * 3 + 2] * kj;
      }
      put(y, x, 0, __ldg(a.lut + clip8(acc0 >> 22)));
      put(y, x, 1, __ldg(a.lut + 256 + clip8(acc1 >> 22)));
      put(y, x, 2, __ldg(a.lut + 512 + clip8(acc2 >> 22)));
    }
  }
}

__global__ void __launch_bounds__(DEC_THREADS)
decode_kernel(const DecodeArgs a) {
  extern __shared__ __align__(16) uint8_t dsm[];
  const int band = blockIdx.x;  // output rows [band*16, band*16+16)
  const int frame = blockIdx.y;
  const int y0 = band * DEC_BAND;
  const int tid = threadIdx.x;

  const int in_bytes = dec_in_bytes(a.src_w, a.max_rows);
  const int tmp_bytes = dec_tmp_bytes(a.max_rows);
  uint8_t* s_in = dsm;
  uint8_t* s_tmp = s_in + in_bytes;
  uint8_t* s_out = s_tmp + tmp_bytes;
  int* s_tab = reinterpret_cast<int*>(s_out + dec_out_bytes(a.out_kind));

  // ---- which source rows does this band need? ----
  int r_lo, r_hi;
  if (a.mode == DEC_PIL_BICUBIC) {
    r_lo = a.v_min[y0];
    r_hi = a.v_min[y0 + DEC_BAND - 1] + a.v_cnt[y0 + DEC_BAND - 1];
  } else {
    const float sc = static_cast<float>(a.src_h) / DEC_OUT;
    float s0 = fmaxf(sc * (y0 + 0.5f) - 0.5f, 0.f), s1 = fmaxf(sc * (y0 + DEC_BAND - 1 + 0.5f) - 0.5f, 0.f);
    r_lo = static_cast<int>(s0);
    r_hi = min(static_cast<int>(s1) + 2, a.src_h);
  }
  const int nrows = r_hi - r_lo;
  const int row_bytes = a.src_w * 3;

  // ---- stage source rows (16 B vectors when the geometry allows, bytes otherwise) ----
  const uint8_t* img = a.ob + static_cast<long long>(frame) * a.frame_stride;
  const long long pitch = static_cast<long long>(a.in_w) * 3;
  const uint8_t* src0 = img + (a.crop_top + r_lo) * pitch + a.crop_left * 3;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src0) | static_cast<uintptr_t>(pitch) |
                        static_cast<uintptr_t>(row_bytes)) & 15) == 0;
  if (vec_ok) {
    const int vpr = row_bytes >> 4;
    for (int i = tid; i < nrows * vpr; i += DEC_THREADS) {
      const int r = i / vpr, c = i - r * vpr;
      reinterpret_cast<uint4*>(s_in + r * row_bytes)[c] =
          __ldg(reinterpret_cast<const uint4*>(src0 + r * pitch) + c);
    }
  } else {
    for (int i = tid; i < nrows * row_bytes; i += DEC_THREADS) {
      const int r = i / row_bytes, c = i - r * row_bytes;
      s_in[r * row_bytes + c] = __ldg(src0 + r * pitch + c);
    }
  }

  float* s_outf = reinterpret_cast<float*>(s_out);
  op_t* s_outh = reinterpret_cast<op_t*>(s_out);
  const int P = a.patch, G = a.grid;
  const bool patch_out = a.out_kind == DEC_OUT_PATCH_BF16;
  // In patch mode the band is staged as [px][c][ph(16)][pw(P)] so that global writes are long runs.
  auto put = [&](int y, int x, int c, float v) {
    if (patch_out) {
      const int px = x / P, pw = x - px * P;
      s_outh[((px * 3 + c) * DEC_BAND + y) * P + pw] = float_to_op(v);
    } else {
      s_outf[(c * DEC_BAND + y) * DEC_OUT + x] = v;
    }
  };

  if (a.mode == DEC_PIL_BICUBIC) {
    // tables -> smem
    int* t_hmin = s_tab;
    int* t_hcnt = t_hmin + DEC_OUT;
    int* t_hk = t_hcnt + DEC_OUT;
    int* t_vmin = t_hk + DEC_OUT * a.h_ksize;
    int* t_vcnt = t_vmin + DEC_BAND;
    int* t_vk = t_vcnt + DEC_BAND;
    const bool fast = a.h_ksize == a.v_ksize && a.h_ksize >= 3 && a.h_ksize <= 8;
    if (!fast) {
      for (int i = tid; i < DEC_OUT; i += DEC_THREADS) { t_hmin[i] = a.h_min[i]; t_hcnt[i] = a.h_cnt[i]; }
      for (int i = tid; i < DEC_OUT * a.h_ksize; i += DEC_THREADS) t_hk[i] = a.h_k[i];
    }
    for (int i = tid; i < DEC_BAND; i += DEC_THREADS) { t_vmin[i] = a.v_min[y0 + i]; t_vcnt[i] = a.v_cnt[y0 + i]; }
    for (int i = tid; i < DEC_BAND * a.v_ksize; i += DEC_THREADS) t_vk[i] = a.v_k[y0 * a.v_ksize + i];
    __syncthreads();
    if (fast) {
      switch (a.h_ksize) {   // taps per output pixel after the host's re-packing: 5 for 256 -> 224, 4 when upscaling
        case 3: bicubic_band_fast<3>(a, s_in, s_tmp, nrows, row_bytes, r_lo, y0, t_vmin, t_vk, put); break;
        case 4: bicubic_band_fast<4>(a, s_in, s_tmp, nrows, row_bytes, r_lo, y0, t_vmin, t_vk, put); break;
        case 5: bicubic_band_fast<5>(a, s_in, s_tmp, nrows, row_bytes, r_lo, y0, t_vmin, t_vk, put); break;
        case 6: bicubic_band_fast<6>(a, s_in, s_tmp, nrows, row_bytes, r_lo, y0, t_vmin, t_vk, put); break;
        case 7: bicubic_band_fast<7>(a, s_in, s_tmp, nrows, row_bytes, r_lo, y0, t_vmin, t_vk, put); break;
        default: bicubic_band_fast<8>(a, s_in, s_tmp, nrows, row_bytes, r_lo, y0, t_vmin, t_vk, put); break;
      }
    } else {

    // pass 1: horizontal, every staged row -> uint8 tmp[r][x][c]
    for (int i = tid; i < nrows * DEC_OUT; i += DEC_THREADS) {
      const int r = i / DEC_OUT, x = i - r * DEC_OUT;
      const int xmin = t_hmin[x], cnt = t_hcnt[x];
      const int* k = t_hk + x * a.h_ksize;
      const uint8_t* p = s_in + r * row_bytes + xmin * 3;
      int acc0 = 1 << 21, acc1 = 1 << 21, acc2 = 1 << 21;
      for (int j = 0; j < cnt; ++j) {
        const int kj = k[j];
        acc0 += p[3 * j] * kj; acc1 += p[3 * j + 1] * kj; acc2 += p[3 * j + 2] * kj;
      }
      uint8_t* q = s_tmp + (r * DEC_OUT + x) * 3;
      q[0] = clip8(acc0 >> 22); q[1] = clip8(acc1 >> 22); q[2] = clip8(acc2 >> 22);
    }
    __syncthreads();
    // pass 2: vertical -> uint8 -> normalised value via the host-built table
    for (int i = tid; i < DEC_BAND * DEC_OUT; i += DEC_THREADS) {
      const int y = i / DEC_OUT, x = i - y * DEC_OUT;
      const int ymin = t_vmin[y] - r_lo, cnt = t_vcnt[y];
      const int* k = t_vk + y * a.v_ksize;
      const uint8_t* p = s_tmp + (ymin * DEC_OUT + x) * 3;
      int acc0 = 1 << 21, acc1 = 1 << 21, acc2 = 1 << 21;
      for (int j = 0; j < cnt; ++j) {
        const int kj = k[j];
        acc0 += p[j * DEC_OUT * 3] * kj; acc1 += p[j * DEC_OUT * 3 + 1] * kj; acc2 += p[j * DEC_OUT * 3 + 2] * kj;
      }
      put(y, x, 0, __ldg(a.lut + clip8(acc0 >> 22)));
      put(y, x, 1, __ldg(a.lut + 256 + clip8(acc1 >> 22)));
      put(y, x, 2, __ldg(a.lut + 512 + clip8(acc2 >> 22)));
    }
    }
  } else {
    __syncthreads();
    // float bilinear, align_corners=False, antialias off (ATen upsample_bilinear2d index/weight rule)
    const float sch = static_cast<float>(a.src_h) / DEC_OUT, scw = static_cast<float>(a.src_w) / DEC_OUT;
    const bool ident = (a.src_h == DEC_OUT) || (a.src_w == DEC_OUT);  // clip_multiscale_adapter.py:127
    for (int i = tid; i < DEC_BAND * DEC_OUT; i += DEC_THREADS) {
      const int y = i / DEC_OUT, x = i - y * DEC_OUT;
      float v[3];
      if (ident) {
        const uint8_t* p = s_in + ((y0 + y) - r_lo) * row_bytes + x * 3;
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
      } else {
        const float sy = fmaxf(sch * ((y0 + y) + 0.5f) - 0.5f, 0.f), sx = fmaxf(scw * (x + 0.5f) - 0.5f, 0.f);
        const int iy0 = static_cast<int>(sy), ix0 = static_cast<int>(sx);
        const int iy1 = iy0 + (iy0 < a.src_h - 1 ? 1 : 0), ix1 = ix0 + (ix0 < a.src_w - 1 ? 1 : 0);
        const float ly1 = sy - iy0, ly0 = 1.f - ly1, lx1 = sx - ix0, lx0 = 1.f - lx1;
        const uint8_t* p0 = s_in + (iy0 - r_lo) * row_bytes;
        const uint8_t* p1 = s_in + (iy1 - r_lo) * row_bytes;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float top = __fadd_rn(__fmul_rn(lx0, p0[ix0 * 3 + c]), __fmul_rn(lx1, p0[ix1 * 3 + c]));
          const float bot = __fadd_rn(__fmul_rn(lx0, p1[ix0 * 3 + c]), __fmul_rn(lx1, p1[ix1 * 3 + c]));
          v[c] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        put(y, x, c, __fdiv_rn(__fsub_rn(__fdiv_rn(v[c], 255.0f), a.mean[c]), a.stdv[c]));
    }
  }
  __syncthreads();

  // ---- write the band out ----
  if (patch_out) {
    op_t* A = reinterpret_cast<op_t*>(a.out);
    const int K = 3 * P * P;
    const int py = y0 / P, ph0 = y0 - py * P;  // band = 16 rows [ph0, ph0+16) of patch row py
    const int vec_per_run = P / 8;             // 16-byte vectors per (px,c,ph) run of P bf16
    const int nvec = G * 3 * DEC_BAND * vec_per_run;
    for (int i = tid; i < nvec; i += DEC_THREADS) {
      const int w8 = i % vec_per_run;
      int rest = i / vec_per_run;
      const int ph = rest % DEC_BAND; rest /= DEC_BAND;
      const int c = rest % 3;
      const int px = rest / 3;
      const size_t grow = static_cast<size_t>(frame) * a.tokens + 1 + py * G + px;
      const size_t gcol = static_cast<size_t>(c) * P * P + (ph0 + ph) * P + w8 * 8;
      *reinterpret_cast<uint4*>(A + grow * K + gcol) = reinterpret_cast<const uint4*>(s_outh)[i];
    }
    if (band == 0) {  // class-token slot: an all-zero A row (the GEMM epilogue adds cls + pos[0])
      uint4* z = reinterpret_cast<uint4*>(A + static_cast<size_t>(frame) * a.tokens * K);
      for (int i = tid; i < K / 8; i += DEC_THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
  } else {
    float* O = reinterpret_cast<float*>(a.out) + static_cast<size_t>(frame) * 3 * DEC_OUT * DEC_OUT;
    for (int i = tid; i < 3 * DEC_BAND * (DEC_OUT / 4); i += DEC_THREADS) {
      const int x4 = i % (DEC_OUT / 4);
      const int rest = i / (DEC_OUT / 4);
      const int y = rest % DEC_BAND, c = rest / DEC_BAND;
      reinterpret_cast<float4*>(O + (static_cast<size_t>(c) * DEC_OUT + y0 + y) * DEC_OUT)[x4] =
          reinterpret_cast<const float4*>(s_outf + (c * DEC_BAND + y) * DEC_OUT)[x4];
    }
  }
}

}  // namespace arp

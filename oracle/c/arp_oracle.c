/* ORACLE / TEST INFRASTRUCTURE — not part of the shipped product.
 *
 * Plain-C restatement of the byte/integer/order-exact pieces of the reference's reward-labeling path.
 * Built by oracle/Makefile into oracle/_build/liboracle.so; used only by tests/ and bench.py's CPU leg.
 *
 *   oracle_pil_bicubic_rgb   PIL.Image.resize(..., BICUBIC) as reached from label_reward.py:109-121
 *                            (Pillow ImagingResample, 8 bits per channel: two passes, uint8 intermediate,
 *                             22-bit fixed-point coefficients)
 *   oracle_normalize_chw     ToTensor + Normalize of the same transform (fp32, /255 then (v-mean)/std)
 *   oracle_discount_cumsum   label_reward.py:247-254, gamma = 1.0 path, sequential fp32 adds
 *   oracle_stack_outputs     label_reward.py:232-245, out[i,f] = x[max(0, i-(F-1-f))]
 *   oracle_episode_index     label_reward.py:82-83, [0] + (nonzero(done[:, -1]) + 1)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

/* bounds[2*xx] = first tap, bounds[2*xx+1] = tap count, kk[xx*ksize + j] = fixed-point weight */
static int precompute(int in_size, int out_size, int** bounds_out, int** kk_out) {
  double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 2.0 * filterscale;
  int ksize = (int)ceil(support) * 2 + 1;
  int* bounds = (int*)malloc(sizeof(int) * 2 * out_size);
  int* kk = (int*)calloc((size_t)out_size * ksize, sizeof(int));
  double* k = (double*)malloc(sizeof(double) * ksize);
  for (int xx = 0; xx < out_size; xx++) {
    double center = (xx + 0.5) * scale, ww = 0.0, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; x++) { k[x] = bicubic_filter((x + xmin - center + 0.5) * ss); ww += k[x]; }
    for (int x = 0; x < xmax; x++) {
      if (ww != 0.0) k[x] /= ww;
      double v = k[x] * (1 << PRECISION_BITS);
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  free(k);
  *bounds_out = bounds;
  *kk_out = kk;
  return ksize;
}

static uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

/* src uint8 [h][w][3] -> dst uint8 [out][out][3] */
int oracle_pil_bicubic_rgb(const uint8_t* src, int h, int w, uint8_t* dst, int out) {
  int *bx, *kx, *by, *ky;
  int ksx = precompute(w, out, &bx, &kx), ksy = precompute(h, out, &by, &ky);
  uint8_t* tmp = (uint8_t*)malloc((size_t)h * out * 3);
  if (!tmp) return -1;
  for (int y = 0; y < h; y++)
    for (int xx = 0; xx < out; xx++) {
      int xmin = bx[2 * xx], n = bx[2 * xx + 1];
      const int* k = kx + (size_t)xx * ksx;
      for (int c = 0; c < 3; c++) {
        int ss = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < n; x++) ss += src[((size_t)y * w + xmin + x) * 3 + c] * k[x];
        tmp[((size_t)y * out + xx) * 3 + c] = clip8(ss);
      }
    }
  for (int yy = 0; yy < out; yy++) {
    int ymin = by[2 * yy], n = by[2 * yy + 1];
    const int* k = ky + (size_t)yy * ksy;
    for (int xx = 0; xx < out; xx++)
      for (int c = 0; c < 3; c++) {
        int ss = 1 << (PRECISION_BITS - 1);
        for (int y = 0; y < n; y++) ss += tmp[((size_t)(ymin + y) * out + xx) * 3 + c] * k[y];
        dst[((size_t)yy * out + xx) * 3 + c] = clip8(ss);
      }
  }
  free(tmp); free(bx); free(kx); free(by); free(ky);
  return 0;
}

/* u8 [n][n][3] -> fp32 [3][n][n]: ToTensor (/255) then Normalize ((v - mean) / std), each rounded to fp32 */
void oracle_normalize_chw(const uint8_t* hwc, int n, float* chw) {
  static const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
  static const float std_[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < n * n; i++) {
      volatile float v = (float)hwc[(size_t)i * 3 + c] / 255.0f;
      volatile float d = v - mean[c];
      chw[(size_t)c * n * n + i] = d / std_[c];
    }
}

void oracle_discount_cumsum(const float* x, int64_t n, float gamma, float* out) {
  if (n <= 0) return;
  out[n - 1] = x[n - 1];
  for (int64_t t = n - 2; t >= 0; t--) {
    volatile float prod = gamma * out[t + 1]; /* numpy rounds the product before the add: no FMA */
    out[t] = x[t] + prod;
  }
}

void oracle_stack_outputs(const float* x, int64_t n, int num_frames, float* out) {
  for (int64_t i = 0; i < n; i++)
    for (int f = 0; f < num_frames; f++) {
      int64_t j = i - (num_frames - 1 - f);
      out[i * num_frames + f] = x[j < 0 ? 0 : j];
    }
}

/* returns the number of entries written to idx (<= cap) */
int64_t oracle_episode_index(const float* done_last, int64_t n, int64_t* idx, int64_t cap) {
  int64_t m = 0;
  if (m < cap) idx[m++] = 0;
  for (int64_t i = 0; i < n; i++)
    if (done_last[i] != 0.0f && m < cap) idx[m++] = i + 1;
  return m;
}

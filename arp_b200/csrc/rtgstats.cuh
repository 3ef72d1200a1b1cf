// Consumer-side return-to-go statistics: what ProcgenDataset.preprocess_rtgs
// (arp_dt/data_procgen.py:132-174) computes from the labeler's per-frame rewards before training —
//   reward_min / reward_max (:147-148), optional `reward - reward_min` (:150-153), the per-episode
//   return-to-go scan + window stack again (:155-168, same arithmetic as label_reward.py:232-254, done by
//   rtg_scan_stack_kernel), and the statistic that sets the conditioning token: max (CoinRun) or the
//   0.9-quantile (other games) of all stacked return-to-go values (:171-174).
// All byte/order-exact: min, max and order statistics are exact selections, the subtraction is one fp32 op.
// The quantile needs the two order statistics around (n-1)*q: an MSB-first 8-bit radix select (4 passes over
// the values, 256-bin histograms) finds the k-th smallest of n floats exactly, for two ranks at once.
#pragma once

#include "common.cuh"

namespace arp {

// order-preserving map float -> uint32 (total order of finite floats and infinities; -0 < +0)
__device__ __forceinline__ uint32_t f32_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// mm[0] = min key, mm[1] = max key (ordered uint32). Initialise to {0xffffffff, 0}.
__global__ void __launch_bounds__(256)
minmax_f32_kernel(const float* __restrict__ x, long long n, uint32_t* __restrict__ mm) {
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t k = f32_to_ordered(x[i]);
    lo = min(lo, k);
    hi = max(hi, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
  }
}

// out[0] = min, out[1] = max as floats
__global__ void minmax_finish_kernel(const uint32_t* __restrict__ mm, float* __restrict__ out) {
  out[0] = ordered_to_f32(mm[0]);
  out[1] = ordered_to_f32(mm[1]);
}

// y[i] = x[i] - *shift   (one rounded fp32 subtraction, numpy's `reward - self.reward_min[image_key]`)
__global__ void __launch_bounds__(256)
sub_scalar_f32_kernel(const float* __restrict__ x, const float* __restrict__ shift, float* __restrict__ y, long long n) {
  const float s = *shift;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = __fsub_rn(x[i], s);
}

// ---- radix select: k-th smallest (0-based rank) of n floats, for NQ ranks at once ----
constexpr int RSEL_NQ = 2;
struct RadixSelectState {
  uint32_t prefix[RSEL_NQ];        // key bits fixed so far (high bits)
  unsigned long long k[RSEL_NQ];   // rank within the elements matching the prefix
  unsigned int hist[RSEL_NQ][256];
};

__global__ void radix_select_init_kernel(RadixSelectState* st, unsigned long long k0, unsigned long long k1) {
  const int t = threadIdx.x;
  if (t < RSEL_NQ) { st->prefix[t] = 0; st->k[t] = t == 0 ? k0 : k1; }
  for (int i = t; i < RSEL_NQ * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
}

// pass p = 0..3 examines key bits [24-8p, 32-8p)
__global__ void __launch_bounds__(256)
radix_select_hist_kernel(const float* __restrict__ x, long long n, RadixSelectState* st, int pass) {
  __shared__ unsigned int sh[RSEL_NQ][256];
  for (int i = threadIdx.x; i < RSEL_NQ * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int shift = 24 - 8 * pass;
  const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
  const uint32_t p0 = st->prefix[0], p1 = st->prefix[1];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t k = f32_to_ordered(x[i]);
    const uint32_t d = (k >> shift) & 0xffu;
    if ((k & himask) == p0) atomicAdd(&sh[0][d], 1u);
    if ((k & himask) == p1) atomicAdd(&sh[1][d], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RSEL_NQ * 256; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&(&st->hist[0][0])[i], v);
  }
}

// one thread per rank walks the 256 bins, fixes the digit, re-bases k, clears the histogram
__global__ void radix_select_pick_kernel(RadixSelectState* st, int pass, float* __restrict__ out) {
  const int q = threadIdx.x;
  if (q < RSEL_NQ) {
    const int shift = 24 - 8 * pass;
    unsigned long long k = st->k[q], acc = 0;
    int d = 0;
    for (; d < 255; ++d) {
      const unsigned long long c = st->hist[q][d];
      if (acc + c > k) break;
      acc += c;
    }
    st->prefix[q] |= static_cast<uint32_t>(d) << shift;
    st->k[q] = k - acc;
    if (pass == 3) out[q] = ordered_to_f32(st->prefix[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RSEL_NQ * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0;
}

}  // namespace arp

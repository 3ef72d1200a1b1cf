// Microtest (dev tool): which (lane, column) does each thread register of the 16-lane tcgen05.ld / st shapes address?
// Writes lane*1000 + column with the 32x32b shape, reads it back with 16x256b.x2; writes thread-tagged values with
// 16x128b.x2 and reads them back with 32x32b.
#include <cstdio>
#include "../../arp_b200/csrc/attention_tc.cuh"
using namespace arp;
__global__ void k(int* out) {
  __shared__ uint32_t slot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<64>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  if (warp == 0) {
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = lane * 1000 + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(base), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory");
    tmem_st_wait();
    uint32_t r[8];
    // lanes 0..15, columns 0..15
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(base) : "memory");
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[lane * 8 + i] = r[i];
    // lanes 16..31
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]) : "r"(base + (16u << 16)) : "memory");
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[256 + lane * 8 + i] = r[i];
    // 16x128b.x2 store: thread-tagged values into columns 32.. of lanes 0..15
    uint32_t w[4];
    for (int i = 0; i < 4; ++i) w[i] = 100000 + lane * 10 + i;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + 32), "r"(w[0]),"r"(w[1]),"r"(w[2]),"r"(w[3]) : "memory");
    tmem_st_wait();
    uint32_t q[16];
    tmem_ld_32x16(base + 32, q);
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) out[512 + lane * 8 + i] = q[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(base); }
}
int main() {
  int* d; cudaMalloc(&d, 768 * 4); cudaMemset(d, 0, 768 * 4);
  k<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  int h[768]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("16x256b.x2 at lane base 0: thread: regs (lane*1000+col)\n");
  for (int t = 0; t < 32; ++t) { printf("t%2d:", t); for (int i = 0; i < 8; ++i) printf(" %6d", h[t * 8 + i]); printf("\n"); }
  printf("16x256b.x2 at lane base 16\n");
  for (int t = 0; t < 32; t += 5) { printf("t%2d:", t); for (int i = 0; i < 8; ++i) printf(" %6d", h[256 + t * 8 + i]); printf("\n"); }
  printf("16x128b.x2 store (100000 + thread*10 + reg) read back by lane (32x32b), columns 32..39\n");
  for (int t = 0; t < 32; ++t) { printf("lane%2d:", t); for (int i = 0; i < 8; ++i) printf(" %6d", h[512 + t * 8 + i]); printf("\n"); }
  return 0;
}

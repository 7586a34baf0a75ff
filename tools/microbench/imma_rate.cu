// Go/no-go probe for moving the Poseidon1 partial-section linear algebra (constant matrix x batch of states) onto the
// tensor cores as u8-limb integer MMA (round-1 verdict, item 6): issue rate of the legacy warp-level
// mma.sync.m16n8k32.u8.u8 (s32 accumulate) on sm_100a, alone and next to an IMAD.WIDE stream on the same warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench/imma_rate.cu -o tools/microbench/imma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MMA_PER_IT, int WIDE_PER_IT>
__global__ void k(int iters, uint32_t seed, long long* cycles, int* sink) {
  uint32_t a[4] = {seed + threadIdx.x, seed * 3 + 1, seed ^ 0x55aa, seed + 7}, b[2] = {seed * 5, seed * 7 + 3};
  int c[8][4] = {};
  uint64_t acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int m = 0; m < MMA_PER_IT; m++) imma(c[m & 7], a, b);
#pragma unroll
    for (int w = 0; w < WIDE_PER_IT; w++)
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[w & 7]) : "r"(a[w & 3]), "r"(b[w & 1]));
  }
  const long long t1 = clock64();
  int s = 0;
  for (int m = 0; m < 8; m++) s += c[m][0] + c[m][1] + c[m][2] + c[m][3] + (int)acc[m] + (int)(acc[m] >> 32);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int M, int W>
void run(const char* name, int warps_per_sm) {
  long long* d_c;
  int* d_s;
  cudaMalloc(&d_c, 8);
  cudaMalloc(&d_s, 148 * 1024 * 4);
  const int iters = 2000;
  k<M, W><<<148, warps_per_sm * 32>>>(iters, 12345, d_c, d_s);
  cudaDeviceSynchronize();
  k<M, W><<<148, warps_per_sm * 32>>>(iters, 12345, d_c, d_s);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  const double per_smsp_warps = warps_per_sm / 4.0;
  printf("%-28s warps/SM %2d: %8.2f cycles per iteration per warp", name, warps_per_sm, (double)c / iters);
  if (M) printf("  -> %6.2f cycles per MMA per SMSP", (double)c / iters / M / per_smsp_warps);
  if (W) printf("  (%d IMAD.WIDE per iteration)", W);
  printf("  [%s]\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_c), cudaFree(d_s);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<16, 0>("16 IMMA", w);
  for (int w : {4, 8, 16, 32}) run<0, 16>("16 IMAD.WIDE", w);
  for (int w : {4, 8, 16, 32}) run<16, 16>("16 IMMA + 16 IMAD.WIDE", w);
  for (int w : {8, 16}) run<4, 16>("4 IMMA + 16 IMAD.WIDE", w);
  return 0;
}

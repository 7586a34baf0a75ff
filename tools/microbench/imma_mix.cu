// Does IMMA.16832.U8.U8 (legacy mma.sync on sm_100a) overlap with the multiplier pipe?  Cycles per loop iteration of
// M IMMAs (independent accumulators) next to W loop-carried IMAD.WIDE (not hoistable: the multiplicand is the running
// accumulator), next to ALU-pipe adds, and next to DFMA, at 1..8 warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench/imma_mix.cu -o tools/microbench/imma_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int M, int W, int A, int D>
__global__ void k(int iters, uint32_t seed, long long* cycles, int* sink) {
  uint32_t a[4] = {seed + threadIdx.x, seed * 3 + 1, seed ^ 0x55aa, seed + 7}, b[2] = {seed * 5, seed * 7 + 3};
  int c[8][4] = {};
  uint64_t acc[8];
  uint32_t s[8];
  double d[8];
  for (int i = 0; i < 8; i++) acc[i] = seed * (i + 1) + threadIdx.x, s[i] = seed + i, d[i] = 1.0 + i + seed;
  const double dm = 1.0 + 1e-9 * seed;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (u < M) imma(c[u & 7], a, b);
      if (u < W) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[u & 7]) : "r"((uint32_t)acc[(u + 1) & 7]), "r"(b[u & 1]));
      if (u < A) asm volatile("add.u32 %0, %0, %1;" : "+r"(s[u & 7]) : "r"(s[(u + 1) & 7]));
      if (u < A) asm volatile("xor.b32 %0, %0, %1;" : "+r"(s[(u + 3) & 7]) : "r"(s[(u + 2) & 7]));
      if (u < D) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[u & 7]) : "d"(dm));
    }
  }
  const long long t1 = clock64();
  int r = 0;
  for (int m = 0; m < 8; m++) r += c[m][0] + c[m][1] + c[m][2] + c[m][3] + (int)acc[m] + (int)(acc[m] >> 32) + s[m] + (int)d[m];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int M, int W, int A, int D>
void run(int warps_per_sm) {
  long long* d_c;
  int* d_s;
  cudaMalloc(&d_c, 8);
  cudaMalloc(&d_s, 148 * 1024 * 4);
  const int iters = 2000;
  for (int rep = 0; rep < 2; rep++) {
    k<M, W, A, D><<<148, warps_per_sm * 32>>>(iters, 12345, d_c, d_s);
    cudaDeviceSynchronize();
  }
  long long c;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  printf("IMMA %2d  IMAD.WIDE %2d  ALU %2d  DFMA %2d  warps/SMSP %d: %8.2f cycles per iteration per SMSP  [%s]\n", M, W, 2 * A, D,
         warps_per_sm / 4, (double)c / iters / (warps_per_sm / 4.0), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_c), cudaFree(d_s);
}

int main() {
  for (int w : {4, 16, 32}) {
    run<16, 0, 0, 0>(w);
    run<0, 16, 0, 0>(w);
    run<16, 16, 0, 0>(w);
    run<8, 16, 0, 0>(w);
    run<4, 16, 0, 0>(w);
    run<0, 0, 16, 0>(w);
    run<16, 0, 16, 0>(w);
    run<0, 0, 0, 16>(w);
    run<16, 0, 0, 16>(w);
    run<8, 16, 16, 0>(w);
  }
  return 0;
}

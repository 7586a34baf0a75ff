// Second probe: true IMAD.WIDE issue cost, and whether the FP64 pipe (DFMA) runs concurrently with it.
// Every chain feeds the low AND high word back (xor) so ptxas can neither hoist the product nor drop a half.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define ITERS 1024
#define CH 8
__device__ __forceinline__ uint32_t wide_step(uint32_t x, uint32_t k) {
  uint32_t lo, hi;
  asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t;}" : "=r"(lo), "=r"(hi) : "r"(x), "r"(k));
  return lo ^ hi;
}
__device__ __forceinline__ uint32_t widea_step(uint32_t x, uint32_t k, uint64_t& acc) {
  uint32_t lo, hi;
  asm volatile("{mad.wide.u32 %2, %3, %4, %2; mov.b64 {%0,%1}, %2;}" : "=r"(lo), "=r"(hi), "+l"(acc) : "r"(x), "r"(k));
  return lo ^ hi;
}
__global__ void k_wide(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CH]; for (int c = 0; c < CH; c++) a[c] = threadIdx.x * 77 + c + seed;
  uint32_t k = seed | 1;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = wide_step(a[c], k);
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < CH; c++) s ^= a[c];
  if (s == 0x12345678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_widea(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CH]; uint64_t acc[CH]; for (int c = 0; c < CH; c++) { a[c] = threadIdx.x * 77 + c + seed; acc[c] = c; }
  uint32_t k = seed | 1;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = widea_step(a[c], k, acc[c]);
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < CH; c++) s ^= a[c] ^ (uint32_t)acc[c];
  if (s == 0x12345678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_imadlo(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CH]; for (int c = 0; c < CH; c++) a[c] = threadIdx.x * 77 + c + seed;
  uint32_t k = seed | 1;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) { asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a[c]) : "r"(k)); a[c] ^= k; }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < CH; c++) s ^= a[c];
  if (s == 0x12345678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dfma(uint32_t* out, long long* cyc, uint32_t seed) {
  double d[CH]; for (int c = 0; c < CH; c++) d[c] = threadIdx.x + c;
  double b = 1.0 + seed * 1e-9, kk = seed * 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(b), "d"(kk));
  }
  long long t1 = clock64();
  double s = 0; for (int c = 0; c < CH; c++) s += d[c];
  if (s == 0.12345) out[0] = 1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// CH wide chains + CH dfma chains interleaved
__global__ void k_wide_dfma(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CH]; double d[CH];
  for (int c = 0; c < CH; c++) { a[c] = threadIdx.x * 77 + c + seed; d[c] = threadIdx.x + c; }
  uint32_t k = seed | 1; double b = 1.0 + seed * 1e-9, kk = seed * 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      a[c] = wide_step(a[c], k);
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(b), "d"(kk));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; double sd = 0; for (int c = 0; c < CH; c++) { s ^= a[c]; sd += d[c]; }
  if (s == 0x12345678 || sd == 0.12345) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 1 wide : 2 dfma
__global__ void k_wide_2dfma(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CH]; double d[2 * CH];
  for (int c = 0; c < CH; c++) { a[c] = threadIdx.x * 77 + c + seed; d[c] = threadIdx.x + c; d[c + CH] = c; }
  uint32_t k = seed | 1; double b = 1.0 + seed * 1e-9, kk = seed * 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      a[c] = wide_step(a[c], k);
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(b), "d"(kk));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c + CH]) : "d"(b), "d"(kk));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; double sd = 0; for (int c = 0; c < CH; c++) { s ^= a[c]; sd += d[c] + d[c + CH]; }
  if (s == 0x12345678 || sd == 0.12345) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// u32 -> f64 via magic number + DADD, and back
__global__ void k_magic(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CH]; for (int c = 0; c < CH; c++) a[c] = threadIdx.x * 77 + c + seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      double x = __hiloint2double(0x43300000, a[c]);
      asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(-4503599627370496.0 + 3.0));
      asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(4503599627370496.0));
      a[c] = __double2loint(x) ^ __double2hiint(x);
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int c = 0; c < CH; c++) s ^= a[c];
  if (s == 0x12345678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <class K> void run(const char* name, K kern, double ops, int nsm, uint32_t* d_out, long long* d_cyc) {
  for (int threads : {256, 512, 1024}) {
    kern<<<nsm, threads>>>(d_out, d_cyc, 12345u);
    cudaDeviceSynchronize();
    kern<<<nsm, threads>>>(d_out, d_cyc, 12345u);
    cudaDeviceSynchronize();
    long long* h = new long long[nsm];
    cudaMemcpy(h, d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; i++) avg += h[i]; avg /= nsm;
    // cycles per (op-group) per warp per SMSP: warps per SMSP = threads/128
    double per = avg / ((double)ITERS * CH * (threads / 128.0));
    printf("%-16s thr=%4d  %.2f cycles per group per warp-slot (group = %s)\n", name, threads, per, name);
    delete[] h;
  }
}
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int nsm = prop.multiProcessorCount;
  uint32_t* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, sizeof(long long) * nsm);
  run("WIDE+LOP", k_wide, 1, nsm, d_out, d_cyc);
  run("WIDEacc+LOP", k_widea, 1, nsm, d_out, d_cyc);
  run("IMAD+LOP", k_imadlo, 1, nsm, d_out, d_cyc);
  run("DFMA", k_dfma, 1, nsm, d_out, d_cyc);
  run("WIDE+LOP+DFMA", k_wide_dfma, 1, nsm, d_out, d_cyc);
  run("WIDE+LOP+2DFMA", k_wide_2dfma, 1, nsm, d_out, d_cyc);
  run("MAGIC(2DADD+LOP)", k_magic, 1, nsm, d_out, d_cyc);
  return 0;
}

// Integer-pipe throughput probe for sm_100a: which of IMAD / IMAD.WIDE / IMAD.HI / IADD3 / DFMA ... the
// Poseidon1 and NTT kernels should be built from.  Prints lane-ops per clock per SM for each instruction
// (and a few mixes) with 8 independent dependency chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 2048
#define CHAINS 8

#define DEF_KERNEL(NAME, DECL, BODY, SINK)                                                        \
  __global__ void NAME(uint32_t* out, long long* cyc, uint32_t seed) {                            \
    DECL;                                                                                         \
    long long t0 = clock64();                                                                     \
    for (int it = 0; it < ITERS; it++) {                                                          \
      _Pragma("unroll") for (int c = 0; c < CHAINS; c++) { BODY; }                                \
    }                                                                                             \
    long long t1 = clock64();                                                                     \
    SINK;                                                                                         \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                              \
  }

#define DECL32 uint32_t a[CHAINS], b = seed | 1, k = seed * 7 + 3; for (int c = 0; c < CHAINS; c++) a[c] = threadIdx.x + c + seed
#define SINK32 uint32_t s = 0; for (int c = 0; c < CHAINS; c++) s ^= a[c]; if (s == 0x12345678) out[0] = s
#define DECL64 uint64_t a[CHAINS]; uint32_t b = seed | 1, k = seed * 7 + 3; for (int c = 0; c < CHAINS; c++) a[c] = threadIdx.x + c + seed
#define SINK64 uint64_t s = 0; for (int c = 0; c < CHAINS; c++) s ^= a[c]; if (s == 0x12345678) out[0] = (uint32_t)s

DEF_KERNEL(k_imad, DECL32, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b), "r"(k)), SINK32)
DEF_KERNEL(k_imad_hi, DECL32, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b), "r"(k)), SINK32)
DEF_KERNEL(k_iadd, DECL32, asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), SINK32)
DEF_KERNEL(k_min, DECL32, asm volatile("min.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), SINK32)
DEF_KERNEL(k_lop, DECL32, asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), SINK32)
DEF_KERNEL(k_shf, DECL32, asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[c]) : "r"(b)), SINK32)
DEF_KERNEL(k_wide, DECL64, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[c]) : "r"(b), "r"(k)), SINK64)
// wide multiply whose multiplicand depends on the previous result (like a Montgomery chain)
DEF_KERNEL(k_wide_dep, DECL64,
           asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(a[c]) : "r"(b)), SINK64)
DEF_KERNEL(k_add64, DECL64, asm volatile("add.u64 %0, %0, %1;" : "+l"(a[c]) : "l"((uint64_t)b << 7 | k)), SINK64)
// mixes
DEF_KERNEL(k_mix_wide_iadd, DECL64,
           asm volatile("{.reg .u32 lo, hi; mad.wide.u32 %0, %1, %2, %0; mov.b64 {lo, hi}, %0; add.u32 lo, lo, %1; "
                        "mov.b64 %0, {lo, hi};}" : "+l"(a[c]) : "r"(b), "r"(k)), SINK64)
DEF_KERNEL(k_mix_imad_iadd, DECL32,
           asm volatile("mad.lo.u32 %0, %0, %1, %2; add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b), "r"(k)), SINK32)
DEF_KERNEL(k_mix_imad_min, DECL32,
           asm volatile("mad.lo.u32 %0, %0, %1, %2; min.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b), "r"(k)), SINK32)
DEF_KERNEL(k_mix_wide_imad, DECL64,
           asm volatile("{.reg .u32 lo, hi; mad.wide.u32 %0, %1, %2, %0; mov.b64 {lo, hi}, %0; mad.lo.u32 lo, lo, %1, %2; "
                        "mov.b64 %0, {lo, hi};}" : "+l"(a[c]) : "r"(b), "r"(k)), SINK64)

__global__ void k_dfma(uint32_t* out, long long* cyc, uint32_t seed) {
  double a[CHAINS], b = 1.0 + seed * 1e-9, k = seed * 1e-7;
  for (int c = 0; c < CHAINS; c++) a[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[c]) : "d"(b), "d"(k));
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CHAINS; c++) s += a[c];
  if (s == 0.12345) out[0] = 1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma(uint32_t* out, long long* cyc, uint32_t seed) {
  float a[CHAINS], b = 1.0f + seed * 1e-9f, k = seed * 1e-7f;
  for (int c = 0; c < CHAINS; c++) a[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(b), "f"(k));
  }
  long long t1 = clock64();
  float s = 0;
  for (int c = 0; c < CHAINS; c++) s += a[c];
  if (s == 0.12345f) out[0] = 1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// I2F / F2I conversions used by a hypothetical FP64 MDS
__global__ void k_i2d(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[CHAINS];
  for (int c = 0; c < CHAINS; c++) a[c] = threadIdx.x + c + seed;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++)
      asm volatile("{.reg .f64 d; cvt.rn.f64.u32 d, %0; cvt.rzi.u32.f64 %0, d;}" : "+r"(a[c]));
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int c = 0; c < CHAINS; c++) s ^= a[c];
  if (s == 0x12345678) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class K>
void run(const char* name, K kern, int ops_per_body, int threads, int blocks_per_sm, int nsm, uint32_t* d_out, long long* d_cyc) {
  int blocks = nsm * blocks_per_sm;
  kern<<<blocks, threads>>>(d_out, d_cyc, 12345u);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(d_out, d_cyc, 12345u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[blocks];
  cudaMemcpy(h, d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; i++) avg += h[i];
  avg /= blocks;
  double lane_ops_per_sm = (double)threads * blocks_per_sm * ITERS * CHAINS * ops_per_body;
  printf("%-18s thr=%4d bps=%d  %.1f lane-ops/clk/SM  (%.3f ms, %.0f cycles, eff clk %.0f MHz)\n", name, threads,
         blocks_per_sm, lane_ops_per_sm / avg, ms, avg, avg / (ms * 1e3));
  delete[] h;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int nsm = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, nsm);
  uint32_t* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, 4);
  cudaMalloc(&d_cyc, sizeof(long long) * nsm * 8);
  for (int threads : {256, 1024}) {
    int bps = threads == 256 ? 4 : 1;
    run("IMAD", k_imad, 1, threads, bps, nsm, d_out, d_cyc);
    run("IMAD.HI", k_imad_hi, 1, threads, bps, nsm, d_out, d_cyc);
    run("IMAD.WIDE", k_wide, 1, threads, bps, nsm, d_out, d_cyc);
    run("IMAD.WIDE dep", k_wide_dep, 1, threads, bps, nsm, d_out, d_cyc);
    run("IADD", k_iadd, 1, threads, bps, nsm, d_out, d_cyc);
    run("ADD64", k_add64, 1, threads, bps, nsm, d_out, d_cyc);
    run("UMIN", k_min, 1, threads, bps, nsm, d_out, d_cyc);
    run("XOR", k_lop, 1, threads, bps, nsm, d_out, d_cyc);
    run("SHF", k_shf, 1, threads, bps, nsm, d_out, d_cyc);
    run("FFMA", k_ffma, 1, threads, bps, nsm, d_out, d_cyc);
    run("DFMA", k_dfma, 1, threads, bps, nsm, d_out, d_cyc);
    run("I2D+D2I", k_i2d, 2, threads, bps, nsm, d_out, d_cyc);
    run("WIDE+IADD", k_mix_wide_iadd, 2, threads, bps, nsm, d_out, d_cyc);
    run("IMAD+IADD", k_mix_imad_iadd, 2, threads, bps, nsm, d_out, d_cyc);
    run("IMAD+UMIN", k_mix_imad_min, 2, threads, bps, nsm, d_out, d_cyc);
    run("WIDE+IMAD", k_mix_wide_imad, 2, threads, bps, nsm, d_out, d_cyc);
  }
  return 0;
}

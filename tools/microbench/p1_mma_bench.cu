// Poseidon1 with the full-round MDS on the tensor cores (csrc/poseidon1_mma.cuh) against the one-state-per-thread form
// (csrc/poseidon1.cuh): bit-exactness on random states and time of a chain of compressions at the leaf-sponge shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leanmultisig_b200/csrc tools/microbench/p1_mma_bench.cu -o tools/microbench/p1_mma_bench
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon1_mma.cuh"
#include "poseidon1_umma.cuh"

using namespace lm;

__constant__ P1Tables c_p1 =
#include "poseidon1_tables.inc"
    ;

#ifndef THREADS
#define THREADS 256
#endif
#ifndef USYNC
#define USYNC 1
#endif
#ifndef MINB
#define MINB 2
#endif

// chain: state <- compress(digest | f(digest, it)), `iters` times; out = digest
__global__ void __launch_bounds__(THREADS, MINB) scalar_chain(const uint32_t* in, uint32_t* out, int iters, int full) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s[16];
  for (int k = 0; k < 16; k++) s[k] = in[16 * i + k];
  if (full) {
    p1_permute<16, P1Tables, true>(s, c_p1);
    for (int k = 0; k < 16; k++) out[16 * i + k] = s[k];
    return;
  }
  for (int it = 0; it < iters; it++) {
    for (int k = 0; k < 8; k++) s[8 + k] = kb_add(s[k], (uint32_t)(it * 8 + k));
    p1_compress<8, P1Tables, true>(s, c_p1);
  }
  for (int k = 0; k < 8; k++) out[16 * i + k] = s[k];
}

__global__ void __launch_bounds__(THREADS, MINB) mma_chain(const uint32_t* in, uint32_t* out, int iters, int full) {
  __shared__ P1wSmem<THREADS / 32> sm;
  p1w_smem_init(sm, c_p1);
  const P1wB B = p1w_b_init();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32 * 32;
  uint32_t f[4][4];
  for (int m = 0; m < 4; m++) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + 16 * (base + g + 8 * m) + 4 * t);
    f[m][0] = v.x, f[m][1] = v.y, f[m][2] = v.z, f[m][3] = v.w;
  }
  if (full) {
    p1w_permute<true>(f, B, sm, c_p1);
    for (int m = 0; m < 4; m++)
      *reinterpret_cast<uint4*>(out + 16 * (base + g + 8 * m) + 4 * t) = make_uint4(f[m][0], f[m][1], f[m][2], f[m][3]);
    return;
  }
  for (int it = 0; it < iters; it++) {
    // lanes t = 2, 3 (elements 8..15) take f(digest) from lanes t = 0, 1
    for (int m = 0; m < 4; m++)
      for (int q = 0; q < 4; q++) {
        const uint32_t d = __shfl_sync(0xffffffffu, f[m][q], (lane & ~3) | (t & 1));
        if (t >= 2) f[m][q] = kb_add(d, (uint32_t)(it * 8 + 4 * (t - 2) + q));
      }
    p1w_compress<true>(f, B, sm, c_p1);
  }
  if (t < 2)
    for (int m = 0; m < 4; m++)
      *reinterpret_cast<uint4*>(out + 16 * (base + g + 8 * m) + 4 * t) = make_uint4(f[m][0], f[m][1], f[m][2], f[m][3]);
}

__global__ void __launch_bounds__(THREADS, MINB) umma_chain(const uint32_t* in, uint32_t* out, int iters, int full, const uint8_t* b_image) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  P1uCtx c = p1u_setup(dsm, b_image, THREADS / 128);
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s[16];
  for (int k = 0; k < 16; k++) s[k] = in[16 * i + k];
  if (full) {
    p1u_permute<16, USYNC != 0>(c, s, c_p1);
    for (int k = 0; k < 16; k++) out[16 * i + k] = s[k];
  } else {
    for (int it = 0; it < iters; it++) {
      for (int k = 0; k < 8; k++) s[8 + k] = kb_add(s[k], (uint32_t)(it * 8 + k));
      p1u_compress<8, USYNC != 0>(c, s, c_p1);
    }
    for (int k = 0; k < 8; k++) out[16 * i + k] = s[k];
  }
  p1u_teardown(c, THREADS / 128);
}

static const P1Tables h_p1 =
#include "poseidon1_tables.inc"
    ;

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 0) : (1ull << 20);
  const int iters = argc > 2 ? atoi(argv[2]) : 8;
  std::vector<uint32_t> h(16 * n);
  uint64_t x = 88172645463325252ull;
  for (auto& v : h) {
    x ^= x << 13, x ^= x >> 7, x ^= x << 17;
    v = (uint32_t)(x % KB_P);
  }
  for (int k = 0; k < 16; k++) h[k] = KB_P - 1, h[16 + k] = 0;
  uint32_t *d_in, *d_a, *d_b;
  cudaMalloc(&d_in, 64 * n), cudaMalloc(&d_a, 64 * n), cudaMalloc(&d_b, 64 * n);
  cudaMemcpy(d_in, h.data(), 64 * n, cudaMemcpyHostToDevice);
  std::vector<uint32_t> a(16 * n), b(16 * n);
  std::vector<uint8_t> img(P1U_B_BYTES);
  p1u_build_b_image(h_p1, img.data());
  uint8_t* d_img;
  cudaMalloc(&d_img, P1U_B_BYTES);
  cudaMemcpy(d_img, img.data(), P1U_B_BYTES, cudaMemcpyHostToDevice);
  const int dyn = p1u_smem_bytes(THREADS / 128);
  cudaFuncSetAttribute(umma_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  const int variant = argc > 3 ? atoi(argv[3]) : 2;  // 1 = mma.sync, 2 = tcgen05
  int rc = 0;
  for (int full = 1; full >= 0; full--) {
    cudaMemset(d_a, 0, 64 * n), cudaMemset(d_b, 0, 64 * n);
    scalar_chain<<<n / THREADS, THREADS>>>(d_in, d_a, iters, full);
    if (variant == 1)
      mma_chain<<<n / THREADS, THREADS>>>(d_in, d_b, iters, full);
    else
      umma_chain<<<n / THREADS, THREADS, dyn>>>(d_in, d_b, iters, full, d_img);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return printf("CUDA error: %s\n", cudaGetErrorString(e)), 1;
    cudaMemcpy(a.data(), d_a, 64 * n, cudaMemcpyDeviceToHost), cudaMemcpy(b.data(), d_b, 64 * n, cudaMemcpyDeviceToHost);
    uint64_t bad = 0, first = ~0ull;
    for (uint64_t i = 0; i < 16 * n; i++)
      if (a[i] != b[i]) {
        if (first == ~0ull) first = i;
        bad++;
      }
    printf("%s: %llu mismatching words of %llu", full ? "permutation" : "compression chain", (unsigned long long)bad,
           (unsigned long long)(16 * n));
    if (bad) printf(" (first: state %llu word %llu scalar %08x mma %08x)", (unsigned long long)(first / 16),
                    (unsigned long long)(first % 16), a[first], b[first]), rc = 1;
    printf("\n");
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int which = 0; which < 2; which++) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (which == 0)
        scalar_chain<<<n / THREADS, THREADS>>>(d_in, d_a, iters, 0);
      else if (variant == 1)
        mma_chain<<<n / THREADS, THREADS>>>(d_in, d_b, iters, 0);
      else
        umma_chain<<<n / THREADS, THREADS, dyn>>>(d_in, d_b, iters, 0, d_img);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-6s %llu states x %d compressions: %.3f ms  (%.3f G compressions/s)\n", which ? (variant == 1 ? "mma.sync" : "tcgen05") : "scalar",
           (unsigned long long)n, iters, best, n * (double)iters / best * 1e-6);
  }
  return rc;
}

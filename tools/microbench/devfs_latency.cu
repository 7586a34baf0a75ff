// Latency of the device challenger's pieces, one warp: cycles per permutation, per EF inverse, per out-of-line EF product.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leanmultisig_b200/csrc tools/microbench/devfs_latency.cu -o /tmp/devfs_latency
#include <cstdio>
#include <cuda_runtime.h>
#include "devfs.cuh"
#include "reduce.cuh"
using namespace lm;

__global__ void k(DevFs* fs, uint32_t* tr, long long* out, uint32_t* sink) {
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  __shared__ uint32_t sbuf[64];
  fs_load_rc(rc_s);
  FsWarp w;
  w.load(fs, tr, rc_s);
  long long t0 = clock64();
  for (int i = 0; i < 16; i++) w.permute();
  long long t1 = clock64();
  Ef a{{w.x | 1, 3, 5, 7, 11}}, inv;
  for (int i = 0; i < 16; i++) {
    fs_ef_inv(a, &inv);
    a = ef_add(a, inv);
  }
  long long t2 = clock64();
  for (int i = 0; i < 64; i++) a = fs_ef_mul(a, inv);
  long long t3 = clock64();
  if ((threadIdx.x & 31) == 0) st_ef(sbuf, a), st_ef(sbuf + 5, inv), st_ef(sbuf + 10, a);
  __syncwarp();
  for (int i = 0; i < 16; i++) w.add_sumcheck_polynomial_bare(sbuf, 3, inv);
  long long t4 = clock64();
  for (int i = 0; i < 16; i++) a = ef_add(a, w.sample_ef()), w.fresh = true;
  long long t5 = clock64();
  w.store();
  if (threadIdx.x == 0) {
    out[0] = (t1 - t0) / 16, out[1] = (t2 - t1) / 16, out[2] = (t3 - t2) / 64, out[3] = (t4 - t3) / 16, out[4] = (t5 - t4) / 16;
    sink[0] = a.c[0];
  }
}

int main() {
  DevFs* fs;
  uint32_t *tr, *sink;
  long long* out;
  cudaMalloc(&fs, sizeof(DevFs));
  cudaMemset(fs, 0, sizeof(DevFs));
  DevFs h{};
  h.cap_words = 1 << 16;
  cudaMemcpy(fs, &h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&tr, 4 << 16);
  cudaMalloc(&sink, 64);
  cudaMalloc(&out, 64);
  for (int rep = 0; rep < 3; rep++) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<1, 32>>>(fs, tr, out, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long o[5];
    cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
    printf("cycles: permute %lld, ef_inv %lld, ef_mul(call) %lld, add_sumcheck_polynomial(3 coeffs) %lld, sample %lld; kernel %.1f us (%s)\n", o[0], o[1],
           o[2], o[3], o[4], ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

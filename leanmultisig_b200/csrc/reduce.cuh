// Shared device helpers: EF loads/stores and the (c0, c2) pair reduction used by every sumcheck round kernel.
#pragma once
#include <cstdint>
#include "kb.cuh"

namespace lm {

__device__ __forceinline__ Ef ld_ef(const uint32_t* p) {
  Ef v;
#pragma unroll
  for (int c = 0; c < 5; c++) v.c[c] = __ldg(p + c);
  return v;
}
__device__ __forceinline__ Ef ld_ef_rw(const uint32_t* p) {
  Ef v;
#pragma unroll
  for (int c = 0; c < 5; c++) v.c[c] = p[c];
  return v;
}
__device__ __forceinline__ void st_ef(uint32_t* p, const Ef& v) {
#pragma unroll
  for (int c = 0; c < 5; c++) p[c] = v.c[c];
}

// block reduction of (c0, c2) and write of one partial (10 words) per CTA
__device__ __forceinline__ void block_reduce_pair(Ef c0, Ef c2, uint32_t* __restrict__ partial) {
  __shared__ Ef red0[32], red2[32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Ef o0, o2;
#pragma unroll
    for (int c = 0; c < 5; c++) {
      o0.c[c] = __shfl_down_sync(0xffffffffu, c0.c[c], off);
      o2.c[c] = __shfl_down_sync(0xffffffffu, c2.c[c], off);
    }
    c0 = ef_add(c0, o0);
    c2 = ef_add(c2, o2);
  }
  const int t = threadIdx.x;
  if ((t & 31) == 0) red0[t >> 5] = c0, red2[t >> 5] = c2;
  __syncthreads();
  if (t < 32) {
    const int nw = blockDim.x >> 5;
    c0 = t < nw ? red0[t] : ef_zero();
    c2 = t < nw ? red2[t] : ef_zero();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Ef o0, o2;
#pragma unroll
      for (int c = 0; c < 5; c++) {
        o0.c[c] = __shfl_down_sync(0xffffffffu, c0.c[c], off);
        o2.c[c] = __shfl_down_sync(0xffffffffu, c2.c[c], off);
      }
      c0 = ef_add(c0, o0);
      c2 = ef_add(c2, o2);
    }
    if (t == 0) {
#pragma unroll
      for (int c = 0; c < 5; c++) partial[10 * blockIdx.x + c] = c0.c[c], partial[10 * blockIdx.x + 5 + c] = c2.c[c];
    }
  }
}

// out[0..5) = sum_k partial[10k + 0..5),  out[5..10) = sum_k partial[10k + 5..10)
static __global__ void sum_pair_partials_kernel(const uint32_t* __restrict__ partial, int n, uint32_t* __restrict__ out) {
  Ef c0 = ef_zero(), c2 = ef_zero();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    c0 = ef_add(c0, ld_ef_rw(partial + 10 * i));
    c2 = ef_add(c2, ld_ef_rw(partial + 10 * i + 5));
  }
  // reuse the CTA reduction; the single CTA writes partial slot 0 of `out`
  block_reduce_pair(c0, c2, out);
}


}  // namespace lm

// Prefix eq tables shared by the LSB-first sumchecks (quotient GKR layers, AIR sessions).
//
// A sumcheck over k variables that binds the least-significant variable first weighs pair j of round `rnd` with
// eq(point[0 .. m), j), m = k - 1 - rnd (the reference's SplitEq, crates/backend/sumcheck/src/split_eq.rs:5-103, rebuilt
// per round there).  Here every table a session will need is built ONCE: eq over point[0 .. t) for t <= h ("hi" prefixes)
// and over point[h .. h + t) ("lo" prefixes), h = k / 2, so that
//     eq(point[0 .. m), j) = HI_min(m,h)[j >> lo_bits] * LO_(m - min(m,h))[j & mask]
// for every round; 2^(h+1) + 2^(k-h) entries in all.  Tables are AoS (5 words per entry), big-endian index.
#pragma once
#include <cstddef>
#include <cstdint>
#include "kb.cuh"
#include "reduce.cuh"

namespace lm {

LM_HD uint32_t eqtab_split(uint32_t k) { return k / 2; }  // h for k variables (k - 1 table coordinates)
LM_HD size_t eqtab_hi(uint32_t t) { return 5 * (((size_t)1 << t) - 1); }
LM_HD size_t eqtab_lo(uint32_t h, uint32_t t) { return 5 * (((size_t)2 << h) - 1) + 5 * (((size_t)1 << t) - 1); }
inline size_t eqtab_words(uint32_t max_vars) {
  const uint32_t h = eqtab_split(max_vars), l = max_vars - h;
  return 5 * (((size_t)2 << h) + ((size_t)2 << l)) + 64;
}

#ifdef __CUDACC__
// tables t = 0 .. cnt over pts[0 .. cnt), table 0 = {first}; whole CTA, ends with a barrier
__device__ __forceinline__ void eqtab_build_prefix(uint32_t* tab, const Ef* pts, int cnt, const Ef& first) {
  if (threadIdx.x == 0) st_ef(tab, first);
  for (int t = 0; t < cnt; t++) {
    __syncthreads();
    const uint32_t* src = tab + 5 * (((size_t)1 << t) - 1);
    uint32_t* dst = tab + 5 * (((size_t)2 << t) - 1);
    const Ef p = pts[t];
    for (uint32_t i = threadIdx.x; i < (1u << t); i += blockDim.x) {
      const Ef e = ld_ef_rw(src + 5 * i);
      const Ef hi = ef_mul(e, p);
      st_ef(dst + 5 * (2 * i), ef_sub(e, hi));
      st_ef(dst + 5 * (2 * i + 1), hi);
    }
  }
  __syncthreads();
}
// every prefix table of a k-variable sumcheck whose point (k coordinates, the last one never enters a table) is `point`
__device__ __forceinline__ void eqtab_build(uint32_t* tab, const Ef* point, uint32_t k, const Ef& scale) {
  const uint32_t h = eqtab_split(k), cnt = k ? k - 1 : 0;
  const uint32_t nh = cnt < h ? cnt : h;
  eqtab_build_prefix(tab, point, (int)nh, scale);
  eqtab_build_prefix(tab + eqtab_lo(h, 0), point + h, (int)(cnt - nh), Ef{{KB_R1, 0, 0, 0, 0}});
}
// weight of pair j in the round with m free variables
struct EqView {
  const uint32_t* hi;
  const uint32_t* lo;
  uint32_t lo_bits;
  __device__ __forceinline__ EqView(const uint32_t* tab, uint32_t k, uint32_t m) {
    const uint32_t h = eqtab_split(k), hb = m < h ? m : h;
    lo_bits = m - hb;
    hi = tab + eqtab_hi(hb);
    lo = tab + eqtab_lo(h, lo_bits);
  }
  __device__ __forceinline__ Ef operator()(uint64_t j) const {
    return ef_mul(ld_ef_rw(hi + 5 * (j >> lo_bits)), ld_ef_rw(lo + 5 * (j & (((uint64_t)1 << lo_bits) - 1))));
  }
};
#endif

}  // namespace lm

// Value types the AIR constraint code is generic over: Fb (base field, first round) and Ef (after the first fold),
// plus the per-session constants.  Shared by air.cu (execution table) and air_generic.cu (extension_op, poseidon16).
#pragma once
#include <cstdint>
#include "kb.cuh"

namespace lm {

constexpr int AIR_LO = 10;  // eq factor split: eq_hi[j >> AIR_LO] * eq_lo[j & 1023]  (SplitEq, split_eq.rs:5-103)

struct AirExtra {
  Ef alpha[16];  // alpha powers, first 13 used
  Ef la[8];      // logup_alphas_eq_poly (first 4 + last are used by the bus column)
  Ef la_last;
  Ef beta;
};

// ---- value types the constraint code is generic over ------------------------------------------------------
struct Fb {
  uint32_t v;
};
__device__ __forceinline__ Fb operator+(Fb a, Fb b) { return Fb{kb_add(a.v, b.v)}; }
__device__ __forceinline__ Fb operator-(Fb a, Fb b) { return Fb{kb_sub(a.v, b.v)}; }
__device__ __forceinline__ Fb operator*(Fb a, Fb b) { return Fb{kb_mul(a.v, b.v)}; }
__device__ __forceinline__ Fb operator-(Fb a) { return Fb{kb_neg(a.v)}; }
__device__ __forceinline__ Ef operator+(const Ef& a, const Ef& b) { return ef_add(a, b); }
__device__ __forceinline__ Ef operator-(const Ef& a, const Ef& b) { return ef_sub(a, b); }
__device__ __forceinline__ Ef operator*(const Ef& a, const Ef& b) { return ef_mul(a, b); }
__device__ __forceinline__ Ef operator-(const Ef& a) {
  Ef r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.c[i] = kb_neg(a.c[i]);
  return r;
}
// constants and scalings
__device__ __forceinline__ Fb add_one(Fb a) { return Fb{kb_add(a.v, KB_R1)}; }
__device__ __forceinline__ Fb sub_one(Fb a) { return Fb{kb_sub(a.v, KB_R1)}; }
__device__ __forceinline__ Ef add_one(Ef a) { return ef_add_base(a, KB_R1); }
__device__ __forceinline__ Ef sub_one(Ef a) {
  a.c[0] = kb_sub(a.c[0], KB_R1);
  return a;
}
__device__ __forceinline__ Fb dbl(Fb a) { return a + a; }
__device__ __forceinline__ Ef dbl(const Ef& a) { return ef_add(a, a); }
__device__ __forceinline__ uint32_t kb_halve(uint32_t a) { return (a & 1) ? (a >> 1) + ((KB_P + 1) >> 1) : (a >> 1); }
__device__ __forceinline__ Fb halve(Fb a) { return Fb{kb_halve(a.v)}; }
__device__ __forceinline__ Ef halve(Ef a) {
#pragma unroll
  for (int i = 0; i < 5; i++) a.c[i] = kb_halve(a.c[i]);
  return a;
}
// EF scalar times value
__device__ __forceinline__ Ef scale(const Ef& s, Fb x) { return ef_mul_base(s, x.v); }
__device__ __forceinline__ Ef scale(const Ef& s, const Ef& x) { return ef_mul(s, x); }
__device__ __forceinline__ Ef add_val(const Ef& e, Fb x) { return ef_add_base(e, x.v); }
__device__ __forceinline__ Ef add_val(const Ef& e, const Ef& x) { return ef_add(e, x); }

}  // namespace lm

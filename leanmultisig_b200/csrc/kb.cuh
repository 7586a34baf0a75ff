// KoalaBear (p = 2^31 - 2^24 + 1) arithmetic for sm_100a, Montgomery form (x * 2^32 mod p) in u32.
//
// Replaces, on the device, the reference's scalar + AVX/NEON field code:
//   crates/backend/koala-bear/src/monty_31/utils.rs:65-127   (monty_add / monty_sub / monty_reduce)
//   crates/backend/koala-bear/src/monty_31/monty_31.rs:677-685 (Mul)
//   crates/backend/koala-bear/src/quintic_extension/extension.rs:531-548 (quintic_mul)
// Values crossing any kernel boundary are canonical in [0, p), exactly what the reference stores
// (monty_31.rs:32-42), so results are bit-identical whatever instruction sequence produced them.
//
// Instruction notes (sm_100a): a Montgomery product is IMAD.WIDE.U32 (a*b), two SHF + IADD3 (m = lo * p^-1),
// IMAD.WIDE.U32 (m*p), IADD3 (hi - hi + p) with the result in (0, 2p]; canonicalisation is one VIADDMNMX.U32.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

#define LM_HD __host__ __device__ __forceinline__

namespace lm {

constexpr uint32_t KB_P = 0x7f000001u;
constexpr uint32_t KB_NEG_MU = 0x7effffffu;  // -p^-1 mod 2^32
constexpr uint32_t KB_R1 = 0x01fffffeu;      // 2^32 mod p  (Montgomery form of 1)
constexpr uint32_t KB_R2 = 0x17f7efe4u;      // 2^64 mod p  (checked in kb_selfcheck)

LM_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// x in [0, 2p) -> [0, p)
LM_HD uint32_t kb_canon(uint32_t x) { return umin32(x, x - KB_P); }
LM_HD uint32_t kb_add(uint32_t a, uint32_t b) { return kb_canon(a + b); }
LM_HD uint32_t kb_sub(uint32_t a, uint32_t b) {
  uint32_t d = a - b;
  return umin32(d, d + KB_P);
}
LM_HD uint32_t kb_neg(uint32_t a) { return a ? KB_P - a : 0u; }

// ---- 32x32 -> 64 multiply-add, pinned to IMAD.WIDE.U32 on the device ------------------------------------
// Measured on B200 (tools/microbench/int_pipes.cu, profiles/r01_int_pipes.txt): IMAD and IMAD.WIDE issue at
// 64 lanes/clk/SM, IMAD.HI at 32, IADD3/IMNMX at 128 on the other pipe.  ptxas rewrites a Montgomery reduction
// whose modulus it can see into IMAD.HI (half rate), and multiplications by small literals into shift/IMAD.HI
// chains, so the modulus and the MDS coefficients are read from constant memory, which it cannot fold.
#ifdef __CUDACC__
struct KbOpaque {
  uint32_t p;
  uint32_t mds[16];
  double mds_d[16];
  uint32_t s24, s31;  // shift amounts of p^-1 = 2^31 + 2^24 + 1, opaque so that ptxas keeps the shifts (see kb_redc_lazy)
};
static __constant__ KbOpaque c_kb = {KB_P,
                                     {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1},
                                     {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1},
                                     24, 31};
#endif
#ifdef __CUDA_ARCH__
#define LM_KB_P_OPAQUE (c_kb.p)
#else
#define LM_KB_P_OPAQUE KB_P
#endif

LM_HD uint64_t mul_wide(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint64_t d;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
  return d;
#else
  return (uint64_t)a * b;
#endif
}
LM_HD uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
#ifdef __CUDA_ARCH__
  uint64_t d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
#else
  return (uint64_t)a * b + c;
#endif
}

// Montgomery reduction without the final conditional subtraction: returns a value congruent to t / 2^32 in
// (t / 2^32, t / 2^32 + p].  Requires t < 2^64 - 2^32 p.
// Device form (monty_reduce of the reference, utils.rs:107-127, with the borrow replaced by + p): m = lo(t) p^-1 mod 2^32
// and p^-1 = 2^31 + 2^24 + 1, so m is two shifts and a 3-input add on the ALU pipe instead of an IMAD on the
// multiplier pipe every kernel here is bound by; u = m p has the same low word as t, hence (t - u) / 2^32 =
// hi(t) - hi(u) with no carry chain.  The shift amounts come from constant memory: with literals ptxas folds the
// shifts back into IMAD lo, 0x81000001.  Per reduction: 4 multiplier-pipe cycles instead of 6 (IMAD 2 + IMAD.WIDE 4).
LM_HD uint32_t kb_redc_lazy(uint64_t t) {
#if defined(__CUDA_ARCH__) && !defined(LM_REDC_OLD)
  const uint32_t lo = (uint32_t)t;
#ifdef LM_REDC_SHIFTS
  const uint32_t m = lo + (lo << c_kb.s24) + (lo << c_kb.s31);
#else
  const uint32_t m = lo * 0x81000001u;
#endif
  const uint64_t u = mul_wide(m, c_kb.p);
  return (uint32_t)(t >> 32) - (uint32_t)(u >> 32) + c_kb.p;
#else
  uint32_t m = (uint32_t)t * KB_NEG_MU;
  uint64_t t2 = mad_wide(m, LM_KB_P_OPAQUE, t);
  return (uint32_t)(t2 >> 32);
#endif
}
// a * b * 2^-32 mod p, result in [0, 2p) provided a * b < 2^32 p (e.g. a < 2p, b < p or both < 1.43p).
LM_HD uint32_t kb_mul_lazy(uint32_t a, uint32_t b) { return kb_redc_lazy(mul_wide(a, b)); }
// canonical product
LM_HD uint32_t kb_mul(uint32_t a, uint32_t b) { return kb_canon(kb_mul_lazy(a, b)); }

// value-preserving (mod p) shrink of a 64-bit accumulator to < 2^57:  hi * (2^32 mod p) + lo
LM_HD uint64_t kb_fold(uint64_t acc) { return mad_wide((uint32_t)(acc >> 32), KB_R1, (uint64_t)(uint32_t)acc); }

// Accumulator for sum_j a_j * c_j with a_j < p + 2^9, c_j < p.  Every product is < 0.2462 * 2^64, so four of them
// fit on top of a folded (< 2^57) accumulator; `mac<K>()` folds when the compile-time term index says so.
struct KbDot {
  uint64_t acc;
  LM_HD explicit KbDot(uint64_t init = 0) : acc(init) {}
  template <int TERM_INDEX>
  LM_HD void mac(uint32_t a, uint32_t c) {
    if (TERM_INDEX > 0 && TERM_INDEX % 4 == 0) acc = kb_fold(acc);
    acc = mad_wide(a, c, acc);
  }
  // sum * 2^-32 mod p in [0, p + 2^25)
  LM_HD uint32_t finish_lazy() const { return kb_redc_lazy(kb_fold(acc)); }
  LM_HD uint32_t finish() const { return kb_canon(finish_lazy()); }
};

// 96-bit accumulator for long dot products (the Poseidon1 partial section: 16-20 terms): no intermediate folds, one
// IMAD.WIDE per term plus a three-word carry chain; finish folds the three words with 2^32 and 2^64 mod p.
struct KbAcc96 {
  uint32_t lo, hi, top;
  LM_HD explicit KbAcc96(uint64_t init = 0) : lo((uint32_t)init), hi((uint32_t)(init >> 32)), top(0) {}
  LM_HD void mac(uint32_t a, uint32_t c) {
    const uint64_t p = mul_wide(a, c);
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
        : "+r"(lo), "+r"(hi), "+r"(top)
        : "r"((uint32_t)p), "r"((uint32_t)(p >> 32)));
#else
    const uint64_t cur = ((uint64_t)hi << 32) | lo, sum = cur + p;
    top += sum < cur;
    lo = (uint32_t)sum, hi = (uint32_t)(sum >> 32);
#endif
  }
  // (top 2^64 + hi 2^32 + lo) * 2^-32 mod p in [0, p + 2^26), provided top < 2^6
  LM_HD uint32_t finish_lazy() const {
    return kb_redc_lazy(mad_wide(top, KB_R2, mad_wide(hi, KB_R1, (uint64_t)lo)));
  }
};

// ---- quintic extension EF = F[X]/(X^5 + X^2 - 1), AoS [c0..c4] -------------------------------------------
struct Ef {
  uint32_t c[5];
};
LM_HD Ef ef_zero() { return Ef{{0, 0, 0, 0, 0}}; }
LM_HD Ef ef_from_base(uint32_t a) { return Ef{{a, 0, 0, 0, 0}}; }
LM_HD Ef ef_add(const Ef& a, const Ef& b) {
  Ef r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.c[i] = kb_add(a.c[i], b.c[i]);
  return r;
}
LM_HD Ef ef_sub(const Ef& a, const Ef& b) {
  Ef r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.c[i] = kb_sub(a.c[i], b.c[i]);
  return r;
}
LM_HD Ef ef_mul_base(const Ef& a, uint32_t b) {
  Ef r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.c[i] = kb_mul(a.c[i], b);
  return r;
}
LM_HD Ef ef_add_base(Ef a, uint32_t b) {
  a.c[0] = kb_add(a.c[0], b);
  return a;
}
// Product mod X^5 = 1 - X^2.  Each output coefficient is one delayed-reduction dot product of length 5
// against pre-combined b terms — the same regrouping as the reference's quintic_mul (extension.rs:531-548),
// with negated terms expressed as (p - x) so all accumulations are unsigned.
LM_HD Ef ef_mul(const Ef& a, const Ef& b) {
  const uint32_t b0 = b.c[0], b1 = b.c[1], b2 = b.c[2], b3 = b.c[3], b4 = b.c[4];
  const uint32_t b0m3 = kb_sub(b0, b3), b1m4 = kb_sub(b1, b4), b4m2 = kb_sub(b4, b2);
  const uint32_t b3m14 = kb_sub(b3, b1m4);
  const uint32_t rows[5][5] = {{b0, b4, b3, b2, b1m4},
                               {b1, b0, b4, b3, b2},
                               {b2, b1m4, b0m3, b4m2, b3m14},
                               {b3, b2, b1m4, b0m3, b4m2},
                               {b4, b3, b2, b1m4, b0m3}};
  Ef r;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    KbDot d;
    d.mac<0>(a.c[0], rows[i][0]);
    d.mac<1>(a.c[1], rows[i][1]);
    d.mac<2>(a.c[2], rows[i][2]);
    d.mac<3>(a.c[3], rows[i][3]);
    d.mac<4>(a.c[4], rows[i][4]);
    r.c[i] = d.finish();
  }
  return r;
}
LM_HD Ef ef_sqr(const Ef& a) { return ef_mul(a, a); }

// the nine distinct entries of ef_mul's row matrix for a fixed right operand
struct EfRows {
  uint32_t b0, b1, b2, b3, b4, b0m3, b1m4, b4m2, b3m14;
};
LM_HD EfRows ef_rows(const Ef& b) {
  EfRows r;
  r.b0 = b.c[0], r.b1 = b.c[1], r.b2 = b.c[2], r.b3 = b.c[3], r.b4 = b.c[4];
  r.b0m3 = kb_sub(r.b0, r.b3), r.b1m4 = kb_sub(r.b1, r.b4), r.b4m2 = kb_sub(r.b4, r.b2);
  r.b3m14 = kb_sub(r.b3, r.b1m4);
  return r;
}
// a * b + c * d with one reduction per coefficient (ten delayed products)
LM_HD Ef ef_mul2_add(const Ef& a, const Ef& b, const Ef& c, const Ef& d) {
  const EfRows x = ef_rows(b), y = ef_rows(d);
  const uint32_t rx[5][5] = {{x.b0, x.b4, x.b3, x.b2, x.b1m4},
                             {x.b1, x.b0, x.b4, x.b3, x.b2},
                             {x.b2, x.b1m4, x.b0m3, x.b4m2, x.b3m14},
                             {x.b3, x.b2, x.b1m4, x.b0m3, x.b4m2},
                             {x.b4, x.b3, x.b2, x.b1m4, x.b0m3}};
  const uint32_t ry[5][5] = {{y.b0, y.b4, y.b3, y.b2, y.b1m4},
                             {y.b1, y.b0, y.b4, y.b3, y.b2},
                             {y.b2, y.b1m4, y.b0m3, y.b4m2, y.b3m14},
                             {y.b3, y.b2, y.b1m4, y.b0m3, y.b4m2},
                             {y.b4, y.b3, y.b2, y.b1m4, y.b0m3}};
  Ef r;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    KbDot t;
    t.mac<0>(a.c[0], rx[i][0]);
    t.mac<1>(a.c[1], rx[i][1]);
    t.mac<2>(a.c[2], rx[i][2]);
    t.mac<3>(a.c[3], rx[i][3]);
    t.mac<4>(a.c[4], rx[i][4]);
    t.mac<5>(c.c[0], ry[i][0]);
    t.mac<6>(c.c[1], ry[i][1]);
    t.mac<7>(c.c[2], ry[i][2]);
    t.mac<8>(c.c[3], ry[i][3]);
    t.mac<9>(c.c[4], ry[i][4]);
    r.c[i] = t.finish();
  }
  return r;
}

}  // namespace lm

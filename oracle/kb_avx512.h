/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 * AVX-512 KoalaBear lanes shared by poseidon1_avx512.c and dft.c: 16 Montgomery residues per vector, multiplied as two
 * sets of 8 x 64-bit lanes with vpmuludq (the shape of the reference's PackedMontyField31AVX512 products,
 * crates/backend/koala-bear/src/monty_31/x86_64_avx512/packing.rs).  Every function carries its own target attribute;
 * callers check lm_or_have_avx512() first. */
#ifndef LM_ORACLE_KB_AVX512_H
#define LM_ORACLE_KB_AVX512_H
#include <immintrin.h>
#include "kb.h"

#define TGT __attribute__((target("avx512f,avx512dq"), always_inline)) static inline
#define TGT_FN __attribute__((target("avx512f,avx512dq")))

/* 8 x u64 lanes, each < 2^32 p: x * 2^-32 mod p in [0, p) (upper halves zero) */
TGT __m512i red64(__m512i x) {
  const __m512i mu = _mm512_set1_epi64(KB_MU), p = _mm512_set1_epi64(KB_P);
  __m512i t = _mm512_mul_epu32(x, mu);      /* low 32 bits: t = x * mu mod 2^32 */
  __m512i u = _mm512_mul_epu32(t, p);       /* u = t * p, same low word as x */
  __m512i d = _mm512_sub_epi64(x, u);       /* multiple of 2^32, possibly negative */
  __m512i hi = _mm512_srai_epi64(d, 32);    /* in (-p, p) */
  __mmask8 neg = _mm512_movepi64_mask(d);
  return _mm512_mask_add_epi64(hi, neg, hi, p);
}
/* 8 x u64 lanes, each an exact integer < 2^42: x mod p in [0, p) */
TGT __m512i mod42(__m512i x) {
  const __m512i p = _mm512_set1_epi64(KB_P);
  __m512i q = _mm512_srli_epi64(x, 31);
  __m512i r = _mm512_sub_epi64(x, _mm512_mul_epu32(q, p));    /* < 17 p */
  q = _mm512_srli_epi64(r, 31);
  r = _mm512_sub_epi64(r, _mm512_mul_epu32(q, p));            /* < 2 p */
  return _mm512_min_epu64(r, _mm512_sub_epi64(r, p));
}
TGT __m512i odd(__m512i x) { return _mm512_srli_epi64(x, 32); }
TGT __m512i join(__m512i e, __m512i o) { return _mm512_or_si512(e, _mm512_slli_epi64(o, 32)); }
/* 16 lanes: canonical sum / product */
TGT __m512i add16(__m512i a, __m512i b) {
  const __m512i p = _mm512_set1_epi32((int)KB_P);
  __m512i s = _mm512_add_epi32(a, b);
  return _mm512_min_epu32(s, _mm512_sub_epi32(s, p));
}
TGT __m512i mul16(__m512i a, __m512i b) {
  __m512i e = red64(_mm512_mul_epu32(a, b));
  __m512i o = red64(_mm512_mul_epu32(odd(a), odd(b)));
  return join(e, o);
}
TGT __m512i cube16(__m512i a) { return mul16(mul16(a, a), a); }
TGT __m512i sub16(__m512i a, __m512i b) {
  const __m512i p = _mm512_set1_epi32((int)KB_P);
  __m512i d = _mm512_sub_epi32(a, b);
  return _mm512_min_epu32(d, _mm512_add_epi32(d, p));
}
#endif

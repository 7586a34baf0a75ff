/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * 16 Poseidon1 compressions at a time with AVX-512, one state per 32-bit lane ("vertical" packing) — the CPU
 * baseline of bench.py.  The reference hashes Merkle leaves and tree layers the same way: PackedKoalaBearAVX512 holds
 * 16 field elements, first_digest_layer / compress_layer walk 16 rows / pairs per step
 * (crates/whir/src/merkle.rs:215-288, crates/backend/symetric/src/merkle.rs:50-90) through permute_simd
 * (crates/backend/koala-bear/src/poseidon1_koalabear_16.rs:934-1016).  This file restates that shape with the oracle's
 * own constants (poseidon1.c::p1_init): it is checked lane for lane against the scalar oracle, which the KAT pins.
 *
 * Arithmetic: a lane vector is split into its even and odd 32-bit lanes, each handled as 8 x 64-bit lanes with
 * vpmuludq.  red64 is monty_reduce (monty_31/utils.rs:107-127) on 64-bit lanes; the circulant MDS uses the small
 * integer coefficients directly (sums < 2^42) followed by an exact two-step quotient estimate (p ~ 2^31).
 *
 * Compiled with per-function target attributes; callers must check lm_or_have_avx512() first.
 */
#include <stdlib.h>
#include <string.h>
#include "kb.h"
#include "oracle.h"
#include "poseidon1_consts.h"

#include "kb_avx512.h"

int lm_or_have_avx512(void) {
  static int cached = -1;
  if (cached < 0) {
    const char *off = getenv("LM_ORACLE_NO_AVX512");
    cached = (!off || !*off) && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq");
  }
  return cached;
}

static const uint32_t MDS_COL[16] = {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1};

/* out = MDS * s with the circulant of small integers; s canonical Montgomery residues */
TGT void mds16(__m512i s[16]) {
  __m512i so[16], out[16];
  for (int j = 0; j < 16; j++) so[j] = odd(s[j]);
  for (int i = 0; i < 16; i++) {
    __m512i ae = _mm512_setzero_si512(), ao = _mm512_setzero_si512();
    for (int j = 0; j < 16; j++) {
      const __m512i c = _mm512_set1_epi64(MDS_COL[(16 + i - j) & 15]);
      ae = _mm512_add_epi64(ae, _mm512_mul_epu32(s[j], c));
      ao = _mm512_add_epi64(ao, _mm512_mul_epu32(so[j], c));
    }
    out[i] = join(mod42(ae), mod42(ao));
  }
  memcpy(s, out, sizeof(out));
}

/* sum_j s[j] * row[j] with one reduction per 2 products (2 p^2 < 2^32 p, the bound monty_reduce needs) */
TGT __m512i dot16(const __m512i s[16], const __m512i so[16], const kb_t row[16]) {
  __m512i re = _mm512_setzero_si512(), ro = _mm512_setzero_si512();
  for (int g = 0; g < 8; g++) {
    __m512i ae = _mm512_setzero_si512(), ao = _mm512_setzero_si512();
    for (int j = 2 * g; j < 2 * g + 2; j++) {
      const __m512i c = _mm512_set1_epi64(row[j]);
      ae = _mm512_add_epi64(ae, _mm512_mul_epu32(s[j], c));
      ao = _mm512_add_epi64(ao, _mm512_mul_epu32(so[j], c));
    }
    re = _mm512_add_epi64(re, red64(ae));
    ro = _mm512_add_epi64(ro, red64(ao));
  }
  /* sums of eight residues < 8 p < 2^34: bring back to [0, p) */
  return join(mod42(re), mod42(ro));
}

TGT void full_round16(__m512i s[16], const kb_t rc[16]) {
  for (int i = 0; i < 16; i++) s[i] = cube16(add16(s[i], _mm512_set1_epi32((int)rc[i])));
  mds16(s);
}

/* permute_generic (:873-912) on 16 states, then the feed-forward of compress_in_place (:1020-1030) */
TGT_FN static void compress16(__m512i s[16]) {
  const p1_consts_t *C = lm_or_p1_consts();
  __m512i in[16];
  memcpy(in, s, sizeof(in));
  for (int r = 0; r < P1_RF_HALF; r++) full_round16(s, C->rc[r]);
  {
    __m512i t[16], to[16], out[16];
    for (int i = 0; i < 16; i++) t[i] = add16(s[i], _mm512_set1_epi32((int)C->first_rc[i])), to[i] = odd(t[i]);
    for (int i = 0; i < 16; i++) out[i] = dot16(t, to, C->m_i[i]);
    memcpy(s, out, sizeof(out));
  }
  for (int r = 0; r < P1_RP; r++) {
    __m512i s0 = cube16(s[0]);
    if (r < P1_RP - 1) s0 = add16(s0, _mm512_set1_epi32((int)C->scalar_rc[r]));
    s[0] = s0;
    __m512i so[16];
    for (int j = 0; j < 16; j++) so[j] = odd(s[j]);
    const __m512i dot = dot16(s, so, C->first_row[r]);
    for (int i = 1; i < 16; i++) s[i] = add16(s[i], mul16(s0, _mm512_set1_epi32((int)C->v[r][i - 1])));
    s[0] = dot;
  }
  for (int r = 0; r < P1_RF_HALF; r++) full_round16(s, C->rc[P1_RF_HALF + P1_RP + r]);
  for (int i = 0; i < 16; i++) s[i] = add16(s[i], in[i]);
}

/* 16 states stored one after the other (16 words each): compress each in place */
TGT_FN void lm_or_poseidon1_compress_x16(uint32_t *states) {
  const __m512i idx = _mm512_mullo_epi32(_mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0), _mm512_set1_epi32(16));
  __m512i s[16];
  for (int k = 0; k < 16; k++) s[k] = _mm512_i32gather_epi32(idx, states + k, 4);
  compress16(s);
  for (int k = 0; k < 16; k++) _mm512_i32scatter_epi32(states + k, idx, s[k], 4);
}

/* leaf digests of 16 consecutive rows starting at `rows`: the row sponge of merkle.c::leaf_digest, restricted to the
 * chunk-aligned case (effective_width and stored_width multiples of 8); the caller falls back to the scalar path
 * otherwise.  zero_state != NULL: start from the pre-absorbed zero suffix (>= 2 trailing zero chunks). */
TGT_FN void lm_or_leaf_digest_x16(const uint32_t *rows, uint32_t stored_width, uint32_t full_width, uint32_t effective_width,
                                  const uint32_t *zero_state, uint32_t *digests) {
  const __m512i lane = _mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  const __m512i ridx = _mm512_mullo_epi32(lane, _mm512_set1_epi32((int)stored_width));
  __m512i s[16];
  int64_t chunk; /* next rate chunk to absorb */
  const uint32_t lim = zero_state ? effective_width : stored_width;
  if (zero_state) {
    for (int k = 0; k < 16; k++) s[k] = _mm512_set1_epi32((int)zero_state[k]);
    chunk = (int64_t)effective_width / 8 - 1;
  } else {
    /* the first compression takes the last two chunks of the virtual row */
    const int64_t n_chunks = full_width / 8;
    for (int half = 0; half < 2; half++) {
      const int64_t c = n_chunks - 2 + half;
      for (int k = 0; k < 8; k++)
        s[8 * half + k] = (uint64_t)(8 * c + k) < lim ? _mm512_i32gather_epi32(ridx, rows + 8 * c + k, 4) : _mm512_setzero_si512();
    }
    compress16(s);
    chunk = n_chunks - 3;
  }
  for (; chunk >= 0; chunk--) {
    for (int k = 0; k < 8; k++)
      s[8 + k] = (uint64_t)(8 * chunk + k) < lim ? _mm512_i32gather_epi32(ridx, rows + 8 * chunk + k, 4) : _mm512_setzero_si512();
    compress16(s);
  }
  const __m512i didx = _mm512_mullo_epi32(lane, _mm512_set1_epi32(8));
  for (int k = 0; k < 8; k++) _mm512_i32scatter_epi32(digests + k, didx, s[k], 4);
}

/* next[i] = C(prev[2i] || prev[2i+1])[0..8) for 16 consecutive parents */
TGT_FN void lm_or_compress_pairs_x16(const uint32_t *prev, uint32_t *next) {
  const __m512i lane = _mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
  const __m512i sidx = _mm512_mullo_epi32(lane, _mm512_set1_epi32(16));
  __m512i s[16];
  for (int k = 0; k < 16; k++) s[k] = _mm512_i32gather_epi32(sidx, prev + k, 4);
  compress16(s);
  const __m512i didx = _mm512_mullo_epi32(lane, _mm512_set1_epi32(8));
  for (int k = 0; k < 8; k++) _mm512_i32scatter_epi32(next + k, didx, s[k], 4);
}

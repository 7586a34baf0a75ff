/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * Multilinear-polynomial helpers.  MLE index convention: variable x_0 is the
 * most-significant bit of the evaluation index.
 *   reference: crates/backend/poly/src/eq_mle.rs:16-26,85-150  eval_eq / compute_eval_eq (scaled eq table)
 *              crates/backend/poly/src/evals.rs:142-347        eval_multilinear_generic
 *              crates/backend/poly/src/point.rs:51-61          expand_from_univariate (y, y^2, y^4, ...)
 *              crates/backend/poly/src/utils.rs:161-186        fold_multilinear (MSB-first fold)
 */
#include <stdlib.h>
#include <string.h>
#include "ext5.h"
#include "oracle.h"

/* out[b] = scalar * prod_i (b_i ? z_i : 1 - z_i), b big-endian over k variables. */
void lm_or_eq_table(const uint32_t *point /* k x 5 */, uint32_t k, const uint32_t scalar[5], uint32_t *out /* 2^k x 5 */) {
  ef_t *o = (ef_t *)out;
  memcpy(&o[0], scalar, sizeof(ef_t));
  uint64_t len = 1;
  for (uint32_t i = 0; i < k; i++) {
    ef_t z;
    memcpy(&z, point + 5 * i, sizeof(z));
    /* new variable becomes the least-significant bit so far */
    for (int64_t b = (int64_t)len - 1; b >= 0; b--) {
      ef_t hi = ef_mul(o[b], z);
      o[2 * b + 1] = hi;
      o[2 * b] = ef_sub(o[b], hi);
    }
    len <<= 1;
  }
}

void lm_or_expand_from_univariate(const uint32_t y[5], uint32_t n, uint32_t *out /* n x 5 */) {
  ef_t cur;
  memcpy(&cur, y, sizeof(cur));
  for (uint32_t i = 0; i < n; i++) {
    memcpy(out + 5 * i, &cur, sizeof(cur));
    cur = ef_sqr(cur);
  }
}

/* Evaluate the MLE of 2^n evaluations (dim = 1: base field, dim = 5: extension) at an EF point. */
void lm_or_mle_eval(const uint32_t *evals, uint32_t n, uint32_t dim, const uint32_t *point /* n x 5 */, uint32_t out[5]) {
  uint32_t lo_vars = n / 2, hi_vars = n - lo_vars; /* index = (hi bits | lo bits) */
  uint64_t n_lo = (uint64_t)1 << lo_vars, n_hi = (uint64_t)1 << hi_vars;
  ef_t one = ef_one();
  ef_t *left = (ef_t *)malloc(n_lo * sizeof(ef_t));
  ef_t *right = (ef_t *)malloc(n_hi * sizeof(ef_t));
  lm_or_eq_table(point + 5 * hi_vars, lo_vars, one.c, (uint32_t *)left);
  lm_or_eq_table(point, hi_vars, one.c, (uint32_t *)right);
  ef_t total = ef_zero();
#pragma omp parallel
  {
    ef_t acc = ef_zero();
#pragma omp for schedule(static) nowait
    for (uint64_t hi = 0; hi < n_hi; hi++) {
      ef_t inner = ef_zero();
      const uint32_t *chunk = evals + hi * n_lo * dim;
      if (dim == 1) {
        for (uint64_t lo = 0; lo < n_lo; lo++) inner = ef_add(inner, ef_mul_base(left[lo], chunk[lo]));
      } else {
        for (uint64_t lo = 0; lo < n_lo; lo++) {
          ef_t e;
          memcpy(&e, chunk + 5 * lo, sizeof(e));
          inner = ef_add(inner, ef_mul(left[lo], e));
        }
      }
      acc = ef_add(acc, ef_mul(inner, right[hi]));
    }
#pragma omp critical
    total = ef_add(total, acc);
  }
  memcpy(out, &total, sizeof(total));
  free(left);
  free(right);
}

/* poly/src/utils.rs:161-186: out[i] = in[i] + r * (in[i + N/2] - in[i]); EF result. */
void lm_or_fold_msb(const uint32_t *in, uint64_t n_in, uint32_t dim, const uint32_t r[5], uint32_t *out) {
  ef_t rr;
  memcpy(&rr, r, sizeof(rr));
  uint64_t half = n_in / 2;
  ef_t *o = (ef_t *)out;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < half; i++) {
    if (dim == 1) {
      kb_t a = in[i], b = in[i + half];
      o[i] = ef_add_base(ef_mul_base(rr, kb_sub(b, a)), a);
    } else {
      ef_t a, b;
      memcpy(&a, in + 5 * i, sizeof(a));
      memcpy(&b, in + 5 * (i + half), sizeof(b));
      o[i] = ef_add(a, ef_mul(rr, ef_sub(b, a)));
    }
  }
}

/* Thin wrappers so Python can exercise the field code directly. */
void lm_or_ef_mul(const uint32_t a[5], const uint32_t b[5], uint32_t out[5]) {
  ef_t x, y;
  memcpy(&x, a, sizeof(x));
  memcpy(&y, b, sizeof(y));
  ef_t z = ef_mul(x, y);
  memcpy(out, &z, sizeof(z));
}
void lm_or_ef_inv(const uint32_t a[5], uint32_t out[5]) {
  ef_t x;
  memcpy(&x, a, sizeof(x));
  ef_t z = ef_inv(x);
  memcpy(out, &z, sizeof(z));
}
uint32_t lm_or_kb_mul(uint32_t a, uint32_t b) { return kb_mul(a, b); }
uint32_t lm_or_kb_from_u32(uint32_t a) { return kb_from_u32(a); }
uint32_t lm_or_kb_to_u32(uint32_t a) { return kb_to_u32(a); }
uint32_t lm_or_kb_inv(uint32_t a) { return kb_inv(a); }
uint32_t lm_or_kb_two_adic_generator(uint32_t bits) { return kb_two_adic_generator(bits); }

/* OpenMP thread count of the oracle (bench.py's CPU arms: torchrun exports OMP_NUM_THREADS=1 to its workers) */
#include <omp.h>
void lm_or_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}
int lm_or_max_threads(void) { return omp_get_max_threads(); }

/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * WHIR-open side: statement weights, product sumcheck rounds, STIR equality updates.
 *   reference: crates/whir/src/open.rs:518-584              combine_statement
 *              crates/backend/poly/src/eq_mle.rs:16-83      eval_eq_scaled / compute_sparse_eval_eq
 *              crates/backend/poly/src/next_mle.rs:35-58    matrix_next_mle_folded
 *              crates/whir/src/open.rs:337-382              add_new_equality / add_new_base_equality
 *              crates/backend/poly/src/eq_mle.rs:372-430    compute_eval_eq_base_packed_batched
 *              crates/backend/sumcheck/src/product_computation.rs:127-170,307-315  c0 / c2 of a round
 *              crates/backend/sumcheck/src/product_computation.rs:242-304          fold + next round
 *              crates/backend/poly/src/evals.rs:44-56       evals_to_coeffs
 */
#include <stdlib.h>
#include <string.h>
#include "ext5.h"
#include "oracle.h"

/* weights[(selector << m) + x] += scalar * eq(point, x), x < 2^m  (eq_mle.rs:40-50 with INITIALIZED = true) */
void lm_or_weights_add_eq(uint32_t *weights, uint64_t selector, const uint32_t *point, uint32_t m,
                          const uint32_t scalar[5]) {
  uint64_t n = (uint64_t)1 << m;
  ef_t *tab = (ef_t *)malloc(n * sizeof(ef_t));
  lm_or_eq_table(point, m, scalar, (uint32_t *)tab);
  ef_t *w = (ef_t *)weights + (selector << m);
  for (uint64_t x = 0; x < n; x++) w[x] = ef_add(w[x], tab[x]);
  free(tab);
}

/* next_mle.rs:35-58: res[y] = next_mle(oc, y) for y in {0,1}^n (dense vector, not scaled) */
void lm_or_next_mle_folded(const uint32_t *oc /* n x 5 */, uint32_t n, uint32_t *res /* 2^n x 5 */) {
  ef_t *r = (ef_t *)res;
  uint64_t len = (uint64_t)1 << n;
  for (uint64_t i = 0; i < len; i++) r[i] = ef_zero();
  ef_t one = ef_one();
  for (uint32_t k = 0; k < n; k++) {
    ef_t z;
    memcpy(&z, oc + 5 * (n - k - 1), sizeof(z));
    ef_t prod = ef_sub(one, z);
    for (uint32_t j = n - k; j < n; j++) {
      ef_t t;
      memcpy(&t, oc + 5 * j, sizeof(t));
      prod = ef_mul(prod, t);
    }
    uint32_t pre = n - k - 1;
    ef_t *eq = (ef_t *)malloc(((uint64_t)1 << pre) * sizeof(ef_t));
    lm_or_eq_table(oc, pre, prod.c, (uint32_t *)eq);
    for (uint64_t b = 0; b < ((uint64_t)1 << pre); b++) {
      uint64_t i = (b << (k + 1)) + ((uint64_t)1 << k);
      r[i] = ef_add(r[i], eq[b]);
    }
    free(eq);
  }
  ef_t all = one;
  for (uint32_t j = 0; j < n; j++) {
    ef_t t;
    memcpy(&t, oc + 5 * j, sizeof(t));
    all = ef_mul(all, t);
  }
  r[len - 1] = ef_add(r[len - 1], all);
}

/* weights[(selector << m) + x] += scalar * next_mle(point, x) */
void lm_or_weights_add_next(uint32_t *weights, uint64_t selector, const uint32_t *point, uint32_t m,
                            const uint32_t scalar[5]) {
  uint64_t n = (uint64_t)1 << m;
  ef_t *tab = (ef_t *)malloc(n * sizeof(ef_t));
  lm_or_next_mle_folded(point, m, (uint32_t *)tab);
  ef_t s;
  memcpy(&s, scalar, sizeof(s));
  ef_t *w = (ef_t *)weights + (selector << m);
  for (uint64_t x = 0; x < n; x++) w[x] = ef_add(w[x], ef_mul(tab[x], s));
  free(tab);
}

/* add_new_base_equality (open.rs:360-382): weights[x] += sum_q scalars[q] * eq(points[q], x), base-field points */
void lm_or_weights_add_base_eq(uint32_t *weights, uint32_t m, const uint32_t *points /* q x m */, uint32_t n_q,
                               const uint32_t *scalars /* q x 5 */) {
  uint64_t n = (uint64_t)1 << m;
  kb_t *tab = (kb_t *)malloc(n * sizeof(kb_t));
  ef_t *w = (ef_t *)weights;
  for (uint32_t q = 0; q < n_q; q++) {
    tab[0] = KB_ONE;
    uint64_t len = 1;
    for (uint32_t i = 0; i < m; i++) {
      kb_t z = points[(uint64_t)q * m + i];
      for (int64_t b = (int64_t)len - 1; b >= 0; b--) {
        kb_t hi = kb_mul(tab[b], z);
        tab[2 * b + 1] = hi;
        tab[2 * b] = kb_sub(tab[b], hi);
      }
      len <<= 1;
    }
    ef_t s;
    memcpy(&s, scalars + 5 * q, sizeof(s));
#pragma omp parallel for schedule(static)
    for (uint64_t x = 0; x < n; x++) w[x] = ef_add(w[x], ef_mul_base(s, tab[x]));
  }
  free(tab);
}

/* One product-sumcheck round on p (dim 1 or 5) and w (EF), both of n entries (product_computation.rs:127-170):
 * c0 = sum_{i<n/2} w[i] p[i],  c2 = sum (w[i+n/2] - w[i]) (p[i+n/2] - p[i]) */
void lm_or_prod_round(const uint32_t *p, uint32_t dim, const uint32_t *w, uint64_t n, uint32_t c0[5], uint32_t c2[5]) {
  uint64_t half = n / 2;
  const ef_t *we = (const ef_t *)w;
  ef_t a0 = ef_zero(), a2 = ef_zero();
#pragma omp parallel
  {
    ef_t l0 = ef_zero(), l2 = ef_zero();
#pragma omp for schedule(static) nowait
    for (uint64_t i = 0; i < half; i++) {
      ef_t dw = ef_sub(we[i + half], we[i]);
      if (dim == 1) {
        l0 = ef_add(l0, ef_mul_base(we[i], p[i]));
        l2 = ef_add(l2, ef_mul_base(dw, kb_sub(p[i + half], p[i])));
      } else {
        ef_t x0, x1;
        memcpy(&x0, p + 5 * i, sizeof(x0));
        memcpy(&x1, p + 5 * (i + half), sizeof(x1));
        l0 = ef_add(l0, ef_mul(we[i], x0));
        l2 = ef_add(l2, ef_mul(dw, ef_sub(x1, x0)));
      }
    }
#pragma omp critical
    {
      a0 = ef_add(a0, l0);
      a2 = ef_add(a2, l2);
    }
  }
  memcpy(c0, &a0, sizeof(a0));
  memcpy(c2, &a2, sizeof(a2));
}

/* evals.rs:44-56 on EF data, in place */
void lm_or_evals_to_coeffs(uint32_t *data, uint64_t n) {
  ef_t *d = (ef_t *)data;
  for (uint64_t half = 1; half < n; half <<= 1)
    for (uint64_t i = 0; i < n; i += 2 * half)
      for (uint64_t j = 0; j < half; j++) d[i + j + half] = ef_sub(d[i + j + half], d[i + j]);
  unsigned log_n = 0;
  while (((uint64_t)1 << log_n) < n) log_n++;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t j = 0;
    for (unsigned b = 0; b < log_n; b++)
      if (i & ((uint64_t)1 << b)) j |= (uint64_t)1 << (log_n - 1 - b);
    if (i < j) {
      ef_t t = d[i];
      d[i] = d[j];
      d[j] = t;
    }
  }
}

/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * Quotient GKR for sum_i n_i / d_i (Logup).
 *   reference: crates/sub_protocols/src/quotient_gkr/layers.rs:124-189   sum_quotients_2_by_2 (adjacent pairs)
 *              crates/sub_protocols/src/quotient_gkr/mod.rs:31-141       prove_gkr_quotient / prove_gkr_layer
 *              crates/sub_protocols/src/quotient_gkr/sumcheck_utils.rs:65-79,282-359,491-503  pair_coeffs, rounds, bare poly
 *              crates/utils/src/multilinear.rs:76-98                     finger_print
 * Natural order throughout (the reference's chunk-bit-reversed storage is a CPU SIMD device): a layer of 2^(K+1)
 * entries splits into even (l) and odd (r) halves; the layer above is (nl dr + nr dl, dl dr); the per-layer
 * sumcheck folds the least-significant variable first.  Entries past the active prefix are (0, 1).
 */
#include <stdlib.h>
#include <string.h>
#include "ext5.h"
#include "oracle.h"

/* layers.rs:124-153 on full power-of-two layers (padding (0,1) materialised): nums/dens have n entries of
 * num_dim (1 or 5) / 5 words; outputs n/2 EF entries each */
void lm_or_gkr_layer_up(const uint32_t *nums, uint32_t num_dim, const uint32_t *dens, uint64_t n, uint32_t *out_nums,
                        uint32_t *out_dens) {
  const ef_t *d = (const ef_t *)dens;
  ef_t *on = (ef_t *)out_nums, *od = (ef_t *)out_dens;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n / 2; i++) {
    ef_t d0 = d[2 * i], d1 = d[2 * i + 1];
    if (num_dim == 1) {
      on[i] = ef_add(ef_mul_base(d1, nums[2 * i]), ef_mul_base(d0, nums[2 * i + 1]));
    } else {
      ef_t n0, n1;
      memcpy(&n0, nums + 5 * (2 * i), sizeof(n0));
      memcpy(&n1, nums + 5 * (2 * i + 1), sizeof(n1));
      on[i] = ef_add(ef_mul(d1, n0), ef_mul(d0, n1));
    }
    od[i] = ef_mul(d0, d1);
  }
}

/* One round of the layer sumcheck on the 4 columns (nl, nr, dl, dr), each of n EF entries:
 *   c0 = sum_j eq(eq_point, j) G(lo_j),  c2 = sum_j eq(eq_point, j) G(hi_j - lo_j),
 *   G(nl, nr, dl, dr) = nl dr + nr dl + alpha dl dr,  lo_j = row 2j, hi_j = row 2j+1   (sumcheck_utils.rs:65-79) */
void lm_or_gkr_round(const uint32_t *nl, const uint32_t *nr, const uint32_t *dl, const uint32_t *dr, uint64_t n,
                     const uint32_t *eq_point, const uint32_t alpha[5], uint32_t c0[5], uint32_t c2[5]) {
  uint64_t half = n / 2;
  unsigned lv = 0;
  while (((uint64_t)1 << lv) < half) lv++;
  ef_t *eq = (ef_t *)malloc((half ? half : 1) * sizeof(ef_t));
  ef_t one = ef_one(), al;
  memcpy(&al, alpha, sizeof(al));
  lm_or_eq_table(eq_point, lv, one.c, (uint32_t *)eq);
  const ef_t *NL = (const ef_t *)nl, *NR = (const ef_t *)nr, *DL = (const ef_t *)dl, *DR = (const ef_t *)dr;
  ef_t a0 = ef_zero(), a2 = ef_zero();
  for (uint64_t j = 0; j < half; j++) {
    ef_t g0 = ef_add(ef_add(ef_mul(NL[2 * j], DR[2 * j]), ef_mul(NR[2 * j], DL[2 * j])), ef_mul(al, ef_mul(DL[2 * j], DR[2 * j])));
    ef_t xnl = ef_sub(NL[2 * j + 1], NL[2 * j]), xnr = ef_sub(NR[2 * j + 1], NR[2 * j]);
    ef_t xdl = ef_sub(DL[2 * j + 1], DL[2 * j]), xdr = ef_sub(DR[2 * j + 1], DR[2 * j]);
    ef_t g2 = ef_add(ef_add(ef_mul(xnl, xdr), ef_mul(xnr, xdl)), ef_mul(al, ef_mul(xdl, xdr)));
    a0 = ef_add(a0, ef_mul(eq[j], g0));
    a2 = ef_add(a2, ef_mul(eq[j], g2));
  }
  memcpy(c0, &a0, sizeof(a0));
  memcpy(c2, &a2, sizeof(a2));
  free(eq);
}

/* finger_print (crates/utils/src/multilinear.rs:76-86): c - sum_i alphas[i] * data[i]  with base-field data.
 * data: n_rows x n_data (row-major), alphas: n_data x 5, out: n_rows x 5 */
void lm_or_finger_print(const uint32_t *data, uint64_t n_rows, uint32_t n_data, const uint32_t *alphas,
                        const uint32_t c[5], uint32_t *out) {
  ef_t cc;
  memcpy(&cc, c, sizeof(cc));
  const ef_t *al = (const ef_t *)alphas;
  ef_t *o = (ef_t *)out;
#pragma omp parallel for schedule(static)
  for (uint64_t r = 0; r < n_rows; r++) {
    ef_t s = ef_zero();
    for (uint32_t i = 0; i < n_data; i++) s = ef_add(s, ef_mul_base(al[i], data[r * n_data + i]));
    o[r] = ef_sub(cc, s);
  }
}

/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * Poseidon1-KoalaBear width 16: 4 + 4 full rounds, 20 partial rounds, S-box x^3,
 * circulant MDS with first column MDS_COL.
 *   reference: crates/backend/koala-bear/src/poseidon1_koalabear_16.rs
 *     :11-22   parameters + MDS first column
 *     :699-815 round constants (oracle/poseidon1_rc.inc)
 *     :873-923 permute_generic / full_round  (sparse partial-round form)
 *     :399-480 compute_equivalent_matrices, :482-505 equivalent_round_constants
 *     :1020-1030 compress_in_place = permute(x) + x
 *     :1066-1093 known-answer test
 *
 * Two formulations are kept on purpose:
 *   lm_or_poseidon1_permute_dense  - textbook Poseidon (add full RC vector, S-box, dense MDS)
 *   lm_or_poseidon1_permute        - the reference's optimised form (sparse partial rounds)
 * They must agree with each other and with the KAT (tests/test_oracle_poseidon.py).
 */
#include <string.h>
#include "kb.h"
#include "oracle.h"

#include "poseidon1_consts.h"
#define W P1_W
#define RF_HALF P1_RF_HALF
#define RP P1_RP
#define NR P1_NR

static const uint32_t RC_CANON[NR * W] = {
#include "poseidon1_rc.inc"
};
static const uint32_t MDS_COL[W] = {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1};

static p1_consts_t C;

static void mat_mul(kb_t out[W][W], const kb_t a[W][W], const kb_t b[W][W]) {
  kb_t tmp[W][W];
  for (int i = 0; i < W; i++)
    for (int j = 0; j < W; j++) {
      kb_t s = 0;
      for (int k = 0; k < W; k++) s = kb_add(s, kb_mul(a[i][k], b[k][j]));
      tmp[i][j] = s;
    }
  memcpy(out, tmp, sizeof(tmp));
}
static void mat_vec(kb_t out[W], const kb_t m[W][W], const kb_t v[W]) {
  kb_t tmp[W];
  for (int i = 0; i < W; i++) {
    kb_t s = 0;
    for (int j = 0; j < W; j++) s = kb_add(s, kb_mul(m[i][j], v[j]));
    tmp[i] = s;
  }
  memcpy(out, tmp, sizeof(tmp));
}
/* Gauss-Jordan inverse of the n x n matrix stored with row stride W. */
static void mat_inv(kb_t *inv, const kb_t *m, int n) {
  kb_t a[W][W], b[W][W];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      a[i][j] = m[i * W + j];
      b[i][j] = (i == j) ? KB_ONE : 0;
    }
  for (int col = 0; col < n; col++) {
    int piv = col;
    while (a[piv][col] == 0) piv++;
    if (piv != col)
      for (int j = 0; j < n; j++) {
        kb_t t = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = t;
        t = b[col][j]; b[col][j] = b[piv][j]; b[piv][j] = t;
      }
    kb_t pinv = kb_inv(a[col][col]);
    for (int j = 0; j < n; j++) {
      a[col][j] = kb_mul(a[col][j], pinv);
      b[col][j] = kb_mul(b[col][j], pinv);
    }
    for (int i = 0; i < n; i++) {
      if (i == col || a[i][col] == 0) continue;
      kb_t f = a[i][col];
      for (int j = 0; j < n; j++) {
        a[i][j] = kb_sub(a[i][j], kb_mul(f, a[col][j]));
        b[i][j] = kb_sub(b[i][j], kb_mul(f, b[col][j]));
      }
    }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) inv[i * W + j] = b[i][j];
}

static void p1_init(void) {
  if (C.ready) return;
  for (int r = 0; r < NR; r++)
    for (int i = 0; i < W; i++) C.rc[r][i] = kb_from_u32(RC_CANON[r * W + i]);
  for (int i = 0; i < W; i++)
    for (int j = 0; j < W; j++) C.mds[i][j] = kb_from_u32(MDS_COL[(W + i - j) % W]);

  /* equivalent_round_constants (:482-505): push the partial-round constant vectors
   * backwards through MDS^-1 so only lane 0 keeps a per-round constant. */
  kb_t mds_inv[W][W];
  mat_inv(&mds_inv[0][0], &C.mds[0][0], W);
  kb_t opt[RP];
  kb_t tmp[W];
  memcpy(tmp, C.rc[RF_HALF + RP - 1], sizeof(tmp));
  for (int i = RP - 2; i >= 0; i--) {
    kb_t back[W];
    mat_vec(back, mds_inv, tmp);
    opt[i + 1] = back[0];
    memcpy(tmp, C.rc[RF_HALF + i], sizeof(tmp));
    for (int j = 1; j < W; j++) tmp[j] = kb_add(tmp[j], back[j]);
  }
  memcpy(C.first_rc, tmp, sizeof(tmp));
  for (int r = 0; r < RP - 1; r++) C.scalar_rc[r] = opt[r + 1];

  /* compute_equivalent_matrices (:399-480) */
  kb_t mds_t[W][W], m_mul[W][W], m_i[W][W];
  for (int i = 0; i < W; i++)
    for (int j = 0; j < W; j++) mds_t[i][j] = C.mds[j][i];
  memcpy(m_mul, mds_t, sizeof(m_mul));
  kb_t vs[RP][W], ws[RP][W];
  for (int it = 0; it < RP; it++) {
    kb_t w[W], hat_inv[W][W];
    for (int j = 0; j < W; j++) vs[it][j] = (j < W - 1) ? m_mul[0][j + 1] : 0;
    for (int i = 0; i < W - 1; i++) w[i] = m_mul[i + 1][0];
    /* inverse of bottom-right 15x15 block */
    kb_t sub[W][W];
    memset(sub, 0, sizeof(sub));
    for (int i = 0; i < W - 1; i++)
      for (int j = 0; j < W - 1; j++) sub[i][j] = m_mul[i + 1][j + 1];
    mat_inv(&hat_inv[0][0], &sub[0][0], W - 1);
    for (int i = 0; i < W; i++) {
      kb_t s = 0;
      if (i < W - 1)
        for (int k = 0; k < W - 1; k++) s = kb_add(s, kb_mul(hat_inv[i][k], w[k]));
      ws[it][i] = s;
    }
    memcpy(m_i, m_mul, sizeof(m_i));
    m_i[0][0] = KB_ONE;
    for (int k = 1; k < W; k++) m_i[k][0] = 0, m_i[0][k] = 0;
    mat_mul(m_mul, mds_t, m_i);
  }
  for (int i = 0; i < W; i++)
    for (int j = 0; j < W; j++) C.m_i[i][j] = m_i[j][i];
  for (int r = 0; r < RP; r++) {
    int src = RP - 1 - r; /* collections are reversed into application order */
    memcpy(C.v[r], vs[src], sizeof(C.v[r]));
    C.first_row[r][0] = C.mds[0][0];
    for (int i = 1; i < W; i++) C.first_row[r][i] = ws[src][i - 1];
  }
  C.ready = 1;
}

static inline void mds_dense(kb_t s[W]) {
  kb_t out[W];
  for (int i = 0; i < W; i++) {
    /* small integer constants on Montgomery-form state: sum < 2^42, one reduction */
    uint64_t acc = 0;
    for (int j = 0; j < W; j++) acc += (uint64_t)MDS_COL[(W + i - j) % W] * s[j];
    out[i] = (kb_t)(acc % KB_P);
  }
  memcpy(s, out, sizeof(out));
}

static inline void full_round(kb_t s[W], const kb_t rc[W]) {
  for (int i = 0; i < W; i++) s[i] = kb_cube(kb_add(s[i], rc[i]));
  mds_dense(s);
}

void lm_or_poseidon1_permute_dense(uint32_t s[16]) {
  p1_init();
  for (int r = 0; r < RF_HALF; r++) full_round(s, C.rc[r]);
  for (int r = 0; r < RP; r++) {
    for (int i = 0; i < W; i++) s[i] = kb_add(s[i], C.rc[RF_HALF + r][i]);
    s[0] = kb_cube(s[0]);
    mds_dense(s);
  }
  for (int r = 0; r < RF_HALF; r++) full_round(s, C.rc[RF_HALF + RP + r]);
}

void lm_or_poseidon1_permute(uint32_t s[16]) {
  p1_init();
  for (int r = 0; r < RF_HALF; r++) full_round(s, C.rc[r]);
  for (int i = 0; i < W; i++) s[i] = kb_add(s[i], C.first_rc[i]);
  mat_vec(s, C.m_i, s);
  for (int r = 0; r < RP; r++) {
    kb_t s0 = kb_cube(s[0]);
    if (r < RP - 1) s0 = kb_add(s0, C.scalar_rc[r]);
    s[0] = s0;
    kb_t dot = 0;
    for (int j = 0; j < W; j++) dot = kb_add(dot, kb_mul(s[j], C.first_row[r][j]));
    for (int i = 1; i < W; i++) s[i] = kb_add(s[i], kb_mul(s0, C.v[r][i - 1]));
    s[0] = dot;
  }
  for (int r = 0; r < RF_HALF; r++) full_round(s, C.rc[RF_HALF + RP + r]);
}

/* compress_in_place (:1020): state <- permute(state) + state */
void lm_or_poseidon1_compress(uint32_t s[16]) {
  kb_t in[W];
  memcpy(in, s, sizeof(in));
  lm_or_poseidon1_permute(s);
  for (int i = 0; i < W; i++) s[i] = kb_add(s[i], in[i]);
}

void lm_or_poseidon1_init(void) { p1_init(); }

/* constants in application order, for the Poseidon16 AIR / trace generator (air.c) */
const p1_consts_t *lm_or_p1_consts(void) {
  p1_init();
  return &C;
}

/* Batched entry points for the Python harness: n states, 16 u32 each. */
void lm_or_poseidon1_permute_batch(uint32_t *states, uint64_t n, int dense) {
  p1_init();
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) {
    if (dense)
      lm_or_poseidon1_permute_dense(states + 16 * i);
    else
      lm_or_poseidon1_permute(states + 16 * i);
  }
}
void lm_or_poseidon1_compress_batch(uint32_t *states, uint64_t n) {
  p1_init();
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) lm_or_poseidon1_compress(states + 16 * i);
}

/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).  Poseidon1-KoalaBear-16 constants shared by poseidon1.c and air.c. */
#ifndef LM_ORACLE_P1_CONSTS_H
#define LM_ORACLE_P1_CONSTS_H
#include "kb.h"
#define P1_W 16
#define P1_RF_HALF 4
#define P1_RP 20
#define P1_NR (2 * P1_RF_HALF + P1_RP)
typedef struct {
  kb_t rc[P1_NR][P1_W];         /* Montgomery form */
  kb_t mds[P1_W][P1_W];         /* mds[i][j] = col[(i - j) mod 16] */
  kb_t first_rc[P1_W];          /* sparse form: vector added before m_i */
  kb_t m_i[P1_W][P1_W];         /* dense transition matrix */
  kb_t first_row[P1_RP][P1_W];  /* [mds00, w_hat[0..15)] */
  kb_t v[P1_RP][P1_W];          /* rank-1 column update, v[r][15] = 0 */
  kb_t scalar_rc[P1_RP - 1];    /* added to lane 0 after the S-box of rounds 0..RP-2 */
  int ready;
} p1_consts_t;
const p1_consts_t *lm_or_p1_consts(void);
#endif

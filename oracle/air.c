/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * AIR ("SuperSpartan") sumcheck of the lean_vm execution table.
 *   reference: crates/lean_vm/src/tables/execution/air.rs:42-130   ExecutionTable::eval (13 constraints, degree 5)
 *              crates/lean_vm/src/tables/utils.rs:5-21             eval_virtual_bus_column
 *              crates/backend/air/src/constraint_folder/normal.rs:49-62   alpha-power folding
 *              crates/backend/sumcheck/src/sc_computation.rs:19-26 flat = point[..n_columns], shift = point[n_columns..]
 *              crates/sub_protocols/src/air_sumcheck.rs:225-287,560-634   round evaluations at z = 0, 2, .., d; LSB fold
 *              crates/sub_protocols/src/air_sumcheck.rs:683-694    compute_shifted_columns
 * The reference stores columns bit-reversed inside 2^12 chunks to keep SIMD lanes busy; that layout is invisible
 * at the session interface: round r binds the least-significant remaining variable of the natural row index,
 * and eq_factor's LAST entry belongs to it.  The reference also replaces the all-padding tail by a closed form;
 * the plain sum over the whole hypercube below is the same field element whenever the tail rows are constant.
 */
#include <stdlib.h>
#include <string.h>
#include "ext5.h"
#include "oracle.h"

#define EXEC_N_COLS 20
#define EXEC_N_SHIFT 2
#define EXEC_N_CONSTRAINTS 13
#define EXEC_DEGREE 5
#define LOGUP_PRECOMPILE_DOMAINSEP 1

enum { COL_PC, COL_FP, COL_ADDR_A, COL_ADDR_B, COL_ADDR_C, COL_VAL_A, COL_VAL_B, COL_VAL_C, COL_OP_A, COL_OP_B, COL_OP_C,
       COL_FLAG_A, COL_FLAG_B, COL_FLAG_C, COL_FLAG_C_FP, COL_FLAG_AB_FP, COL_MUL, COL_JUMP, COL_AUX, COL_PRECOMPILE_DATA };

typedef struct {
  const ef_t *alpha_powers; /* >= 13 */
  const ef_t *la;           /* logup_alphas_eq_poly, n_la entries */
  uint32_t n_la;
  ef_t beta;
  ef_t acc;
  int idx;
} folder_t;

static void assert_zero(folder_t *f, ef_t x) {
  f->acc = ef_add(f->acc, ef_mul(f->alpha_powers[f->idx], x));
  f->idx++;
}

static ef_t ef_c(uint32_t canon) { return ef_from_base(kb_from_u32(canon)); }

/* point: 20 flat columns then 2 shift columns (pc, fp at the next row) */
static ef_t exec_eval(const ef_t *pt, const ef_t *alpha_powers, const ef_t *la, uint32_t n_la, ef_t beta) {
  folder_t f = {alpha_powers, la, n_la, beta, ef_zero(), 0};
  const ef_t one = ef_one();
  const ef_t *flat = pt, *shift = pt + EXEC_N_COLS;
  ef_t pc_shift = shift[COL_PC], fp_shift = shift[COL_FP];
  ef_t op_a = flat[COL_OP_A], op_b = flat[COL_OP_B], op_c = flat[COL_OP_C];
  ef_t flag_a = flat[COL_FLAG_A], flag_b = flat[COL_FLAG_B], flag_c = flat[COL_FLAG_C];
  ef_t flag_c_fp = flat[COL_FLAG_C_FP], flag_ab_fp = flat[COL_FLAG_AB_FP];
  ef_t mul = flat[COL_MUL], jump = flat[COL_JUMP], aux = flat[COL_AUX], pdata = flat[COL_PRECOMPILE_DATA];
  ef_t val_a = flat[COL_VAL_A], val_b = flat[COL_VAL_B], val_c = flat[COL_VAL_C];
  ef_t pc = flat[COL_PC], fp = flat[COL_FP];
  ef_t addr_a = flat[COL_ADDR_A], addr_b = flat[COL_ADDR_B], addr_c = flat[COL_ADDR_C];

  ef_t om_a = ef_neg(ef_sub(ef_add(flag_a, flag_ab_fp), one));
  ef_t om_b = ef_neg(ef_sub(ef_add(flag_b, flag_ab_fp), one));
  ef_t om_c = ef_neg(ef_sub(ef_add(flag_c, flag_c_fp), one));

  ef_t fp_op_a = ef_add(fp, op_a), fp_op_b = ef_add(fp, op_b), fp_op_c = ef_add(fp, op_c);
  ef_t nu_a = ef_add(ef_add(ef_mul(flag_a, op_a), ef_mul(om_a, val_a)), ef_mul(flag_ab_fp, fp_op_a));
  ef_t nu_b = ef_add(ef_add(ef_mul(flag_b, op_b), ef_mul(om_b, val_b)), ef_mul(flag_ab_fp, fp_op_b));
  ef_t nu_c = ef_add(ef_add(ef_mul(flag_c, op_c), ef_mul(om_c, val_c)), ef_mul(flag_c_fp, fp_op_c));
  ef_t pc_plus_one = ef_add(pc, one);
  ef_t nu_a_minus_one = ef_sub(nu_a, one);

  ef_t add = ef_sub(ef_mul(aux, ef_c(2)), ef_mul(aux, aux));
  ef_t deref = ef_mul(ef_mul(aux, ef_sub(aux, one)), ef_from_base(kb_inv(kb_from_u32(2))));
  ef_t is_precompile = ef_neg(ef_sub(ef_add(ef_add(ef_add(add, mul), deref), jump), one));

  /* virtual bus column (tables/utils.rs:5-21) */
  {
    ef_t data[4] = {pdata, nu_a, nu_b, nu_c};
    ef_t s = ef_zero();
    for (int i = 0; i < 4; i++) s = ef_add(s, ef_mul(la[i], data[i]));
    s = ef_add(s, ef_mul(la[n_la - 1], ef_c(LOGUP_PRECOMPILE_DOMAINSEP)));
    assert_zero(&f, ef_add(ef_mul(s, beta), is_precompile));
  }
  assert_zero(&f, ef_mul(om_a, ef_sub(addr_a, fp_op_a)));
  assert_zero(&f, ef_mul(om_b, ef_sub(addr_b, fp_op_b)));
  assert_zero(&f, ef_mul(om_c, ef_sub(addr_c, fp_op_c)));
  assert_zero(&f, ef_mul(add, ef_sub(nu_b, ef_add(nu_a, nu_c))));
  assert_zero(&f, ef_mul(mul, ef_sub(nu_b, ef_mul(nu_a, nu_c))));
  assert_zero(&f, ef_mul(deref, ef_sub(addr_b, ef_add(val_a, op_b))));
  assert_zero(&f, ef_mul(deref, ef_sub(val_b, nu_c)));
  ef_t jc = ef_mul(jump, nu_a);
  assert_zero(&f, ef_mul(jc, nu_a_minus_one));
  assert_zero(&f, ef_mul(jc, ef_sub(pc_shift, nu_b)));
  assert_zero(&f, ef_mul(jc, ef_sub(fp_shift, nu_c)));
  ef_t njc = ef_neg(ef_sub(jc, one));
  assert_zero(&f, ef_mul(njc, ef_sub(pc_shift, pc_plus_one)));
  assert_zero(&f, ef_mul(njc, ef_sub(fp_shift, fp)));
  return f.acc;
}

/* Evaluate the folded constraint polynomial at one point (22 EF values). */
void lm_or_air_exec_eval(const uint32_t *point, const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la,
                         const uint32_t beta[5], uint32_t out[5]) {
  ef_t b;
  memcpy(&b, beta, sizeof(b));
  ef_t r = exec_eval((const ef_t *)point, (const ef_t *)alpha_powers, (const ef_t *)la, n_la, b);
  memcpy(out, &r, sizeof(r));
}

/* air_sumcheck.rs:683-694: shifted[i] = col[i + 1], last row repeated */
void lm_or_shift_column(const uint32_t *col, uint64_t n, uint32_t *out) {
  memcpy(out, col + 1, (n - 1) * sizeof(uint32_t));
  out[n - 1] = col[n - 1];
}

/* One round of the execution-table AIR sumcheck over 22 columns of n rows (dim words per entry, column-major:
 * cols[c * n * dim ...]).  eq_point: the len-1 = log2(n)-1 leading entries of eq_factor (those of the variables
 * that stay free), n x 5 words.  out: EXEC_DEGREE evaluations at z = 0, 2, 3, 4, 5 of
 *   sum_j eq(eq_point, j) * C(col(2j) + z (col(2j+1) - col(2j)))                (air_sumcheck.rs:560-634) */
void lm_or_air_exec_round(const uint32_t *cols, uint64_t n, uint32_t dim, const uint32_t *eq_point,
                          const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la, const uint32_t beta[5],
                          uint32_t *out /* 5 x 5 */) {
  const int NC = EXEC_N_COLS + EXEC_N_SHIFT;
  uint64_t half = n / 2;
  unsigned lv = 0;
  while (((uint64_t)1 << lv) < half) lv++;
  ef_t *eq = (ef_t *)malloc((half ? half : 1) * sizeof(ef_t));
  ef_t one = ef_one();
  lm_or_eq_table(eq_point, lv, one.c, (uint32_t *)eq);
  ef_t b;
  memcpy(&b, beta, sizeof(b));
  ef_t acc[EXEC_DEGREE];
  for (int z = 0; z < EXEC_DEGREE; z++) acc[z] = ef_zero();
#pragma omp parallel
  {
    ef_t loc[EXEC_DEGREE];
    for (int z = 0; z < EXEC_DEGREE; z++) loc[z] = ef_zero();
#pragma omp for schedule(static) nowait
    for (uint64_t j = 0; j < half; j++) {
      ef_t lo[22], diff[22], pt[22];
      for (int c = 0; c < NC; c++) {
        const uint32_t *base = cols + (uint64_t)c * n * dim;
        ef_t a, h;
        if (dim == 1) {
          a = ef_from_base(base[2 * j]);
          h = ef_from_base(base[2 * j + 1]);
        } else {
          memcpy(&a, base + 5 * (2 * j), sizeof(a));
          memcpy(&h, base + 5 * (2 * j + 1), sizeof(h));
        }
        lo[c] = a;
        diff[c] = ef_sub(h, a);
        pt[c] = a;
      }
      static const int ZS[EXEC_DEGREE] = {0, 2, 3, 4, 5};
      int cur = 0;
      for (int zi = 0; zi < EXEC_DEGREE; zi++) {
        while (cur < ZS[zi]) {
          for (int c = 0; c < NC; c++) pt[c] = ef_add(pt[c], diff[c]);
          cur++;
        }
        ef_t v = exec_eval(pt, (const ef_t *)alpha_powers, (const ef_t *)la, n_la, b);
        loc[zi] = ef_add(loc[zi], ef_mul(v, eq[j]));
      }
    }
#pragma omp critical
    for (int z = 0; z < EXEC_DEGREE; z++) acc[z] = ef_add(acc[z], loc[z]);
  }
  memcpy(out, acc, sizeof(acc));
  free(eq);
}

/* fold the least-significant variable: out[j] = c[2j] + r (c[2j+1] - c[2j]); EF output */
void lm_or_fold_lsb(const uint32_t *in, uint64_t n_in, uint32_t dim, const uint32_t r[5], uint32_t *out) {
  ef_t rr;
  memcpy(&rr, r, sizeof(rr));
  ef_t *o = (ef_t *)out;
#pragma omp parallel for schedule(static)
  for (uint64_t j = 0; j < n_in / 2; j++) {
    if (dim == 1) {
      kb_t a = in[2 * j], b = in[2 * j + 1];
      o[j] = ef_add_base(ef_mul_base(rr, kb_sub(b, a)), a);
    } else {
      ef_t a, b;
      memcpy(&a, in + 5 * (2 * j), sizeof(a));
      memcpy(&b, in + 5 * (2 * j + 1), sizeof(b));
      o[j] = ef_add(a, ef_mul(rr, ef_sub(b, a)));
    }
  }
}

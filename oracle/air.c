/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * AIR ("SuperSpartan") sumcheck of the lean_vm tables: execution (id 0), extension_op (id 1), poseidon16 (id 2).
 *   reference: crates/lean_vm/src/tables/execution/air.rs:42-130   ExecutionTable::eval (13 constraints, degree 5)
 *              crates/lean_vm/src/tables/extension_op/air.rs:44-163   ExtensionOpPrecompile::eval (29 + 13 columns, degree 6)
 *              crates/lean_vm/src/tables/poseidon_16/mod.rs:294-548   Poseidon16Precompile::eval (109 columns, degree 10)
 *              crates/lean_vm/src/tables/poseidon_16/trace_gen.rs:10-165  fill_trace_poseidon_16
 *              crates/backend/air/src/lib.rs:59-84                 assert_eq / assert_bool (= (1 - x) x) / declare_values
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
#include "poseidon1_consts.h"

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

/* ---- extension_op table (extension_op/air.rs:44-163): 29 flat columns + shifts of the first 13 ------------------ */
#define EXT_N_COLS 29
#define EXT_N_SHIFT 13
#define EXT_DEGREE 6
enum { XC_IS_BE, XC_START, XC_LEN, XC_FLAG_ADD, XC_FLAG_MUL, XC_FLAG_POLY_EQ, XC_IDX_A, XC_IDX_B, XC_COMP, XC_IDX_RES = 13,
       XC_VA = 14, XC_VB = 19, XC_VRES = 24 };

/* quintic_mul on AIR values (extension.rs:531-548 through quintic_mul_air): product in F[X]/(X^5 + X^2 - 1) of two
 * quintuples of (possibly extension-valued) expressions */
static void quintic_mul_air(const ef_t a[5], const ef_t b[5], ef_t out[5]) {
  ef_t d[9];
  for (int k = 0; k < 9; k++) d[k] = ef_zero();
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) d[i + j] = ef_add(d[i + j], ef_mul(a[i], b[j]));
  out[0] = ef_sub(ef_add(d[0], d[5]), d[8]);
  out[1] = ef_add(d[1], d[6]);
  out[2] = ef_add(ef_sub(ef_add(d[2], d[7]), d[5]), d[8]);
  out[3] = ef_add(ef_sub(d[3], d[6]), d[8]);
  out[4] = ef_sub(d[4], d[7]);
}

static void bus_or_declare(folder_t *f, int bus, ef_t flag, const ef_t data[4]) {
  if (!bus) return; /* BUS = false: declare_values only (air/src/lib.rs:82-84) */
  ef_t s = ef_zero();
  for (int i = 0; i < 4; i++) s = ef_add(s, ef_mul(f->la[i], data[i]));
  s = ef_add(s, ef_mul(f->la[f->n_la - 1], ef_c(LOGUP_PRECOMPILE_DOMAINSEP)));
  assert_zero(f, ef_add(ef_mul(s, f->beta), flag));
}

static ef_t bool_check(ef_t x) { return ef_mul(ef_sub(ef_one(), x), x); } /* field.rs:207: (1 - x) x */

static ef_t ext_op_eval(const ef_t *pt, int bus, const ef_t *alpha_powers, const ef_t *la, uint32_t n_la, ef_t beta) {
  folder_t f = {alpha_powers, la, n_la, beta, ef_zero(), 0};
  const ef_t one = ef_one();
  const ef_t *flat = pt, *shift = pt + EXT_N_COLS;
  ef_t is_be = flat[XC_IS_BE], start = flat[XC_START], len = flat[XC_LEN];
  ef_t flag_add = flat[XC_FLAG_ADD], flag_mul = flat[XC_FLAG_MUL], flag_poly_eq = flat[XC_FLAG_POLY_EQ];
  ef_t idx_a = flat[XC_IDX_A], idx_b = flat[XC_IDX_B], idx_r = flat[XC_IDX_RES];
  const ef_t *va = flat + XC_VA, *vb = flat + XC_VB, *vres = flat + XC_VRES, *comp = flat + XC_COMP;
  const ef_t *comp_shift = shift + XC_COMP;
  ef_t start_shift = shift[XC_START];

  ef_t active = ef_add(ef_add(flag_add, flag_mul), flag_poly_eq);
  ef_t activation_flag = ef_mul(start, active);
  ef_t aux = ef_add(ef_add(ef_add(ef_add(ef_mul(is_be, ef_c(4)), ef_mul(flag_add, ef_c(8))), ef_mul(flag_mul, ef_c(16))),
                           ef_mul(flag_poly_eq, ef_c(32))),
                    ef_mul(len, ef_c(64)));
  ef_t data[4] = {aux, idx_a, idx_b, idx_r};
  bus_or_declare(&f, bus, activation_flag, data);

  ef_t is_ee = ef_neg(ef_sub(is_be, one));
  ef_t not_start_shift = ef_neg(ef_sub(start_shift, one));
  ef_t va_f[5], comp_tail[5];
  for (int k = 0; k < 5; k++) {
    va_f[k] = k == 0 ? va[0] : ef_mul(va[k], is_ee);
    comp_tail[k] = ef_mul(comp_shift[k], not_start_shift);
  }
  assert_zero(&f, bool_check(is_be));
  assert_zero(&f, bool_check(start));
  assert_zero(&f, bool_check(flag_add));
  assert_zero(&f, bool_check(flag_mul));
  assert_zero(&f, bool_check(flag_poly_eq));
  for (int k = 0; k < 5; k++) assert_zero(&f, ef_mul(ef_sub(comp[k], ef_add(ef_add(va_f[k], vb[k]), comp_tail[k])), flag_add));
  ef_t prod[5];
  quintic_mul_air(va_f, vb, prod);
  for (int k = 0; k < 5; k++) assert_zero(&f, ef_mul(ef_sub(comp[k], ef_add(prod[k], comp_tail[k])), flag_mul));
  ef_t pe[5], cso[5], per[5];
  for (int k = 0; k < 5; k++) {
    pe[k] = ef_sub(ef_sub(ef_add(prod[k], prod[k]), va_f[k]), vb[k]);
    if (k == 0) pe[k] = ef_add(pe[k], one);
    cso[k] = k == 0 ? ef_add(ef_mul(comp_shift[0], not_start_shift), start_shift) : ef_mul(comp_shift[k], not_start_shift);
  }
  quintic_mul_air(pe, cso, per);
  for (int k = 0; k < 5; k++) assert_zero(&f, ef_mul(ef_sub(comp[k], per[k]), flag_poly_eq));
  for (int k = 0; k < 5; k++) assert_zero(&f, ef_mul(ef_sub(comp[k], vres[k]), start));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(ef_sub(len, shift[XC_LEN]), one)));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(is_be, shift[XC_IS_BE])));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(flag_add, shift[XC_FLAG_ADD])));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(flag_mul, shift[XC_FLAG_MUL])));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(flag_poly_eq, shift[XC_FLAG_POLY_EQ])));
  ef_t a_inc = ef_add(is_be, ef_mul(is_ee, ef_c(5)));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(ef_sub(shift[XC_IDX_A], idx_a), a_inc)));
  assert_zero(&f, ef_mul(not_start_shift, ef_sub(ef_sub(shift[XC_IDX_B], idx_b), ef_c(5))));
  assert_zero(&f, ef_mul(start_shift, ef_sub(len, one)));
  return f.acc;
}

/* ---- poseidon16 table (poseidon_16/mod.rs:294-548): 109 columns, no shifts, degree 10 --------------------------- */
#define P16_N_COLS 109
#define P16_DEGREE 10
enum { PC_FLAG, PC_INDEX_B, PC_INDEX_RES, PC_FLAG_HALF, PC_FLAG_HARD, PC_OFFSET_HARD, PC_EFF_FIRST, PC_EFF_SECOND, PC_FLAG_PERMUTE,
       PC_INPUTS = 9, PC_BEGIN = 25, PC_PARTIAL = 57, PC_END = 77, PC_OUT_LEFT = 93, PC_OUT_RIGHT = 101 };

static void p16_mds(ef_t s[16], const p1_consts_t *C) {
  ef_t out[16];
  for (int i = 0; i < 16; i++) {
    ef_t acc = ef_zero();
    for (int j = 0; j < 16; j++) acc = ef_add(acc, ef_mul_base(s[j], C->mds[i][j]));
    out[i] = acc;
  }
  memcpy(s, out, sizeof(out));
}
static ef_t ef_cube(ef_t x) { return ef_mul(ef_mul(x, x), x); }
static void p16_two_full_rounds(ef_t s[16], const p1_consts_t *C, int r) {
  for (int h = 0; h < 2; h++) {
    for (int i = 0; i < 16; i++) s[i] = ef_cube(ef_add_base(s[i], C->rc[r + h][i]));
    p16_mds(s, C);
  }
}

static ef_t poseidon16_eval(const ef_t *c, int bus, const ef_t *alpha_powers, const ef_t *la, uint32_t n_la, ef_t beta) {
  const p1_consts_t *C = lm_or_p1_consts();
  folder_t f = {alpha_powers, la, n_la, beta, ef_zero(), 0};
  const ef_t one = ef_one();
  ef_t flag_half = c[PC_FLAG_HALF], flag_hard = c[PC_FLAG_HARD], flag_perm = c[PC_FLAG_PERMUTE];
  ef_t pdata = ef_add(ef_add(ef_add(ef_add(one, ef_mul(flag_half, ef_c(4))), ef_mul(flag_hard, ef_c(8))),
                             ef_mul(ef_mul(flag_hard, c[PC_OFFSET_HARD]), ef_c(16))),
                      ef_mul(flag_perm, ef_c(2)));
  ef_t om_hard = ef_sub(one, flag_hard);
  ef_t index_a = ef_sub(c[PC_EFF_SECOND], ef_mul(om_hard, ef_c(4)));
  ef_t data[4] = {pdata, index_a, c[PC_INDEX_B], c[PC_INDEX_RES]};
  bus_or_declare(&f, bus, c[PC_FLAG], data);
  assert_zero(&f, bool_check(c[PC_FLAG]));
  assert_zero(&f, bool_check(flag_half));
  assert_zero(&f, bool_check(flag_hard));
  assert_zero(&f, bool_check(flag_perm));
  assert_zero(&f, ef_mul(flag_perm, ef_add(flag_half, flag_hard)));
  assert_zero(&f, ef_mul(flag_hard, ef_sub(c[PC_OFFSET_HARD], c[PC_EFF_FIRST])));
  assert_zero(&f, ef_mul(om_hard, ef_sub(index_a, c[PC_EFF_FIRST])));

  ef_t s[16];
  memcpy(s, c + PC_INPUTS, sizeof(s));
  for (int r = 0; r < 2; r++) {
    p16_two_full_rounds(s, C, 2 * r);
    for (int i = 0; i < 16; i++) {
      assert_zero(&f, ef_sub(s[i], c[PC_BEGIN + 16 * r + i]));
      s[i] = c[PC_BEGIN + 16 * r + i];
    }
  }
  /* sparse partial rounds */
  for (int i = 0; i < 16; i++) s[i] = ef_add_base(s[i], C->first_rc[i]);
  {
    ef_t out[16];
    for (int i = 0; i < 16; i++) {
      ef_t acc = ef_zero();
      for (int j = 0; j < 16; j++) acc = ef_add(acc, ef_mul_base(s[j], C->m_i[i][j]));
      out[i] = acc;
    }
    memcpy(s, out, sizeof(out));
  }
  for (int r = 0; r < P1_RP; r++) {
    assert_zero(&f, ef_sub(ef_cube(s[0]), c[PC_PARTIAL + r]));
    s[0] = c[PC_PARTIAL + r];
    if (r < P1_RP - 1) s[0] = ef_add_base(s[0], C->scalar_rc[r]);
    ef_t old = s[0], dot = ef_zero();
    for (int j = 0; j < 16; j++) dot = ef_add(dot, ef_mul_base(s[j], C->first_row[r][j]));
    s[0] = dot;
    for (int i = 1; i < 16; i++) s[i] = ef_add(s[i], ef_mul_base(old, C->v[r][i - 1]));
  }
  p16_two_full_rounds(s, C, P1_RF_HALF + P1_RP);
  for (int i = 0; i < 16; i++) {
    assert_zero(&f, ef_sub(s[i], c[PC_END + i]));
    s[i] = c[PC_END + i];
  }
  p16_two_full_rounds(s, C, P1_RF_HALF + P1_RP + 2);
  ef_t not_perm = ef_sub(one, flag_perm);
  ef_t last4 = ef_sub(not_perm, flag_half);
  for (int i = 0; i < 8; i++) {
    ef_t gate = i < 4 ? not_perm : last4;
    assert_zero(&f, ef_mul(gate, ef_sub(ef_add(s[i], c[PC_INPUTS + i]), c[PC_OUT_LEFT + i])));
    assert_zero(&f, ef_mul(flag_perm, ef_sub(s[i], c[PC_OUT_LEFT + i])));
    assert_zero(&f, ef_mul(flag_perm, ef_sub(s[i + 8], c[PC_OUT_RIGHT + i])));
  }
  return f.acc;
}

/* fill_trace_poseidon_16 (trace_gen.rs:10-165): columns 25..109 from the control columns and the inputs.
 * cols: column-major base-field matrix, 109 columns of n rows. */
void lm_or_poseidon16_fill_trace(uint32_t *cols, uint64_t n) {
  const p1_consts_t *C = lm_or_p1_consts();
#pragma omp parallel for schedule(static)
  for (uint64_t row = 0; row < n; row++) {
    kb_t in[16], s[16];
    for (int i = 0; i < 16; i++) s[i] = in[i] = cols[(uint64_t)(PC_INPUTS + i) * n + row];
    kb_t t[16];
#define FULL2(R)                                                            \
  for (int h = 0; h < 2; h++) {                                             \
    for (int i = 0; i < 16; i++) {                                          \
      kb_t x = kb_add(s[i], C->rc[(R) + h][i]);                             \
      s[i] = kb_mul(kb_mul(x, x), x);                                       \
    }                                                                       \
    for (int i = 0; i < 16; i++) {                                          \
      kb_t acc = 0;                                                         \
      for (int j = 0; j < 16; j++) acc = kb_add(acc, kb_mul(C->mds[i][j], s[j])); \
      t[i] = acc;                                                           \
    }                                                                       \
    memcpy(s, t, sizeof(t));                                                \
  }
    for (int r = 0; r < 2; r++) {
      FULL2(2 * r)
      for (int i = 0; i < 16; i++) cols[(uint64_t)(PC_BEGIN + 16 * r + i) * n + row] = s[i];
    }
    for (int i = 0; i < 16; i++) s[i] = kb_add(s[i], C->first_rc[i]);
    for (int i = 0; i < 16; i++) {
      kb_t acc = 0;
      for (int j = 0; j < 16; j++) acc = kb_add(acc, kb_mul(C->m_i[i][j], s[j]));
      t[i] = acc;
    }
    memcpy(s, t, sizeof(t));
    for (int r = 0; r < P1_RP; r++) {
      s[0] = kb_mul(kb_mul(s[0], s[0]), s[0]);
      cols[(uint64_t)(PC_PARTIAL + r) * n + row] = s[0];
      if (r < P1_RP - 1) s[0] = kb_add(s[0], C->scalar_rc[r]);
      kb_t old = s[0], dot = 0;
      for (int j = 0; j < 16; j++) dot = kb_add(dot, kb_mul(s[j], C->first_row[r][j]));
      s[0] = dot;
      for (int i = 1; i < 16; i++) s[i] = kb_add(s[i], kb_mul(old, C->v[r][i - 1]));
    }
    FULL2(P1_RF_HALF + P1_RP)
    for (int i = 0; i < 16; i++) cols[(uint64_t)(PC_END + i) * n + row] = s[i];
    FULL2(P1_RF_HALF + P1_RP + 2)
#undef FULL2
    kb_t fp = cols[(uint64_t)PC_FLAG_PERMUTE * n + row];
    kb_t nfp = kb_sub(KB_ONE, fp);
    for (int i = 0; i < 8; i++) {
      kb_t comp = kb_add(s[i], in[i]);
      cols[(uint64_t)(PC_OUT_LEFT + i) * n + row] = kb_add(kb_mul(nfp, comp), kb_mul(fp, s[i]));
      cols[(uint64_t)(PC_OUT_RIGHT + i) * n + row] = kb_mul(fp, s[i + 8]);
    }
  }
}

/* ---- table dispatch ------------------------------------------------------------------------------------------
 * table = id | LM_OR_AIR_NO_BUS: the BUS = false instantiation (no bus constraint, alpha index starts at the first
 * ordinary constraint), as used by sub_protocols/tests/prove_poseidon_16.rs:39 */
#define LM_OR_AIR_NO_BUS 0x100u
static int air_shape(uint32_t table, int *n_cols, int *n_shift, int *degree) {
  switch (table & 0xff) {
    case 0: *n_cols = EXEC_N_COLS, *n_shift = EXEC_N_SHIFT, *degree = EXEC_DEGREE; return 0;
    case 1: *n_cols = EXT_N_COLS, *n_shift = EXT_N_SHIFT, *degree = EXT_DEGREE; return 0;
    case 2: *n_cols = P16_N_COLS, *n_shift = 0, *degree = P16_DEGREE; return 0;
  }
  return -1;
}
int lm_or_air_shape(uint32_t table, uint32_t out[3]) {
  int a, b, c;
  if (air_shape(table, &a, &b, &c)) return -1;
  out[0] = a, out[1] = b, out[2] = c;
  return 0;
}

static ef_t air_eval(uint32_t table, const ef_t *pt, const ef_t *alpha_powers, const ef_t *la, uint32_t n_la, ef_t beta) {
  int bus = !(table & LM_OR_AIR_NO_BUS);
  switch (table & 0xff) {
    case 0: return exec_eval(pt, alpha_powers, la, n_la, beta);
    case 1: return ext_op_eval(pt, bus, alpha_powers, la, n_la, beta);
    default: return poseidon16_eval(pt, bus, alpha_powers, la, n_la, beta);
  }
}

/* Evaluate the folded constraint polynomial at one point (n_cols + n_shift EF values). */
void lm_or_air_eval(uint32_t table, const uint32_t *point, const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la,
                    const uint32_t beta[5], uint32_t out[5]) {
  ef_t b;
  memcpy(&b, beta, sizeof(b));
  ef_t r = air_eval(table, (const ef_t *)point, (const ef_t *)alpha_powers, (const ef_t *)la, n_la, b);
  memcpy(out, &r, sizeof(r));
}
void lm_or_air_exec_eval(const uint32_t *point, const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la,
                         const uint32_t beta[5], uint32_t out[5]) {
  lm_or_air_eval(0, point, alpha_powers, la, n_la, beta, out);
}

/* air_sumcheck.rs:683-694: shifted[i] = col[i + 1], last row repeated */
void lm_or_shift_column(const uint32_t *col, uint64_t n, uint32_t *out) {
  memcpy(out, col + 1, (n - 1) * sizeof(uint32_t));
  out[n - 1] = col[n - 1];
}

/* One round of a table's AIR sumcheck over its n_cols + n_shift columns of n rows (dim words per entry,
 * column-major: cols[c * n * dim ...]).  eq_point: the log2(n)-1 leading entries of eq_factor (those of the
 * variables that stay free).  out: `degree` evaluations at z = 0, 2, 3, .., degree of
 *   sum_j eq(eq_point, j) * C(col(2j) + z (col(2j+1) - col(2j)))                (air_sumcheck.rs:560-634) */
void lm_or_air_round(uint32_t table, const uint32_t *cols, uint64_t n, uint32_t dim, const uint32_t *eq_point,
                     const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la, const uint32_t beta[5],
                     uint32_t *out /* degree x 5 */) {
  int n_cols, n_shift, DEG;
  if (air_shape(table, &n_cols, &n_shift, &DEG)) return;
  const int NC = n_cols + n_shift;
  uint64_t half = n / 2;
  unsigned lv = 0;
  while (((uint64_t)1 << lv) < half) lv++;
  ef_t *eq = (ef_t *)malloc((half ? half : 1) * sizeof(ef_t));
  ef_t one = ef_one();
  lm_or_eq_table(eq_point, lv, one.c, (uint32_t *)eq);
  ef_t b;
  memcpy(&b, beta, sizeof(b));
  ef_t acc[16];
  for (int z = 0; z < DEG; z++) acc[z] = ef_zero();
#pragma omp parallel
  {
    ef_t loc[16];
    for (int z = 0; z < DEG; z++) loc[z] = ef_zero();
    ef_t diff[128], pt[128];
#pragma omp for schedule(static) nowait
    for (uint64_t j = 0; j < half; j++) {
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
        diff[c] = ef_sub(h, a);
        pt[c] = a;
      }
      int cur = 0;
      for (int zi = 0; zi < DEG; zi++) {
        int z = zi == 0 ? 0 : zi + 1;
        while (cur < z) {
          for (int c = 0; c < NC; c++) pt[c] = ef_add(pt[c], diff[c]);
          cur++;
        }
        ef_t v = air_eval(table, pt, (const ef_t *)alpha_powers, (const ef_t *)la, n_la, b);
        loc[zi] = ef_add(loc[zi], ef_mul(v, eq[j]));
      }
    }
#pragma omp critical
    for (int z = 0; z < DEG; z++) acc[z] = ef_add(acc[z], loc[z]);
  }
  memcpy(out, acc, DEG * sizeof(ef_t));
  free(eq);
}
void lm_or_air_exec_round(const uint32_t *cols, uint64_t n, uint32_t dim, const uint32_t *eq_point,
                          const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la, const uint32_t beta[5],
                          uint32_t *out /* 5 x 5 */) {
  lm_or_air_round(0, cols, n, dim, eq_point, alpha_powers, la, n_la, beta, out);
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

/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * Quintic extension EF = F[X]/(X^5 + X^2 - 1), element = 5 consecutive
 * Montgomery-form u32 (coefficient of X^0 first).
 *   reference: crates/backend/koala-bear/src/quintic_extension/extension.rs:26-36 (layout)
 *              .../extension.rs:531-548 (quintic_mul), mod.rs:60-92 (add/sub/base-mul)
 * All results are canonical, so the schoolbook product + reduction below is
 * bit-identical to the reference's 5-dot-product formulation.
 */
#ifndef LM_ORACLE_EXT5_H
#define LM_ORACLE_EXT5_H
#include "kb.h"

typedef struct {
  kb_t c[5];
} ef_t;

static inline ef_t ef_zero(void) {
  ef_t r = {{0, 0, 0, 0, 0}};
  return r;
}
static inline ef_t ef_one(void) {
  ef_t r = {{KB_ONE, 0, 0, 0, 0}};
  return r;
}
static inline ef_t ef_from_base(kb_t a) {
  ef_t r = {{a, 0, 0, 0, 0}};
  return r;
}
static inline ef_t ef_add(ef_t a, ef_t b) {
  ef_t r;
  for (int i = 0; i < 5; i++) r.c[i] = kb_add(a.c[i], b.c[i]);
  return r;
}
static inline ef_t ef_sub(ef_t a, ef_t b) {
  ef_t r;
  for (int i = 0; i < 5; i++) r.c[i] = kb_sub(a.c[i], b.c[i]);
  return r;
}
static inline ef_t ef_neg(ef_t a) {
  ef_t r;
  for (int i = 0; i < 5; i++) r.c[i] = kb_neg(a.c[i]);
  return r;
}
static inline ef_t ef_mul_base(ef_t a, kb_t b) {
  ef_t r;
  for (int i = 0; i < 5; i++) r.c[i] = kb_mul(a.c[i], b);
  return r;
}
static inline ef_t ef_add_base(ef_t a, kb_t b) {
  a.c[0] = kb_add(a.c[0], b);
  return a;
}
/* product mod X^5 = 1 - X^2  (X^6 = X - X^3, X^7 = X^2 - X^4, X^8 = X^3 + X^2 - 1) */
static inline ef_t ef_mul(ef_t a, ef_t b) {
  kb_t d[9];
  for (int k = 0; k < 9; k++) d[k] = 0;
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) d[i + j] = kb_add(d[i + j], kb_mul(a.c[i], b.c[j]));
  ef_t r;
  r.c[0] = kb_sub(kb_add(d[0], d[5]), d[8]);
  r.c[1] = kb_add(d[1], d[6]);
  r.c[2] = kb_add(kb_add(kb_sub(kb_sub(d[2], d[5]), 0), d[7]), d[8]);
  r.c[3] = kb_add(kb_sub(d[3], d[6]), d[8]);
  r.c[4] = kb_sub(d[4], d[7]);
  return r;
}
static inline ef_t ef_sqr(ef_t a) { return ef_mul(a, a); }
static inline int ef_eq(ef_t a, ef_t b) {
  for (int i = 0; i < 5; i++)
    if (a.c[i] != b.c[i]) return 0;
  return 1;
}
static inline ef_t ef_pow(ef_t a, uint64_t e) {
  ef_t r = ef_one();
  while (e) {
    if (e & 1) r = ef_mul(r, a);
    a = ef_sqr(a);
    e >>= 1;
  }
  return r;
}
/* a^(p^5 - 2) by square-and-multiply over the 155-bit exponent (slow; host-side only). */
static inline ef_t ef_inv(ef_t a) {
  /* exponent p^5 - 2 as little-endian 32-bit limbs, computed at first use */
  static uint32_t limbs[5];
  static int init = 0;
  if (!init) {
    uint32_t acc[6] = {1, 0, 0, 0, 0, 0};
    for (int k = 0; k < 5; k++) {
      uint64_t carry = 0;
      for (int i = 0; i < 6; i++) {
        uint64_t v = (uint64_t)acc[i] * KB_P + carry;
        acc[i] = (uint32_t)v;
        carry = v >> 32;
      }
    }
    /* subtract 2 */
    uint64_t borrow = 2;
    for (int i = 0; i < 6 && borrow; i++) {
      uint64_t v = (uint64_t)acc[i];
      if (v >= borrow) {
        acc[i] = (uint32_t)(v - borrow);
        borrow = 0;
      } else {
        acc[i] = (uint32_t)(v + (1ull << 32) - borrow);
        borrow = 1;
      }
    }
    for (int i = 0; i < 5; i++) limbs[i] = acc[i];
    init = 1;
  }
  ef_t r = ef_one();
  for (int i = 4; i >= 0; i--)
    for (int b = 31; b >= 0; b--) {
      r = ef_sqr(r);
      if ((limbs[i] >> b) & 1) r = ef_mul(r, a);
    }
  return r;
}

#endif

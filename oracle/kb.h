/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's KoalaBear arithmetic.  Nothing under
 * oracle/ is part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * KoalaBear: p = 2^31 - 2^24 + 1, elements stored as u32 in Montgomery form
 * (x * 2^32 mod p), always canonical in [0, p).
 *   reference: crates/backend/koala-bear/src/koala_bear.rs:22-25 (p, mu)
 *              crates/backend/koala-bear/src/monty_31/utils.rs:65-127 (add/sub/reduce)
 *              crates/backend/koala-bear/src/monty_31/monty_31.rs:677-685 (mul)
 */
#ifndef LM_ORACLE_KB_H
#define LM_ORACLE_KB_H
#include <stdint.h>
#include <stddef.h>

#define KB_P 0x7f000001u
#define KB_MU 0x81000001u /* p^-1 mod 2^32 (non-negated convention) */

typedef uint32_t kb_t;

/* monty_31/utils.rs:65 */
static inline kb_t kb_add(kb_t a, kb_t b) {
  uint32_t s = a + b;
  return s >= KB_P ? s - KB_P : s;
}
/* monty_31/utils.rs:83 */
static inline kb_t kb_sub(kb_t a, kb_t b) {
  uint32_t d = a - b;
  return a < b ? d + KB_P : d;
}
static inline kb_t kb_neg(kb_t a) { return a ? KB_P - a : 0; }
/* monty_31/utils.rs:107: x in [0, 2^32 p) -> x * 2^-32 mod p in [0,p) */
static inline kb_t kb_monty_reduce(uint64_t x) {
  uint64_t t = (x * (uint64_t)KB_MU) & 0xffffffffull;
  uint64_t u = t * (uint64_t)KB_P;
  uint64_t d = x - u;
  uint32_t hi = (uint32_t)(d >> 32);
  return x < u ? hi + KB_P : hi;
}
static inline kb_t kb_mul(kb_t a, kb_t b) { return kb_monty_reduce((uint64_t)a * b); }
static inline kb_t kb_sqr(kb_t a) { return kb_mul(a, a); }
static inline kb_t kb_cube(kb_t a) { return kb_mul(kb_sqr(a), a); }
/* canonical integer -> Montgomery form (monty_31/utils.rs:9) */
static inline kb_t kb_from_u32(uint32_t x) { return (kb_t)((((uint64_t)x) << 32) % KB_P); }
static inline kb_t kb_from_i64(int64_t x) {
  int64_t r = x % (int64_t)KB_P;
  if (r < 0) r += KB_P;
  return kb_from_u32((uint32_t)r);
}
/* Montgomery form -> canonical integer (monty_31/utils.rs:50) */
static inline uint32_t kb_to_u32(kb_t a) { return kb_monty_reduce((uint64_t)a); }
#define KB_ONE 0x01fffffeu /* 2^32 mod p */
#define KB_ZERO 0u
static inline kb_t kb_double(kb_t a) { return kb_add(a, a); }
static inline kb_t kb_pow(kb_t a, uint64_t e) {
  kb_t r = KB_ONE;
  while (e) {
    if (e & 1) r = kb_mul(r, a);
    a = kb_sqr(a);
    e >>= 1;
  }
  return r;
}
static inline kb_t kb_inv(kb_t a) { return kb_pow(a, (uint64_t)KB_P - 2); }
/* halve: monty_31/utils.rs:94 */
static inline kb_t kb_halve(kb_t a) { return (a & 1) ? (a >> 1) + ((KB_P + 1) >> 1) : (a >> 1); }

/* 2^k-th roots of unity, canonical values; koala_bear.rs:46-54 (TWO_ADIC_GENERATORS). */
static const uint32_t KB_TWO_ADIC_GEN_CANON[25] = {
    0x1,        0x7f000000, 0x7e010002, 0x6832fe4a, 0x8dbd69c,  0xa28f031,  0x5c4a5b99, 0x29b75a80, 0x17668b8a,
    0x27ad539b, 0x334d48c7, 0x7744959c, 0x768fc6fa, 0x303964b2, 0x3e687d4d, 0x45a60e61, 0x6e2f4d7a, 0x163bd499,
    0x6c4a8a45, 0x143ef899, 0x514ddcad, 0x484ef19b, 0x205d63c3, 0x68e7dd49, 0x6ac49f88,
};
static inline kb_t kb_two_adic_generator(unsigned bits) { return kb_from_u32(KB_TWO_ADIC_GEN_CANON[bits]); }

#endif

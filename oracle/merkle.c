/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * Leaf sponge + Merkle tree over a row-major KoalaBear matrix.
 *   reference: crates/backend/symetric/src/sponge.rs:7-25   hash_slice (verifier form)
 *              .../sponge.rs:28-48  precompute_zero_suffix_state
 *              .../sponge.rs:52-108 hash_rtl_iter / absorb_rtl_chunks (prover form)
 *              crates/backend/symetric/src/compression.rs:5-15 compress (2-to-1)
 *              crates/backend/symetric/src/merkle.rs:21-47,50-90,92-125 tree / open / verify
 *              crates/whir/src/merkle.rs:59-88,205-288 build_merkle_tree_koalabear, open, first_digest_layer*
 * WIDTH = 16, RATE = OUT = DIGEST = 8.
 */
#include <stdlib.h>
#include <string.h>
#include "kb.h"
#include "oracle.h"

#define WIDTH 16
#define RATE 8

/* sponge.rs:7-25: data length multiple of 8, at least 16. */
void lm_or_hash_slice(const uint32_t *data, uint64_t len, uint32_t out[8]) {
  uint32_t st[WIDTH];
  uint64_t n_chunks = len / RATE;
  memcpy(st, data + len - WIDTH, sizeof(st));
  lm_or_poseidon1_compress(st);
  for (int64_t c = (int64_t)n_chunks - 3; c >= 0; c--) {
    memcpy(st + WIDTH - RATE, data + c * RATE, RATE * sizeof(uint32_t));
    lm_or_poseidon1_compress(st);
  }
  memcpy(out, st, 8 * sizeof(uint32_t));
}

/* sponge.rs:28-48 */
void lm_or_zero_suffix_state(uint32_t n_zero_chunks, uint32_t st[16]) {
  memset(st, 0, WIDTH * sizeof(uint32_t));
  lm_or_poseidon1_compress(st);
  for (uint32_t k = 0; k + 2 < n_zero_chunks; k++) {
    memset(st + WIDTH - RATE, 0, RATE * sizeof(uint32_t));
    lm_or_poseidon1_compress(st);
  }
}

/* Prover-side leaf digest following first_digest_layer / _with_initial_state
 * (whir/src/merkle.rs:215-288) with the right-to-left element iterator of
 * matrix.rs:96-110: the row is `stored_width` wide in memory, hashed as if
 * zero-extended to `full_width`; only the first `effective_width` entries are
 * assumed non-zero when >= 2 trailing rate chunks are zero. */
static void leaf_digest(const uint32_t *row, uint32_t stored_width, uint32_t full_width, uint32_t effective_width,
                        const uint32_t *zero_state, uint32_t out[8]) {
  uint32_t st[WIDTH];
  uint32_t n_zero_chunks = (full_width - effective_width) / RATE;
  /* right-to-left stream of elements: position `pos` counts down over the virtual row */
  int64_t pos;
  if (n_zero_chunks >= 2) {
    memcpy(st, zero_state, sizeof(st));
    uint32_t n_pad = (RATE - effective_width % RATE) % RATE;
    pos = (int64_t)effective_width + n_pad - 1; /* virtual width: effective rounded up to the rate */
  } else {
    pos = (int64_t)full_width - 1;
    for (int k = WIDTH - 1; k >= 0; k--, pos--) st[k] = ((uint64_t)pos < stored_width) ? row[pos] : 0;
    lm_or_poseidon1_compress(st);
  }
  while (pos >= 0) {
    for (int k = WIDTH - 1; k >= WIDTH - RATE; k--, pos--) {
      uint32_t lim = (n_zero_chunks >= 2) ? effective_width : stored_width;
      st[k] = ((uint64_t)pos < lim) ? row[pos] : 0;
    }
    lm_or_poseidon1_compress(st);
  }
  memcpy(out, st, 8 * sizeof(uint32_t));
}

/* First digest layer for an h x stored_width matrix; digests is h x 8. */
void lm_or_first_digest_layer(const uint32_t *mat, uint64_t h, uint32_t stored_width, uint32_t full_width,
                              uint32_t effective_width, uint32_t *digests) {
  uint32_t zero_state[WIDTH];
  uint32_t n_zero_chunks = (full_width - effective_width) / RATE;
  lm_or_poseidon1_init();
  if (n_zero_chunks >= 2) lm_or_zero_suffix_state(n_zero_chunks, zero_state);
  /* 16 rows per step on AVX-512 hosts (the reference's packing width, whir/src/merkle.rs:215-288), scalar otherwise and
   * for the ragged cases; both paths are compared in tests/test_oracle_poseidon.py */
  uint64_t h16 = 0;
  if (lm_or_have_avx512() && stored_width % 8 == 0 && effective_width % 8 == 0 && full_width >= 24 &&
      (n_zero_chunks >= 2 ? effective_width <= stored_width && effective_width >= 8 : 1) &&
      (uint64_t)stored_width * 16 < 0x7fffffffull) {
    h16 = h / 16 * 16;
#pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < h16; r += 16)
      lm_or_leaf_digest_x16(mat + r * stored_width, stored_width, full_width, effective_width,
                            n_zero_chunks >= 2 ? zero_state : NULL, digests + 8 * r);
  }
#pragma omp parallel for schedule(static)
  for (uint64_t r = h16; r < h; r++)
    leaf_digest(mat + r * stored_width, stored_width, full_width, effective_width, zero_state, digests + 8 * r);
}

/* compression.rs:5-15 */
void lm_or_compress_pair(const uint32_t left[8], const uint32_t right[8], uint32_t out[8]) {
  uint32_t st[WIDTH];
  memcpy(st, left, 8 * sizeof(uint32_t));
  memcpy(st + 8, right, 8 * sizeof(uint32_t));
  lm_or_poseidon1_compress(st);
  memcpy(out, st, 8 * sizeof(uint32_t));
}

/* symetric/src/merkle.rs:50-90 for power-of-two layers: next[i] = C(prev[2i] || prev[2i+1]). */
void lm_or_compress_layer(const uint32_t *prev, uint64_t n_prev, uint32_t *next) {
  lm_or_poseidon1_init();
  const uint64_t n = n_prev / 2;
  uint64_t n16 = 0;
  if (lm_or_have_avx512()) {
    n16 = n / 16 * 16;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n16; i += 16) lm_or_compress_pairs_x16(prev + 16 * i, next + 8 * i);
  }
#pragma omp parallel for schedule(static)
  for (uint64_t i = n16; i < n; i++) lm_or_compress_pair(prev + 16 * i, prev + 16 * i + 8, next + 8 * i);
}

/* Whole tree: layers stored back to back, layer 0 = h digests, then h/2, ... 1.
 * `layers` must hold (2h - 1) * 8 u32.  Returns nothing; root = last 8 words. */
void lm_or_merkle_tree(const uint32_t *mat, uint64_t h, uint32_t stored_width, uint32_t full_width,
                       uint32_t effective_width, uint32_t *layers) {
  lm_or_first_digest_layer(mat, h, stored_width, full_width, effective_width, layers);
  uint32_t *prev = layers;
  for (uint64_t n = h; n > 1; n >>= 1) {
    uint32_t *next = prev + 8 * n;
    lm_or_compress_layer(prev, n, next);
    prev = next;
  }
}

/* whir/src/merkle.rs:205-211 + symetric/src/merkle.rs:43-47:
 * row zero-extended to full_width, siblings layer[l][(index >> l) ^ 1] for l < log_h. */
void lm_or_merkle_open(const uint32_t *mat, uint64_t h, uint32_t stored_width, uint32_t full_width,
                       const uint32_t *layers, uint64_t index, uint32_t *out_row, uint32_t *out_path) {
  memset(out_row, 0, full_width * sizeof(uint32_t));
  memcpy(out_row, mat + index * stored_width, stored_width * sizeof(uint32_t));
  const uint32_t *layer = layers;
  uint32_t l = 0;
  for (uint64_t n = h; n > 1; n >>= 1, l++) {
    memcpy(out_path + 8 * l, layer + 8 * ((index >> l) ^ 1), 8 * sizeof(uint32_t));
    layer += 8 * n;
  }
}

/* symetric/src/merkle.rs:92-125 */
int lm_or_merkle_verify(const uint32_t root[8], uint32_t log_h, uint64_t index, const uint32_t *row,
                        uint32_t full_width, const uint32_t *path) {
  uint32_t cur[8], nxt[8];
  lm_or_hash_slice(row, full_width, cur);
  for (uint32_t l = 0; l < log_h; l++) {
    if ((index & 1) == 0)
      lm_or_compress_pair(cur, path + 8 * l, nxt);
    else
      lm_or_compress_pair(path + 8 * l, cur, nxt);
    memcpy(cur, nxt, sizeof(cur));
    index >>= 1;
  }
  return memcmp(cur, root, sizeof(cur)) == 0;
}

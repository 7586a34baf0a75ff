/* ORACLE — TEST INFRASTRUCTURE ONLY (see kb.h).
 *
 * Reed-Solomon encode of the stacked multilinear witness: block gather followed
 * by the radix-2 "DFT on evaluations".
 *   reference: crates/whir/src/utils.rs:69-150  reorder_and_dft, prepare_evals_for_fft_unpacked
 *              crates/whir/src/dft.rs:52-62     roots_of_unity_table
 *              crates/whir/src/dft.rs:79-144    dft_batch_by_evals (layer order: half-block 1,2,4,...,h/2)
 *              crates/whir/src/dft.rs:546-568   butterflies (a,b) -> (a + t(b-a), a - t(b-a))
 *              crates/whir/src/dft.rs:147-155   EF matrix = 5x wider base matrix
 * The reference groups layers (L1-sized initial chunk, then 3 per pass) purely for
 * cache reasons; the arithmetic is the plain layer-by-layer network below.
 */
#include <stdlib.h>
#include <string.h>
#include "kb.h"
#include "kb_avx512.h"
#include "oracle.h"

/* utils.rs:128-150.  evals: 2^n_vars elements of `dim` u32 each (dim = 1 base, 5 extension).
 * out: (2^(n_vars + log_inv_rate - folding)) rows x dft_n_cols elements. */
void lm_or_prepare_evals(const uint32_t *evals, uint32_t n_vars, uint32_t dim, uint32_t folding_factor,
                         uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t *out) {
  uint64_t full_len = ((uint64_t)1 << n_vars) << log_inv_rate;
  uint64_t block_size = full_len >> folding_factor;
  uint32_t log_block = n_vars + log_inv_rate - folding_factor;
  uint64_t out_len = block_size * dft_n_cols;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < out_len; i++) {
    uint64_t block_index = i % dft_n_cols;
    uint64_t offset_in_block = i / dft_n_cols;
    uint64_t src = ((block_index << log_block) + offset_in_block) >> log_inv_rate;
    for (uint32_t d = 0; d < dim; d++) out[i * dim + d] = evals[src * dim + d];
  }
}

/* One butterfly (dft.rs:546-568) between two rows of w elements, 16 columns per vector (w % 16 == 0). */
TGT void bfly_rows16(uint32_t *lo, uint32_t *hi, uint64_t w, kb_t t) {
  const __m512i tv = _mm512_set1_epi32((int)t);
  for (uint64_t c = 0; c < w; c += 16) {
    __m512i a = _mm512_loadu_si512(lo + c), b = _mm512_loadu_si512(hi + c);
    __m512i x = mul16(sub16(b, a), tv);
    _mm512_storeu_si512(lo + c, add16(a, x));
    _mm512_storeu_si512(hi + c, sub16(a, x));
  }
}

/* The same layer network as the scalar loop below, scheduled the way the reference schedules it for the cache
 * (dft.rs:79-144: a first run of layers on chunks that fit the cache, then three layers per pass over the matrix) and
 * with the butterflies of one row pair vectorised over the columns. */
TGT_FN static void dft_batch_avx512(uint32_t *mat, uint64_t h, uint64_t w, unsigned log_h, const kb_t *roots) {
  const unsigned LB = log_h < 10 ? log_h : 10;
#pragma omp parallel for schedule(static)
  for (uint64_t blk = 0; blk < (h >> LB); blk++) {
    uint32_t *base = mat + (blk << LB) * w;
    for (unsigned l = 0; l < LB; l++) {
      const uint64_t m = (uint64_t)1 << l, stride = h >> (l + 1);
      for (uint64_t pair = 0; pair < ((uint64_t)1 << LB) / 2; pair++) {
        const uint64_t b2 = pair >> l, i = pair & (m - 1);
        uint32_t *lo = base + (b2 * 2 * m + i) * w;
        bfly_rows16(lo, lo + m * w, w, roots[i * stride]);
      }
    }
  }
  for (unsigned l0 = LB; l0 < log_h;) {
    const unsigned g = log_h - l0 < 3 ? log_h - l0 : 3;
#pragma omp parallel for schedule(static)
    for (uint64_t u = 0; u < (h >> g); u++) {
      const uint64_t lo_part = u & (((uint64_t)1 << l0) - 1), hi_part = u >> l0;
      const uint64_t base_row = (hi_part << (l0 + g)) | lo_part;
      for (unsigned s = 0; s < g; s++) {
        const unsigned l = l0 + s;
        const uint64_t m = (uint64_t)1 << l;
        for (unsigned q = 0; q < (1u << g); q++) {
          if (q & (1u << s)) continue;
          const uint64_t ra = base_row + ((uint64_t)q << l0);
          bfly_rows16(mat + ra * w, mat + (ra + m) * w, w, roots[(ra & (m - 1)) * (h >> (l + 1))]);
        }
      }
    }
    l0 += g;
  }
}

/* dft.rs:79-144 on an h x w base-field matrix, in place. */
void lm_or_dft_batch_by_evals(uint32_t *mat, uint64_t h, uint64_t w) {
  if (h < 2) return;
  unsigned log_h = 0;
  while (((uint64_t)1 << log_h) < h) log_h++;
  /* nth_roots = 1, g, g^2, ..., g^(h/2-1);  layer with half-block m uses stride h/(2m) */
  kb_t g = kb_two_adic_generator(log_h);
  kb_t *roots = (kb_t *)malloc((h / 2) * sizeof(kb_t));
  roots[0] = KB_ONE;
  for (uint64_t i = 1; i < h / 2; i++) roots[i] = kb_mul(roots[i - 1], g);
  if (lm_or_have_avx512() && w % 16 == 0 && !getenv("LM_ORACLE_SCALAR_DFT")) {
    dft_batch_avx512(mat, h, w, log_h, roots);
    free(roots);
    return;
  }
  for (uint64_t m = 1; m < h; m <<= 1) {
    uint64_t stride = h / (2 * m);
#pragma omp parallel for schedule(static)
    for (uint64_t pair = 0; pair < h / 2; pair++) {
      uint64_t blk = pair / m, i = pair % m;
      uint32_t *lo = mat + (blk * 2 * m + i) * w;
      uint32_t *hi = lo + m * w;
      kb_t t = roots[i * stride];
      for (uint64_t c = 0; c < w; c++) {
        kb_t a = lo[c], b = hi[c];
        kb_t x = kb_mul(kb_sub(b, a), t);
        lo[c] = kb_add(a, x);
        hi[c] = kb_sub(a, x);
      }
    }
  }
  free(roots);
}

/* utils.rs:69-95: gather + DFT.  Output matrix has 2^(n_vars+log_inv_rate-folding) rows
 * and dft_n_cols * dim base columns. */
void lm_or_reorder_and_dft(const uint32_t *evals, uint32_t n_vars, uint32_t dim, uint32_t folding_factor,
                           uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t *out) {
  lm_or_prepare_evals(evals, n_vars, dim, folding_factor, log_inv_rate, dft_n_cols, out);
  uint64_t h = (uint64_t)1 << (n_vars + log_inv_rate - folding_factor);
  lm_or_dft_batch_by_evals(out, h, (uint64_t)dft_n_cols * dim);
}

/* Layers [l_first, log_h) of the evals-DFT network applied to a SUBSET of rows held locally, for the row-sharded
 * multi-GPU commit: local row (m, j') with m < n_blocks, j' < run  <->  global row
 * m * block + offset + j'.  Only layers whose partner row is also local may be requested, i.e.
 * 2^l_first >= block (partner = other m, same j').  Same butterfly and twiddles as dft.rs:79-144,546-568. */
void lm_or_dft_layers_mapped(uint32_t *mat, uint64_t w, uint32_t log_h, uint32_t l_first, uint64_t n_blocks,
                             uint64_t run, uint64_t block, uint64_t offset) {
  uint64_t h = (uint64_t)1 << log_h;
  kb_t g = kb_two_adic_generator(log_h);
  for (uint32_t l = l_first; l < log_h; l++) {
    uint64_t m_stride = ((uint64_t)1 << l) / block; /* partner block distance */
#pragma omp parallel for schedule(static)
    for (uint64_t idx = 0; idx < n_blocks / 2 * run; idx++) {
      uint64_t pair = idx / run, jp = idx % run;
      uint64_t m_lo = (pair / m_stride) * 2 * m_stride + pair % m_stride;
      uint64_t m_hi = m_lo + m_stride;
      uint64_t grow = m_lo * block + offset + jp; /* global row of the low element */
      uint64_t e = (grow & (((uint64_t)1 << l) - 1)) << (log_h - l - 1);
      kb_t t = kb_pow(g, e);
      uint32_t *lo = mat + (m_lo * run + jp) * w, *hi = mat + (m_hi * run + jp) * w;
      for (uint64_t c = 0; c < w; c++) {
        kb_t a = lo[c], b = hi[c];
        kb_t x = kb_mul(kb_sub(b, a), t);
        lo[c] = kb_add(a, x);
        hi[c] = kb_sub(a, x);
      }
    }
    (void)h;
  }
}

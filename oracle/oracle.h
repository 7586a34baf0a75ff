/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * C restatement of the reference's (leanEthereum/leanMultisig) algorithms for the
 * proving hot path, used exclusively as the checker in tests/, in
 * __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference arm.
 * The product (leanmultisig_b200/) never links, imports or calls anything here.
 *
 * Parity pin status: the Poseidon1 permutation is pinned by the reference's
 * known-answer test (poseidon1_koalabear_16.rs:1066-1093).  The reference holds no
 * stored vectors for Merkle roots, codewords or sumcheck polynomials, and it cannot
 * be compiled here (Rust, no toolchain) — those are "parity unpinned" by stored
 * vectors and are pinned structurally instead: prover-form vs verifier-form of the
 * leaf sponge, DFT vs direct MLE evaluation (the reference's own property test,
 * whir/src/dft.rs:582-604), and prover -> verifier acceptance.
 */
#ifndef LM_ORACLE_H
#define LM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* poseidon1.c */
void lm_or_poseidon1_init(void);
void lm_or_poseidon1_permute_dense(uint32_t s[16]);
void lm_or_poseidon1_permute(uint32_t s[16]);
void lm_or_poseidon1_compress(uint32_t s[16]);
void lm_or_poseidon1_permute_batch(uint32_t *states, uint64_t n, int dense);
void lm_or_poseidon1_compress_batch(uint32_t *states, uint64_t n);
/* poseidon1_avx512.c: 16 states per call, one per 32-bit lane; only when lm_or_have_avx512() (LM_ORACLE_NO_AVX512=1 forces
 * the scalar path everywhere) */
int lm_or_have_avx512(void);
void lm_or_poseidon1_compress_x16(uint32_t *states);
void lm_or_leaf_digest_x16(const uint32_t *rows, uint32_t stored_width, uint32_t full_width, uint32_t effective_width,
                           const uint32_t *zero_state, uint32_t *digests);
void lm_or_compress_pairs_x16(const uint32_t *prev, uint32_t *next);

/* merkle.c */
void lm_or_hash_slice(const uint32_t *data, uint64_t len, uint32_t out[8]);
void lm_or_zero_suffix_state(uint32_t n_zero_chunks, uint32_t st[16]);
void lm_or_first_digest_layer(const uint32_t *mat, uint64_t h, uint32_t stored_width, uint32_t full_width,
                              uint32_t effective_width, uint32_t *digests);
void lm_or_compress_pair(const uint32_t left[8], const uint32_t right[8], uint32_t out[8]);
void lm_or_compress_layer(const uint32_t *prev, uint64_t n_prev, uint32_t *next);
void lm_or_merkle_tree(const uint32_t *mat, uint64_t h, uint32_t stored_width, uint32_t full_width,
                       uint32_t effective_width, uint32_t *layers);
void lm_or_merkle_open(const uint32_t *mat, uint64_t h, uint32_t stored_width, uint32_t full_width,
                       const uint32_t *layers, uint64_t index, uint32_t *out_row, uint32_t *out_path);
int lm_or_merkle_verify(const uint32_t root[8], uint32_t log_h, uint64_t index, const uint32_t *row,
                        uint32_t full_width, const uint32_t *path);

/* dft.c */
void lm_or_prepare_evals(const uint32_t *evals, uint32_t n_vars, uint32_t dim, uint32_t folding_factor,
                         uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t *out);
void lm_or_dft_batch_by_evals(uint32_t *mat, uint64_t h, uint64_t w);
void lm_or_reorder_and_dft(const uint32_t *evals, uint32_t n_vars, uint32_t dim, uint32_t folding_factor,
                           uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t *out);

void lm_or_dft_layers_mapped(uint32_t *mat, uint64_t w, uint32_t log_h, uint32_t l_first, uint64_t n_blocks,
                             uint64_t run, uint64_t block, uint64_t offset);

/* poly.c */
void lm_or_eq_table(const uint32_t *point, uint32_t k, const uint32_t scalar[5], uint32_t *out);
void lm_or_expand_from_univariate(const uint32_t y[5], uint32_t n, uint32_t *out);
void lm_or_mle_eval(const uint32_t *evals, uint32_t n, uint32_t dim, const uint32_t *point, uint32_t out[5]);
void lm_or_fold_msb(const uint32_t *in, uint64_t n_in, uint32_t dim, const uint32_t r[5], uint32_t *out);
void lm_or_ef_mul(const uint32_t a[5], const uint32_t b[5], uint32_t out[5]);
void lm_or_ef_inv(const uint32_t a[5], uint32_t out[5]);
uint32_t lm_or_kb_mul(uint32_t a, uint32_t b);
uint32_t lm_or_kb_from_u32(uint32_t a);
uint32_t lm_or_kb_to_u32(uint32_t a);
uint32_t lm_or_kb_inv(uint32_t a);
uint32_t lm_or_kb_two_adic_generator(uint32_t bits);
void lm_or_set_num_threads(int n);
int lm_or_max_threads(void);

/* sumcheck.c */
void lm_or_weights_add_eq(uint32_t *weights, uint64_t selector, const uint32_t *point, uint32_t m,
                          const uint32_t scalar[5]);
void lm_or_next_mle_folded(const uint32_t *oc, uint32_t n, uint32_t *res);
void lm_or_weights_add_next(uint32_t *weights, uint64_t selector, const uint32_t *point, uint32_t m,
                            const uint32_t scalar[5]);
void lm_or_weights_add_base_eq(uint32_t *weights, uint32_t m, const uint32_t *points, uint32_t n_q,
                               const uint32_t *scalars);
void lm_or_prod_round(const uint32_t *p, uint32_t dim, const uint32_t *w, uint64_t n, uint32_t c0[5], uint32_t c2[5]);
void lm_or_evals_to_coeffs(uint32_t *data, uint64_t n);

/* air.c */
int lm_or_air_shape(uint32_t table, uint32_t out[3]); /* n_cols, n_shift, degree */
void lm_or_air_eval(uint32_t table, const uint32_t *point, const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la,
                    const uint32_t beta[5], uint32_t out[5]);
void lm_or_air_round(uint32_t table, const uint32_t *cols, uint64_t n, uint32_t dim, const uint32_t *eq_point,
                     const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la, const uint32_t beta[5], uint32_t *out);
void lm_or_poseidon16_fill_trace(uint32_t *cols, uint64_t n);
void lm_or_air_exec_eval(const uint32_t *point, const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la,
                         const uint32_t beta[5], uint32_t out[5]);
void lm_or_shift_column(const uint32_t *col, uint64_t n, uint32_t *out);
void lm_or_air_exec_round(const uint32_t *cols, uint64_t n, uint32_t dim, const uint32_t *eq_point,
                          const uint32_t *alpha_powers, const uint32_t *la, uint32_t n_la, const uint32_t beta[5],
                          uint32_t *out);
void lm_or_fold_lsb(const uint32_t *in, uint64_t n_in, uint32_t dim, const uint32_t r[5], uint32_t *out);

/* gkr.c */
void lm_or_gkr_layer_up(const uint32_t *nums, uint32_t num_dim, const uint32_t *dens, uint64_t n, uint32_t *out_nums,
                        uint32_t *out_dens);
void lm_or_gkr_round(const uint32_t *nl, const uint32_t *nr, const uint32_t *dl, const uint32_t *dr, uint64_t n,
                     const uint32_t *eq_point, const uint32_t alpha[5], uint32_t c0[5], uint32_t c2[5]);
void lm_or_finger_print(const uint32_t *data, uint64_t n_rows, uint32_t n_data, const uint32_t *alphas,
                        const uint32_t c[5], uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif

/* leanmultisig_b200 — C ABI of the B200-native proving hot path for leanEthereum/leanMultisig.
 *
 * Every entry point replaces one call site of the reference's Rust prover (the reference has no FFI of its own;
 * INTEGRATION.md shows the Rust `extern "C"` block and the three-line patch per call site).  Conventions:
 *   - field element  F  = u32 in Montgomery form, canonical in [0, p), p = 2^31 - 2^24 + 1
 *                        (crates/backend/koala-bear/src/monty_31/monty_31.rs:32-42)
 *   - extension      EF = 5 consecutive F, coefficient of X^0 first, F[X]/(X^5 + X^2 - 1)
 *                        (crates/backend/koala-bear/src/quintic_extension/extension.rs:26-36)
 *   - digest            = 8 F (crates/backend/symetric/src/merkle.rs:11)
 *   - multilinear index = variable x_0 is the most-significant bit (crates/backend/poly/src/evals.rs:219)
 *   - every function returns LM_OK (0) or a negative status and never unwinds; lm_last_error() gives the text.
 *     The reference panics on violated preconditions; the Rust shim turns a non-zero status into a panic.
 *   - host pointers are borrowed for the duration of the call only; results either go to caller-allocated
 *     buffers or stay on the device behind an opaque handle with an explicit *_free.
 *   - calls on one lm_ctx are serialised by the caller (the prover drives them from the thread that owns the
 *     Fiat-Shamir transcript, crates/backend/fiat-shamir/src/traits.rs:15-44); each call is synchronous with
 *     respect to its host-visible outputs.  lm_dev_* calls only enqueue on the context's stream.
 */
#ifndef LEANMULTISIG_B200_H
#define LEANMULTISIG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LM_OK 0
#define LM_ERR_INVALID (-1)   /* violated precondition (the reference would panic) */
#define LM_ERR_CUDA (-2)      /* CUDA runtime failure, see lm_last_error() */
#define LM_ERR_NO_DEVICE (-3) /* no CUDA device: there is no CPU fallback */
#define LM_ERR_OOM (-4)

#define LM_DIGEST_ELEMS 8

typedef struct lm_ctx lm_ctx;   /* one per GPU: stream, twiddle table, scratch arena */
typedef struct lm_tree lm_tree; /* committed matrix + all Merkle layers (+ the polynomial), device resident */

const char* lm_last_error(void);
int lm_device_count(void);
/* number of CUDA kernels this library has launched so far in this process */
uint64_t lm_kernel_launches(void);

/* Replaces setup_prover / precompute_dft_twiddles (src/lib.rs:13-16, crates/whir/src/utils.rs:200-202) and the
 * OnceLock Poseidon constants (poseidon1_koalabear_16.rs:575): selects `device`, creates a stream and uploads
 * the twiddle table for transforms of up to 2^max_log_domain rows (<= 24). */
int lm_init(int device, uint32_t max_log_domain, lm_ctx** out_ctx);
int lm_destroy(lm_ctx* ctx);
/* Run subsequent work on an existing CUDA stream (cudaStream_t passed as void*); NULL restores the own stream. */
int lm_set_stream(lm_ctx* ctx, void* cuda_stream);
int lm_sync(lm_ctx* ctx);
/* Pin / unpin a caller-owned host buffer so that commits read it at full PCIe rate (cudaHostRegister). */
int lm_host_register(void* ptr, size_t bytes);
int lm_host_unregister(void* ptr);

/* ---- WHIR commit: crates/whir/src/commit.rs:64-85 (reorder_and_dft + MerkleData::build) -------------------
 * evals: 2^n_vars elements of elem_dim words (1 = base field, 5 = extension), entries >= actual_len are zero
 * (only the first actual_len are read).  Builds the Reed-Solomon codeword matrix of
 * 2^(n_vars + log_inv_rate - folding_factor) rows x 2^folding_factor columns (columns that are entirely zero
 * are not stored, commit.rs:70-74), hashes every row with the Poseidon1 sponge and builds the Merkle tree.
 * out_root receives the root; the handle owns codeword, digest layers and a device copy of the live evals. */
int lm_commit(lm_ctx* ctx, const uint32_t* evals, uint32_t n_vars, uint32_t elem_dim, uint64_t actual_len,
              uint32_t folding_factor, uint32_t log_inv_rate, lm_tree** out_tree, uint32_t out_root[8]);
/* Same, evals already resident on this context's device (e.g. the output of an on-device fold). */
int lm_commit_dev(lm_ctx* ctx, const uint32_t* d_evals, uint32_t n_vars, uint32_t elem_dim, uint64_t actual_len,
                  uint32_t folding_factor, uint32_t log_inv_rate, int retain_evals, lm_tree** out_tree,
                  uint32_t out_root[8]);
/* MerkleData::open (crates/whir/src/commit.rs:34-45 -> crates/whir/src/merkle.rs:205-211): for each index the
 * full row zero-extended to full width (n x full_width words) and the sibling path, leaf level first
 * (n x log2(height) x 8 words). */
/* Witness-side steps in front of the first commitment.
 * lm_commit_stacked: stack_polynomials_and_commit (crates/sub_protocols/src/stacked_pcs.rs:100-157) without the host-side
 * global_polynomial: every segment (memory, memory_acc, bytecode_acc, the table columns) is copied straight to its offset
 * of the device-resident polynomial (zero elsewhere, actual_len = end of the last segment) and committed like lm_commit.
 * lm_access_counts: memory_acc / bytecode_acc (crates/lean_prover/src/prove_execution.rs:91-110): for every index column k
 * (Montgomery-form addresses) acc[addr + j] += 1, j < n_values[k]; out_acc: table_len field elements (Montgomery). */
typedef struct {
  const uint32_t* data; /* host, base field */
  uint64_t len, offset; /* in elements */
} lm_segment;
int lm_commit_stacked(lm_ctx* ctx, const lm_segment* segments, uint32_t n_segments, uint32_t n_vars,
                      uint32_t folding_factor, uint32_t log_inv_rate, lm_tree** out_tree, uint32_t out_root[8]);
int lm_access_counts(lm_ctx* ctx, const uint32_t* const* index_cols, const uint64_t* n_rows, const uint32_t* n_values,
                     uint32_t n_cols, uint64_t table_len, uint32_t* out_acc);
int lm_open(lm_tree* tree, const uint64_t* indices, uint32_t n, uint32_t* out_rows, uint32_t* out_paths);
/* lm_open plus the STIR answer of every opened leaf (crates/whir/src/open.rs:161-190): out_evals[q] = the leaf of index q,
 * read as a multilinear polynomial in fold_vars = log2(elements per leaf) variables, evaluated at fold_point (fold_vars x 5,
 * the folding randomness of the round); the fold runs on the device next to the gather. */
int lm_open_fold(lm_tree* tree, const uint64_t* indices, uint32_t n, const uint32_t* fold_point, uint32_t fold_vars,
                 uint32_t* out_rows, uint32_t* out_paths, uint32_t* out_evals);
/* Verifier side of the same openings: replaces the per-query loops of
 * crates/whir/src/verify.rs:229-345 (verify_stir_challenges: merkle_verify of every opening of a round against the
 * round's root, crates/whir/src/merkle.rs:115-150 -> crates/backend/symetric/src/merkle.rs:92-122 with hash_slice,
 * sponge.rs:7-25) and, when fold_point is given, the fold of every opened leaf at the round's folding randomness that
 * follows it (verify.rs: answers -> evaluate(folding_randomness)).  All pointers are HOST memory.
 *   rows      n x width words (leaf data as the proof carries it: width = elements per leaf x elem_dim, a multiple of 8, >= 16)
 *   paths     n x log_height x 8 words, leaf level first (restored paths: fiat-shamir/src/merkle_pruning.rs)
 *   out_ok    n bytes: 1 = the opening hashes to root.  An index >= 2^log_height or a word that is not a canonical
 *             Montgomery residue fails its opening (out_ok = 0); it is not an error of the call.
 *   out_evals n x 5 words (only with fold_point / fold_vars; 2^fold_vars x elem_dim == width), may be NULL otherwise */
int lm_verify_openings(lm_ctx* ctx, const uint32_t root[8], uint32_t log_height, const uint64_t* indices, uint32_t n,
                       const uint32_t* rows, uint32_t width, uint32_t elem_dim, const uint32_t* paths,
                       const uint32_t* fold_point, uint32_t fold_vars, uint8_t* out_ok, uint32_t* out_evals);
/* height (rows), full row width in words, stored row width in words, elem_dim */
int lm_tree_shape(const lm_tree* tree, uint64_t* height, uint32_t* full_width, uint32_t* stored_width,
                  uint32_t* elem_dim);
/* Out-of-domain sample: polynomial.evaluate(point) on the committed polynomial (commit.rs:89-92,
 * crates/backend/poly/src/evals.rs:142).  point: n_vars x 5 words; out: 5 words. */
int lm_tree_eval(lm_tree* tree, const uint32_t* point, uint32_t out[5]);
/* Copy device-resident pieces back (parity tests, debugging): codeword matrix height x stored_width and the
 * digest layers (2*height - 1) x 8, leaf layer first. */
int lm_tree_read_codeword(lm_tree* tree, uint32_t* out);
int lm_tree_read_layers(lm_tree* tree, uint32_t* out);
int lm_tree_free(lm_tree* tree);

/* MleRef::evaluate on a host polynomial (crates/backend/poly/src/evals.rs:142-347): evals has 2^n_vars
 * elements of elem_dim words of which the first live_len may be non-zero. */
int lm_mle_eval(lm_ctx* ctx, const uint32_t* evals, uint32_t n_vars, uint32_t elem_dim, uint64_t live_len,
                const uint32_t* point, uint32_t out[5]);

/* ---- WHIR open: product sumcheck session -----------------------------------------------------------------
 * Replaces SumcheckSingle (crates/whir/src/open.rs:323-446) and the bodies it calls; the Fiat-Shamir transcript
 * stays with the caller, so the reference's run_product_sumcheck / sumcheck_prove_many_rounds loops
 * (crates/backend/sumcheck/src/product_computation.rs:37-125, prove.rs:86-151) become, per round:
 *   lm_sc_round / lm_sc_fold_round -> (c0, c2);  c1 = sum - 2 c0 - c2;  transcript absorbs (c0, c1, c2), grinds,
 *   samples r;  after the last round lm_sc_fold(r).
 * The session holds the polynomial table p (base field until first folded, then EF) and the weight table w (EF),
 * both of 2^n_vars entries, MSB-first folding (crates/backend/poly/src/utils.rs:161-186). */
typedef struct lm_sumcheck lm_sumcheck;
/* p = the polynomial committed in `tree` (borrowed from the tree, which must outlive the session); w = 0 */
int lm_sc_new_from_tree(lm_tree* tree, lm_sumcheck** out);
/* p = host polynomial (2^n_vars entries of elem_dim words, first live_len possibly non-zero); w = 0 */
int lm_sc_new(lm_ctx* ctx, const uint32_t* evals, uint32_t n_vars, uint32_t elem_dim, uint64_t live_len,
              lm_sumcheck** out);
/* combine_statement terms (open.rs:518-584): w[(selector << m) + x] += scalar * eq(point, x)   (point: m x 5)
 * resp. scalar * next_mle(point, x) (crates/backend/poly/src/next_mle.rs:35).  Also add_new_equality
 * (open.rs:337-358) with selector 0 and m = n_vars. */
int lm_sc_add_eq(lm_sumcheck* sc, uint64_t selector, const uint32_t* point, uint32_t m, const uint32_t scalar[5]);
/* n_statements eq statements that share selector and point length, added in ONE pass over the weight table (combine_statement
 * adds them one after the other; the sums are the same field elements): points n_statements x m x 5, scalars n_statements x 5 */
int lm_sc_add_eq_batch(lm_sumcheck* sc, uint64_t selector, const uint32_t* points, uint32_t m, const uint32_t* scalars,
                       uint32_t n_statements);
int lm_sc_add_next(lm_sumcheck* sc, uint64_t selector, const uint32_t* point, uint32_t m, const uint32_t scalar[5]);
/* weights[base + (b << shift) + offset] += scalar * eq(point, b) for b < 2^pre (point: pre x 5 words, first coordinate = most
 * significant bit of b).  The building block of the next-row statement (crates/backend/poly/src/next_mle.rs:35-58: term k has
 * shift = k + 1, offset = 2^k) in the form a row-range SHARD of the weight table needs, where the index bits that select the
 * shard are fixed and drop out of the local index (leanmultisig_b200/sharded.py). */
int lm_sc_add_strided_eq(lm_sumcheck* sc, uint64_t base, uint32_t shift, uint64_t offset, const uint32_t* point, uint32_t pre,
                         const uint32_t scalar[5]);
/* add_new_base_equality (open.rs:360-382): w[x] += sum_q scalars[q] * eq(points[q], x); points: n_q x n_vars
 * base-field words, scalars: n_q x 5 */
int lm_sc_add_base_eq(lm_sumcheck* sc, const uint32_t* points, uint32_t n_q, const uint32_t* scalars);
/* round polynomial h(X) = c0 + c1 X + c2 X^2 of the current tables (product_computation.rs:127-170) */
int lm_sc_round(lm_sumcheck* sc, uint32_t c0[5], uint32_t c2[5]);
/* fold both tables with challenge r (n_vars decreases by one) */
int lm_sc_fold(lm_sumcheck* sc, const uint32_t r[5]);
/* fold with r and compute the next round polynomial in the same pass (product_computation.rs:242-304) */
int lm_sc_fold_round(lm_sumcheck* sc, const uint32_t r[5], uint32_t c0[5], uint32_t c2[5]);
int lm_sc_num_vars(const lm_sumcheck* sc, uint32_t* n_vars, uint32_t* poly_dim);
/* copy the current tables to the host: poly 2^n_vars x poly_dim words, weights 2^n_vars x 5 (NULL to skip) */
int lm_sc_read(lm_sumcheck* sc, uint32_t* out_poly, uint32_t* out_weights);
/* OOD sample on the current (folded) polynomial (open.rs:96-99); point: n_vars x 5 */
int lm_sc_eval_poly(lm_sumcheck* sc, const uint32_t* point, uint32_t out[5]);
/* commit the current (folded) polynomial: reorder_and_dft + MerkleData::build of a WHIR round (open.rs:81-91) */
int lm_sc_commit_poly(lm_sumcheck* sc, uint32_t folding_factor, uint32_t log_inv_rate, lm_tree** out_tree,
                      uint32_t out_root[8]);
/* Row-sharded sumcheck (leanmultisig_b200/sharded.py): after the folding rounds that are local to a rank the
 * folded tables are exchanged between devices.  lm_sc_export_dev copies the current EF tables (2^n_vars x 5 words each)
 * into caller-owned DEVICE buffers, lm_sc_new_from_dev starts a session from gathered DEVICE tables. */
int lm_sc_export_dev(lm_sumcheck* sc, uint32_t* d_poly_out, uint32_t* d_weights_out);
int lm_sc_new_from_dev(lm_ctx* ctx, const uint32_t* d_poly, const uint32_t* d_weights, uint32_t n_vars, lm_sumcheck** out);
int lm_sc_free(lm_sumcheck* sc);

/* ---- AIR ("SuperSpartan") sumcheck session --------------------------------------------------------------
 * Replaces AirSumcheckSession (crates/sub_protocols/src/air_sumcheck.rs:45-292), the implementor of
 * trait OuterSumcheckSession (air_sumcheck.rs:34-42) that prove_batched_air_sumcheck (air_sumcheck.rs:636-681)
 * drives.  A Rust struct implementing the trait forwards compute_bare_round_poly -> lm_air_round (+ the p(1) /
 * Lagrange step it already does on a handful of values), process_challenge -> lm_air_fold,
 * final_column_evals -> lm_air_final.  table_id: 0 = execution table (crates/lean_vm/src/tables/execution/air.rs:42-130,
 * 20 columns), 1 = extension_op (tables/extension_op/air.rs:44-163, 29 columns), 2 = poseidon16
 * (tables/poseidon_16/mod.rs:294-548, 109 columns); table_id | LM_AIR_NO_BUS selects the BUS = false instantiation of
 * tables 1 and 2 (no bus constraint: logup_alphas_eq may be NULL with n_la = 0, bus_beta is ignored).
 * cols: n_cols host pointers to base-field columns of 2^log_rows entries in natural row order (the shifted
 * columns are derived on the device).  eq_factor: log_rows x 5; the LAST entry belongs to the variable bound
 * first.  alpha_powers: n_alpha x 5 (ExtraDataForBuses::alpha_powers), logup_alphas_eq: n_la x 5, bus_beta: 5. */
typedef struct lm_air lm_air;
#define LM_AIR_EXECUTION 0u
#define LM_AIR_EXTENSION_OP 1u
#define LM_AIR_POSEIDON16 2u
#define LM_AIR_NO_BUS 0x100u
int lm_air_new(lm_ctx* ctx, uint32_t table_id, const uint32_t* const* cols, uint32_t n_cols, uint32_t log_rows,
               const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha,
               const uint32_t* logup_alphas_eq, uint32_t n_la, const uint32_t bus_beta[5], lm_air** out);
/* One row-range shard of a table split over several GPUs (leanmultisig_b200/sharded.py): cols hold the
 * 2^log_rows rows of this shard, eq_factor the log_rows entries of the variables inside the shard, eq_scale the eq value
 * of the shard's row-index prefix (multiplied into every weight, so the per-rank round sums only need adding up),
 * halo_next_row[k] = first row of column k (k < n_shift) in the NEXT shard, the one-row halo of
 * compute_shifted_columns (air_sumcheck.rs:683-694); NULL for the last shard (the last row repeats). */
int lm_air_new_shard(lm_ctx* ctx, uint32_t table_id, const uint32_t* const* cols, uint32_t n_cols, uint32_t log_rows,
                     const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha,
                     const uint32_t* logup_alphas_eq, uint32_t n_la, const uint32_t bus_beta[5],
                     const uint32_t* halo_next_row, const uint32_t eq_scale[5], lm_air** out);
/* Session over already folded columns: cols_ef = (n_cols + n_shift) columns of 2^log_rows EF entries (5 words each,
 * column after column).  Used for the last log2(G) rounds of a sharded sumcheck, after the all-gather of the per-shard
 * column values. */
int lm_air_new_folded(lm_ctx* ctx, uint32_t table_id, const uint32_t* cols_ef, uint32_t n_cols_total, uint32_t log_rows,
                      const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha,
                      const uint32_t* logup_alphas_eq, uint32_t n_la, const uint32_t bus_beta[5], lm_air** out);
/* Session over columns that ALREADY live on this context's device, e.g. inside the committed stacked polynomial (the
 * columns of one table are consecutive 2^log_rows-element segments there, stacked_pcs.rs:118-135): d_cols = n_cols
 * columns of 2^log_rows base-field elements, column after column.  The execution table BORROWS them (no copy; they must
 * stay valid and unchanged until lm_air_free), the wide tables copy them next to their shifted columns on the device. */
int lm_air_new_dev(lm_ctx* ctx, uint32_t table_id, const uint32_t* d_cols, uint32_t n_cols, uint32_t log_rows,
                   const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha, const uint32_t* logup_alphas_eq,
                   uint32_t n_la, const uint32_t bus_beta[5], lm_air** out);
int lm_air_info(const lm_air* air, uint32_t* n_vars, uint32_t* degree, uint32_t* n_cols_total);
/* out_evals: degree x 5 words = sum_j eq(j) C(row pair j at z) for z = 0, 2, 3, .., degree over the WHOLE
 * hypercube (no separate padding term), before the missing_mul_factor scaling (air_sumcheck.rs:242-249) */
int lm_air_round(lm_air* air, uint32_t* out_evals);
/* bind the least-significant remaining variable to r (fold_multilinear_at_bit, crates/backend/poly/src/utils.rs:117) */
int lm_air_fold(lm_air* air, const uint32_t r[5]);
/* after the last fold: the (n_cols + n_shift) column evaluations, 5 words each (air_sumcheck.rs:289-291) */
int lm_air_final(lm_air* air, uint32_t* out);
int lm_air_free(lm_air* air);
/* fill_trace_poseidon_16 (crates/lean_vm/src/tables/poseidon_16/trace_gen.rs:10-165): cols = the 109 host columns of
 * the poseidon16 table, n_rows entries each; reads flag_permute (column 8) and the 16 inputs (columns 9..24), writes
 * columns 25..108 (two post-full-round states, 20 partial-round S-box outputs, one more state, the 16 outputs). */
int lm_poseidon16_fill_trace(lm_ctx* ctx, uint32_t* const* cols, uint64_t n_rows);

/* ---- Logup: fingerprints + quotient GKR -----------------------------------------------------------------
 * lm_finger_print replaces finger_print(_packed) (crates/utils/src/multilinear.rs:76-98): out[r] = c - sum_i
 * alphas[i] * data[r][i] for n_rows rows of n_data base-field words (row-major); alphas: n_data x 5; out: n_rows x 5.
 *
 * lm_gkr_* replaces prove_gkr_quotient (crates/sub_protocols/src/quotient_gkr/mod.rs:31-141) with the transcript
 * left to the caller.  nums (base field) / dens (EF) are the ACTIVE prefix in NATURAL order (the reference's
 * chunk-bit-reversed packing of logup.rs:88-199 is a CPU SIMD layout and is not used); the rest of the next
 * power of two is (0, 1).  lm_gkr_new runs the whole up pass (sum_quotients_2_by_2, layers.rs:124-189) and keeps
 * every layer.  Per layer, top to bottom (mod.rs:73-141, sumcheck_utils.rs:282-359):
 *   lm_gkr_layer_begin(K, claim_point[K x 5], alpha);  K times { lm_gkr_round -> (c0_raw, c2_raw);  host:
 *   build_bare_from_coeffs (sumcheck_utils.rs:491-503), transcript, r;  lm_gkr_fold(r) };  lm_gkr_layer_end ->
 *   inner evals [nl, nr, dl, dr].  Rounds bind the least-significant variable first; eq_alpha of a round is the
 *   last remaining coordinate of the claim point.  Sums run over the whole hypercube (no separate padding term). */
typedef struct lm_gkr lm_gkr;
int lm_finger_print(lm_ctx* ctx, const uint32_t* data, uint64_t n_rows, uint32_t n_data, const uint32_t* alphas,
                    const uint32_t c[5], uint32_t* out);
int lm_gkr_new(lm_ctx* ctx, const uint32_t* nums, const uint32_t* dens, uint64_t active_len, lm_gkr** out);
/* the same from device-resident arrays (d_nums: active_len F, d_dens: active_len x 5 words); both are copied */
int lm_gkr_new_dev(lm_ctx* ctx, const uint32_t* d_nums, const uint32_t* d_dens, uint64_t active_len, lm_gkr** out);
/* One row-range shard of the fraction table split over several GPUs (leanmultisig_b200/sharded.py): the shard
 * has 2^n_vars rows of which the first active_len (possibly 0) are given, the rest are (0, 1); the up pass stops at
 * 2^top_vars fractions per shard (5 - log2(G), so that the gathered tops are the 2^5 values the prover sends). */
int lm_gkr_new_shard(lm_ctx* ctx, const uint32_t* nums, const uint32_t* dens, uint64_t active_len, uint32_t n_vars,
                     uint32_t top_vars, lm_gkr** out);
/* lm_gkr_layer_begin on a shard: point holds the claim_vars coordinates inside the shard, eq_scale the eq value of the
 * shard's row prefix (multiplied into every weight, so that the per-rank (c0, c2) only need adding up). */
int lm_gkr_layer_begin_shard(lm_gkr* gkr, uint32_t claim_vars, const uint32_t* claim_point, const uint32_t alpha[5],
                             const uint32_t eq_scale[5]);
int lm_gkr_num_vars(const lm_gkr* gkr, uint32_t* n_vars);
/* log2 of the number of fractions the up pass stops at: 5 (N_VARS_TO_SEND_GKR_COEFFS), 5 - log2(G) for a shard session */
int lm_gkr_top_vars(const lm_gkr* gkr, uint32_t* top_vars);
/* the 2^top_vars numerators and denominators of the top layer, 5 words each: 32 x 5 for a whole table (mod.rs:64-66) */
int lm_gkr_top(lm_gkr* gkr, uint32_t* top_nums, uint32_t* top_dens);
int lm_gkr_layer_begin(lm_gkr* gkr, uint32_t claim_vars, const uint32_t* claim_point, const uint32_t alpha[5]);
int lm_gkr_round(lm_gkr* gkr, uint32_t c0[5], uint32_t c2[5]);
int lm_gkr_fold(lm_gkr* gkr, const uint32_t r[5]);
int lm_gkr_layer_end(lm_gkr* gkr, uint32_t* inner_evals /* 4 x 5 */);
int lm_gkr_free(lm_gkr* gkr);

/* ---- Logup table assembly ----------------------------------------------------------------------------------
 * Replaces the table build of prove_generic_logup (crates/sub_protocols/src/logup.rs:52-211): numerators (base field)
 * and denominators (EF) of every section, written back to back in NATURAL row order into device buffers that
 * lm_logup_finish hands to the quotient-GKR session without a copy.  The caller (which knows Table::bus() /
 * Table::lookups(), crates/lean_vm/src/tables/table_trait.rs:19-56) lists the sections in the reference's order:
 *   memory, bytecode (+ padding up to the tallest table), then per table (tallest first): [execution only: bytecode
 *   lookup], bus, one section per looked-up value column.
 * Section rows:  numerator = 1 | +num_col[r] | -num_col[r] | 0;
 *                denominator = c + den_sign * (alphas.last * domainsep + sum_i alphas[i] * data_i[r])   (den_sign = +-1)
 *                              or 1 when den_sign = 0 (padding)                 (finger_print_packed, multilinear.rs:87-98)
 * Host columns are uploaded once per (pointer, len) and stay cached for lm_logup_col_eval, which serves the column
 * evaluations that follow the GKR (logup.rs:224-305). */
typedef struct lm_logup lm_logup;
#define LM_LOGUP_NUM_ONE 0u
#define LM_LOGUP_NUM_COL 1u
#define LM_LOGUP_NUM_NEG_COL 2u
#define LM_LOGUP_NUM_ZERO 3u
#define LM_LOGUP_DATA_COL 0u   /* col[offset + r * stride] + value */
#define LM_LOGUP_DATA_ROW 1u   /* the row index r */
#define LM_LOGUP_DATA_CONST 2u /* value */
typedef struct {
  const uint32_t* col; /* host array (Montgomery words) of `len` entries, or NULL */
  uint64_t len, offset, stride;
  uint32_t kind;
  uint32_t value; /* canonical integer constant */
} lm_logup_data;
int lm_logup_new(lm_ctx* ctx, uint64_t total_active_len, const uint32_t c[5], const uint32_t* alphas_eq_poly,
                 uint32_t n_alphas, lm_logup** out);
int lm_logup_section(lm_logup* logup, uint64_t n_rows, uint32_t num_mode, const uint32_t* num_col, int32_t den_sign,
                     uint32_t domainsep, const lm_logup_data* data, uint32_t n_data);
/* MleRef::evaluate of a (cached or new) host column of `len` <= 2^n_vars entries; point: n_vars x 5 */
int lm_logup_col_eval(lm_logup* logup, const uint32_t* col, uint64_t len, uint32_t n_vars, const uint32_t* point,
                      uint32_t out[5]);
/* copy what has been filled so far back to the host (parity tests); either pointer may be NULL */
/* the same for n_cols (<= 256) columns at ONE point: eq tables built once, one launch pair per column, one read-back
 * (logup.rs:224-305 evaluates every looked-up column of a table at the same inner point); out: n_cols x 5 words */
int lm_logup_col_eval_batch(lm_logup* logup, const uint32_t* const* cols, const uint64_t* lens, uint32_t n_cols,
                            uint32_t n_vars, const uint32_t* point, uint32_t* out);
int lm_logup_read(lm_logup* logup, uint32_t* out_nums, uint32_t* out_dens);
/* all total_active_len rows filled: pad to the next power of two with (0, 1) and run the GKR up pass */
int lm_logup_finish(lm_logup* logup, lm_gkr** out);
int lm_logup_free(lm_logup* logup);

/* ---- device-pointer layer (inputs already in HBM; used by the kernel-only benchmark and by lm_* above) -----
 * All pointers are device pointers on ctx's device; work is enqueued on ctx's stream, no synchronisation. */
int lm_dev_alloc(lm_ctx* ctx, size_t bytes, void** out);
int lm_dev_free(lm_ctx* ctx, void* ptr);
int lm_dev_upload(lm_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);   /* synchronous */
int lm_dev_download(lm_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* synchronous */
/* Poseidon1KoalaBear16::permute / compress_in_place on n 16-word states (poseidon1_koalabear_16.rs:873,1020) */
int lm_dev_poseidon1(lm_ctx* ctx, uint32_t* d_states, uint64_t n, int compress);
/* reorder_and_dft (crates/whir/src/utils.rs:69): d_out = 2^(n_vars+log_inv_rate-folding) x (dft_n_cols*dim) */
int lm_dev_reorder_and_dft(lm_ctx* ctx, const uint32_t* d_evals, uint32_t n_vars, uint32_t elem_dim,
                           uint32_t folding_factor, uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_out);
/* EvalsDft::dft_batch_by_evals (crates/whir/src/dft.rs:79), in place on a height x width matrix */
int lm_dev_dft(lm_ctx* ctx, uint32_t* d_mat, uint64_t height, uint64_t width);
/* Local transform of ONE rank of the row-sharded commit with the exchange fused into its last pass: row i of
 * the rank's 2^(n_vars + log_inv_rate - folding) x dft_n_cols result is stored straight into peer_mats[i >> log_run] at
 * local row (rank << log_run) | (i mod 2^log_run), log_run = log2(rows / world) — NVLink stores through peer pointers
 * (CUDA IPC mappings of the other ranks' matrices; peer_mats[rank] = the caller's own), so no separate all-to-all runs.
 * d_work: rows x dft_n_cols words of scratch for the passes before the last one.  The caller orders the ranks (a
 * stream-ordered barrier) before reading its matrix.  Base field only, dft_n_cols % 4 == 0. */
int lm_dev_reorder_and_dft_scatter(lm_ctx* ctx, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding_factor,
                                   uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_work, const uint64_t* peer_mats,
                                   uint32_t world, uint32_t rank);
/* Column-range forms for the pipelined host-input path of the sharded commit (copy of the next column group overlaps the
 * transform / exchange / hash of the current one): the transform + exchange of columns [col_begin, col_begin + col_count)
 * (multiples of 8), the last layers on a column range of the received matrix (multiples of 4), and `count` sponge steps
 * absorbing rate chunks chunk_hi, chunk_hi - 1, .. of every row into the 8-word states kept in d_digests (the call that
 * takes chunk eff_w / 8 - 1 seeds them from the zero-suffix state; needs >= 2 all-zero trailing chunks). */
int lm_dev_reorder_and_dft_scatter_cols(lm_ctx* ctx, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding_factor,
                                        uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_work, const uint64_t* peer_mats,
                                        uint32_t world, uint32_t rank, uint32_t col_begin, uint32_t col_count);
int lm_dev_dft_layers_mapped_cols(lm_ctx* ctx, uint32_t* d_mat, uint64_t width, uint32_t log_h, uint32_t l_first,
                                  uint64_t n_blocks, uint64_t run, uint64_t block, uint64_t offset, uint64_t col_begin,
                                  uint64_t col_count);
int lm_dev_merkle_absorb(lm_ctx* ctx, const uint32_t* d_mat, uint64_t height, uint32_t stored_width, uint32_t full_width,
                         uint32_t effective_width, uint32_t chunk_hi, uint32_t count, uint32_t* d_digests);
/* lm_dev_dft_layers_mapped(_cols) with the result written to a second matrix of the same shape (col_count = 0: all columns).
 * The row-sharded commit reads the matrix its peers stored into and writes the witness's private codeword in the same pass, so
 * the exchange buffer is free for the next commit without a device-to-device copy. */
int lm_dev_dft_layers_mapped_out(lm_ctx* ctx, const uint32_t* d_mat, uint32_t* d_out, uint64_t width, uint32_t log_h,
                                 uint32_t l_first, uint64_t n_blocks, uint64_t run, uint64_t block, uint64_t offset,
                                 uint64_t col_begin, uint64_t col_count);
/* CUDA IPC plumbing for the peer matrices (one process per GPU): the owner exports a buffer it got from lm_dev_alloc, the
 * other ranks open the 64-byte handle from THEIR device (peer access over NVLink is enabled by the open) and close it at
 * the end.  The handle bytes travel through the caller's own channel (torch.distributed all_gather_object here). */
int lm_dev_ipc_export(lm_ctx* ctx, const void* d_ptr, uint8_t handle[64]);
int lm_dev_ipc_open(lm_ctx* ctx, const uint8_t handle[64], void** d_peer);
int lm_dev_ipc_close(lm_ctx* ctx, void* d_peer);
/* the last layers [l_first, log_h) of dft_batch_by_evals on the rows one rank holds after the all-to-all of a
 * row-sharded commit: local row (m, j'), m < n_blocks, j' < run, is global row m * block + offset + j' */
int lm_dev_dft_layers_mapped(lm_ctx* ctx, uint32_t* d_mat, uint64_t width, uint32_t log_h, uint32_t l_first,
                             uint64_t n_blocks, uint64_t run, uint64_t block, uint64_t offset);
/* build_merkle_tree_koalabear (crates/whir/src/merkle.rs:59-88): d_layers = (2*height - 1) x 8 words */
int lm_dev_merkle_tree(lm_ctx* ctx, const uint32_t* d_mat, uint64_t height, uint32_t stored_width,
                       uint32_t full_width, uint32_t effective_width, uint32_t* d_layers);
/* the two halves of lm_dev_merkle_tree, separately launchable (per-kernel timing): first_digest_layer
 * (crates/whir/src/merkle.rs:215-288) into d_layers[0 .. height), then MerkleTree::from_first_layer
 * (crates/backend/symetric/src/merkle.rs:21-35) over the already filled leaf layer */
int lm_dev_merkle_leaves(lm_ctx* ctx, const uint32_t* d_mat, uint64_t height, uint32_t stored_width,
                         uint32_t full_width, uint32_t effective_width, uint32_t* d_layers);
int lm_dev_merkle_levels(lm_ctx* ctx, uint32_t* d_layers, uint64_t height);
/* eval_multilinear (evals.rs:142) with device-resident evals and point; d_out: 5 words */
int lm_dev_mle_eval(lm_ctx* ctx, const uint32_t* d_evals, uint32_t n_vars, uint32_t elem_dim, uint64_t live_len,
                    const uint32_t* d_point, uint32_t* d_out);
/* fold_multilinear (crates/backend/poly/src/utils.rs:161-186): d_out = n_in/2 EF */
int lm_dev_fold_msb(lm_ctx* ctx, const uint32_t* d_in, uint64_t n_in, uint32_t elem_dim, const uint32_t r[5],
                    uint32_t* d_out);
/* fill_trace_poseidon_16 on a device-resident column-major table (109 x n_rows words) */
int lm_dev_poseidon16_fill_trace(lm_ctx* ctx, uint32_t* d_cols, uint64_t n_rows);
/* eval_eq_scaled (crates/backend/poly/src/eq_mle.rs:20-26): d_out = 2^k EF; point is a HOST pointer (k x 5) */
int lm_dev_eq_table(lm_ctx* ctx, const uint32_t* point, uint32_t k, const uint32_t scalar[5], uint32_t* d_out);

/* ---- Fiat-Shamir support ---------------------------------------------------------------------------------
 * The transcript (Challenger / ProverState, crates/backend/fiat-shamir/src/{challenger,prover}.rs) stays on the
 * reference side of the boundary; these two calls are the parts of it that are worth offloading or sharing. */
/* ProverState::pow_grinding's search (fiat-shamir/src/prover.rs:135-167): `state` = the challenger's 16-word state
 * (host), result = the SMALLEST canonical witness w such that lane 8 of permute(state[0..8] | w | 0^7), as a
 * canonical integer, has `bits` low zero bits.  The caller observes the witness exactly as the reference does. */
int lm_pow_grind(lm_ctx* ctx, const uint32_t state[16], uint32_t bits, uint64_t* witness);
/* Poseidon1KoalaBear16::permute_mut on ONE state on the host (poseidon1_koalabear_16.rs:873): the duplex sponge
 * of challenger.rs:32-36 is sequential, one permutation per observation, so it is not a device job. */
int lm_host_poseidon1_permute(uint32_t state[16]);
/* The same permutation through a CPU model of the tensor-core formulation the Merkle kernels run (csrc/poseidon1_umma.cuh:
 * identical B-matrix image, row layout, recombination and block structure, every tcgen05.mma replaced by the integer dot
 * products it stands for).  A test hook: lets the CPU tier pin the formulation against poseidon1_koalabear_16.rs:873 without a
 * GPU.  Not a product path. */
int lm_host_poseidon1_umma_model(uint32_t state[16]);
/* The B-matrix image those kernels stage in shared memory (test hook: the CPU tier checks on the actual constants that every s32
 * accumulator column and every carry-free recombination stays inside its bound).  Returns the image size in bytes; copies it to
 * `out` when `capacity` is large enough (out may be NULL to query the size). */
uint64_t lm_host_poseidon1_umma_image(uint8_t* out, uint64_t capacity);
/* CPU model of the tensor-core statement-weights kernel behind lm_sc_add_eq_batch (csrc/sumcheck.cu weights_gemm_kernel: same
 * image builder, row layout and recombination, the MMAs as integer dot products) - test hook of the CPU tier:
 * w[(x_hi << lo_vars) + x_lo] += sum_k hi_k[x_hi] * lo_k[x_lo], all extension elements as 5 words; hi: n_statements tables of
 * 2^hi_vars entries one after the other, lo likewise. */
int lm_host_eq_gemm_model(uint32_t* w, const uint32_t* hi, const uint32_t* lo, uint32_t n_statements, uint32_t hi_vars,
                          uint32_t lo_vars);

/* ---- Host spine in C++ (csrc/spine.cu), built above the entry points of this header ---------------------------
 * lm_fs mirrors ProverState + Challenger (crates/backend/fiat-shamir/src/prover.rs:28-178, challenger.rs:8-76): the
 * Poseidon1 duplex runs on the host, pow_grinding searches on the device (lm_pow_grind).  A Rust integration keeps its
 * own ProverState and does not need these; they exist so that callers without the Rust spine (the Python mirror, the
 * benchmarks) do not pay an interpreter round trip per sumcheck round.  ctx may be NULL when pow_grinding is not used. */
typedef struct lm_fs lm_fs;
int lm_fs_new(lm_ctx* ctx, lm_fs** out);
int lm_fs_free(lm_fs* fs);
int lm_fs_add_scalars(lm_fs* fs, const uint32_t* words, uint64_t n);   /* add_base_scalars / add_extension_scalars */
int lm_fs_observe(lm_fs* fs, const uint32_t* words, uint64_t n);       /* absorbed, not sent */
int lm_fs_duplex(lm_fs* fs);
/* prover.rs:105-128: coeffs n_coeffs x 5; eq_alpha != NULL absorbs expand_bare_to_full(coeffs, eq_alpha) */
int lm_fs_add_sumcheck_polynomial(lm_fs* fs, const uint32_t* coeffs, uint32_t n_coeffs, const uint32_t* eq_alpha);
int lm_fs_sample(lm_fs* fs, uint32_t n, uint32_t* out /* n x 5 */);     /* sample_vec */
int lm_fs_sample_in_range(lm_fs* fs, uint32_t bits, uint32_t n, uint64_t* out);
int lm_fs_pow_grinding(lm_fs* fs, uint32_t bits);
int lm_fs_transcript_len(const lm_fs* fs, uint64_t* n_words);
int lm_fs_transcript(const lm_fs* fs, uint32_t* out);
int lm_fs_state(const lm_fs* fs, uint32_t state[16], int* rate_fresh);
/* The weight update that closes a WHIR round (crates/whir/src/open.rs:192-231): with comb the combination randomness, the
 * OOD constraints eq(expand(y_k)) get weight comb^k and the STIR constraints eq(expand(gen^idx_q)) weight comb^(n_ood + q);
 * total_io += sum_k comb^k ood_answers[k] + sum_q comb^(n_ood + q) stir_evals[q].  ood_ys: n_ood x 5 (the sampled
 * univariate points), gen: Montgomery form of the folded domain's generator.  Host arithmetic in C++, weights through
 * lm_sc_add_eq / lm_sc_add_base_eq. */
int lm_whir_stir_update(lm_sumcheck* sc, const uint64_t* idx, uint32_t n_q, uint32_t gen, uint32_t num_variables,
                        const uint32_t comb[5], const uint32_t* ood_ys, const uint32_t* ood_answers, uint32_t n_ood,
                        const uint32_t* stir_evals, uint32_t total_io[5]);
/* The sumcheck phase of a WHIR round in one call (sumcheck_prove_many_rounds, crates/backend/sumcheck/src/prove.rs:86-151, with
 * the product computation of product_computation.rs:37-125): n_rounds times { device round (the fold by the previous challenge
 * fused in), c1 from the running sum, absorb (c0, c1, c2) and send (c1, c2), PoW of pow_bits, sample the challenge }, then the
 * fold by the last challenge.  total_io: the claimed sum, updated; out_challenges: n_rounds x 5.  Same transcript as the
 * per-round entry points (lm_sc_round / lm_sc_fold_round + lm_fs_*), without the caller's round trips. */
int lm_whir_sumcheck_rounds(lm_sumcheck* sc, lm_fs* fs, uint32_t n_rounds, uint32_t pow_bits, uint32_t total_io[5],
                            uint32_t* out_challenges);
/* Hand-over of the sponge to a caller-owned transcript and back (state 16 words, rate_fresh as in challenger.rs): a Rust
 * ProverState that wants the device-resident challenger of lm_gkr_prove / lm_air_prove_batched exports its Challenger into
 * an lm_fs, runs the phase, and re-imports state + the transcript words the phase appended. */
int lm_fs_set_state(lm_fs* fs, const uint32_t state[16], int rate_fresh);
/* prove_gkr_quotient (crates/sub_protocols/src/quotient_gkr/mod.rs:31-141) end to end: top values, per-layer sumchecks
 * (compute_round / build_bare_from_coeffs / fold), inner evaluations.  out_point: n_vars x 5.  The challenger of the layer
 * sumchecks runs ON THE DEVICE (csrc/devfs.cuh: Poseidon1 duplex of challenger.rs:8-76, add_sumcheck_polynomial and
 * sample of prover.rs:74-128 inside the round kernels): the launches of all layers are enqueued back to back and the sponge
 * state + appended transcript words come back once, at the end. */
int lm_gkr_prove(lm_gkr* gkr, lm_fs* fs, uint32_t out_quotient[5], uint32_t* out_point, uint32_t out_claim_num[5],
                 uint32_t out_claim_den[5]);
/* the same with the round loop and the sponge on the host (one synchronisation per round); cross-check of lm_gkr_prove */
int lm_gkr_prove_hostloop(lm_gkr* gkr, lm_fs* fs, uint32_t out_quotient[5], uint32_t* out_point, uint32_t out_claim_num[5],
                          uint32_t out_claim_den[5]);
/* prove_batched_air_sumcheck (crates/sub_protocols/src/air_sumcheck.rs:636-681) over n_sessions sessions: eq_factors =
 * the sessions' eq_factor arrays one after the other (n_vars_i x 5 each), sums = n_sessions x 5 initial sums.
 * out_challenges: max_i n_vars_i x 5; afterwards lm_air_final gives each session's column evaluations. */
int lm_air_prove_batched(lm_air* const* airs, uint32_t n_sessions, const uint32_t* eq_factors, const uint32_t* sums,
                         const uint32_t eta[5], lm_fs* fs, uint32_t* out_challenges, uint32_t* out_n_rounds);

/* ---- proof wire format (csrc/wire.cpp, leanmultisig_b200/wire.py): host code -----------------------------------------
 * The compression half of TypeOneMultiSignature::compress / decompress (crates/rec_aggregation/src/type_1_aggregation.rs:
 * 81-89): lz4_flex::compress_prepend_size / decompress_size_prepended = 4-byte little-endian uncompressed length + one LZ4
 * block.  lm_lz4_decompress_size_prepended with dst == NULL only reports the uncompressed length. */
uint64_t lm_lz4_compress_bound(uint64_t n);
int lm_lz4_compress_prepend_size(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* out_len);
int lm_lz4_decompress_size_prepended(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* out_len);

#ifdef __cplusplus
}
#endif
#endif

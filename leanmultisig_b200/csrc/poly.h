// Internal (C++) launch interface of poly.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace lm {
// d_out[b] = scalar * eq(point, b), b < 2^k big-endian; point is k x 5 words on the device
cudaError_t eq_table(cudaStream_t stream, const uint32_t* d_point, int k, const uint32_t scalar[5], uint32_t* d_out);
// scratch (in u32 words) mle_eval needs for an n_vars-variate polynomial
size_t mle_eval_scratch_words(uint32_t n_vars);
// d_out[0..5) = MLE(evals)(point); evals[live_len..) are zero and never read
cudaError_t mle_eval(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim, uint64_t live_len,
                     const uint32_t* d_point, uint32_t* d_scratch, uint32_t* d_out);
// d_out (n_in / 2 EF) = MSB-first fold of d_in (n_in elements of `dim` words) with challenge r
// (entries >= live are zero and not read; EF tables may be folded in place, d_out == d_in)
cudaError_t fold_msb(cudaStream_t stream, const uint32_t* d_in, uint64_t n_in, uint32_t dim, uint64_t live,
                     const uint32_t r[5], uint32_t* d_out);
// counts[addr + j] += 1 (integer counters) for the Montgomery-form addresses of one index column; then integer -> field
cudaError_t access_count(cudaStream_t stream, const uint32_t* d_idx_col, uint64_t n, uint32_t n_values, uint64_t table_len,
                         uint32_t* d_counts, uint32_t* d_bad);
cudaError_t counts_to_monty(cudaStream_t stream, uint32_t* d_counts, uint64_t n);
// n_cols base-field columns (host array of device pointers) at one point; d_out: n_cols x 5 words
cudaError_t mle_eval_batch(cudaStream_t stream, const uint32_t* const* d_cols, const uint64_t* live_lens, uint32_t n_cols,
                           uint32_t n_vars, const uint32_t* d_point, uint32_t* d_scratch, uint32_t* d_out);
// d_out[q] = MLE of row q (2^k elements of `dim` words, index MSB = first coordinate) at d_point (k x 5 words): the STIR answer of
// an opened leaf (open.rs:161-190)
cudaError_t rows_mle_eval(cudaStream_t stream, const uint32_t* d_rows, uint32_t n_rows, uint32_t dim, uint32_t k, const uint32_t* d_point,
                          uint32_t* d_out);
}  // namespace lm

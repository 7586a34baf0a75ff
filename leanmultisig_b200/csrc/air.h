// Internal (C++) launch interface of air.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace lm {
cudaError_t air_shift_column(cudaStream_t stream, const uint32_t* d_col, uint64_t n, uint32_t* d_out);
size_t air_round_scratch_words(uint32_t log_n);
// d_out[5 x 5] = evaluations at z = 0, 2, 3, 4, 5 of the execution-table round polynomial over 22 SoA columns of
// 2^log_n rows (dim words per entry); d_eq_point: log_n - 1 EF entries on the device; the rest are host arrays.
cudaError_t air_exec_round(cudaStream_t stream, const uint32_t* d_cols, uint32_t dim, uint32_t log_n, const uint32_t* d_eq_point,
                           const uint32_t* alpha_powers, const uint32_t* la, uint32_t n_la, const uint32_t beta[5],
                           uint32_t* d_scratch, uint32_t* d_out, const uint32_t* eq_scale = nullptr);
// fold the least-significant variable of n_cols SoA columns (n rows -> n/2 EF rows); d_out must not alias d_in
cudaError_t air_fold_lsb(cudaStream_t stream, const uint32_t* d_in, uint32_t dim, uint64_t n, uint32_t n_cols, const uint32_t r[5],
                         uint32_t* d_out);
// ---- air_generic.cu: extension_op (table 1) and poseidon16 (table 2); table | 0x100 = BUS = false instantiation ----
bool air_table_shape(uint32_t table, uint32_t* n_cols, uint32_t* n_shift, uint32_t* degree, uint32_t* max_constraints);
// d_out[degree x 5] = evaluations at z = 0, 2, .., degree over (n_cols + n_shift) SoA columns of 2^log_n rows
cudaError_t air_generic_round(cudaStream_t stream, uint32_t table, const uint32_t* d_cols, uint32_t dim, uint32_t log_n,
                              const uint32_t* d_eq_point, const uint32_t* alpha_powers, uint32_t n_alpha, const uint32_t* la,
                              uint32_t n_la, const uint32_t beta[5], uint32_t* d_scratch, uint32_t* d_out,
                              const uint32_t* eq_scale = nullptr);
// eq_scale (host, 5 words, nullptr = 1): constant EF factor multiplied into every eq weight — the eq value of the row-range
// prefix when the session covers one shard of a table split over several GPUs.
// columns 25..109 of the poseidon16 table from its flag_permute and input columns (column-major, n rows each)
cudaError_t poseidon16_fill_trace(cudaStream_t stream, uint32_t* d_cols, uint64_t n);
}  // namespace lm

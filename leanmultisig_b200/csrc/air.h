// Internal (C++) launch interface of air.cu / air_generic.cu; the public C ABI is include/leanmultisig_b200.h.
//
// Column storage of a session: base-field columns u32[c][n]; once folded, extension columns as COEFFICIENT PLANES
// u32[c][5][n] (coefficient k of row i of column c at ((c * 5 + k) * n + i)), so that a thread's 8 / 16-byte loads of
// adjacent rows are coalesced across the warp.  With n = 1 this is the [F; 5] layout of the final column evaluations.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "kb.cuh"

namespace lm {
struct AirDev {   // per-session device state
  Ef r[2];        // r[0]: the latest challenge, r[1]: the one before it (both pending while the table is still base field)
  Ef out[12];     // round sums of the last round (z = 0, 2, .., degree)
  uint32_t counter;
};
constexpr int AIR_MAX_BLOCKS = 148 * 4;
// eq tables of the session: eqtab.cuh's prefix tables over the session's eq_factor, built once
cudaError_t air_build_eq_tables(cudaStream_t stream, const uint32_t* d_eq_point, uint32_t n_vars, const uint32_t eq_scale[5],
                                uint32_t* d_tab);

// ---- execution table (air.cu): fused fold + round -----------------------------------------------------------------
// The table is read once per round.  While it is still base field the pending challenges are applied on the fly
// (rows 2j, 2j+1 of the current table = 2 / 4 / 8 base rows folded by 0 / 1 / 2 challenges); the round with two pending
// challenges writes the first extension table (a quarter of the rows), every later round folds the previous extension
// table with the latest challenge while reading it and writes the next one.  The two shifted columns are never
// materialised in the base field: they are rows i + 1 of columns 0 and 1 (`halo` = value after the last row).
enum AirExecMode { AIR_B0 = 0, AIR_B1 = 1, AIR_B2 = 2, AIR_E0 = 3, AIR_E1 = 4 };
struct AirExecArgs {
  const uint32_t* base;  // 20 base columns (modes B*)
  uint64_t n_base;       // rows per base column
  uint32_t halo[2];
  const uint32_t* src;   // extension planes [22][5][rows_src] (modes E*)
  uint32_t* dst;         // extension planes [22][5][2^(m+1)] written by modes B2 and E1
  uint32_t m;            // pairs j < 2^m
  uint32_t k;            // variables of the session (eq tables)
  const uint32_t* eq_tab;
  uint32_t* partial;     // AIR_MAX_BLOCKS x 25 words
  AirDev* d;
};
// alpha_powers: >= 13 x 5, la: logup alphas (n_la x 5, first 4 and last used), beta: bus challenge (host arrays)
// r0_host (5 words, may be nullptr): the first challenge, known to the host; with it round 1 (mode AIR_B1) runs in the base
// field as polynomials in r0 (air.cu, "round 1 in the base field")
cudaError_t air_exec_round(cudaStream_t stream, int mode, const AirExecArgs& a, const uint32_t* alpha_powers, const uint32_t* la,
                           uint32_t n_la, const uint32_t beta[5], const uint32_t* r0_host = nullptr);
// the pending folds applied to the last rows: d_out[22][5]
cudaError_t air_exec_final(cudaStream_t stream, int mode, const AirExecArgs& a, uint32_t* d_out);

// ---- generic pieces (extension_op, poseidon16, sessions started from folded columns) -------------------------------
cudaError_t air_shift_column(cudaStream_t stream, const uint32_t* d_col, uint64_t n, uint32_t* d_out);
// [n_cols][n][5] -> planes [n_cols][5][n]
cudaError_t air_aos_to_planes(cudaStream_t stream, const uint32_t* d_aos, uint32_t n_cols, uint64_t n, uint32_t* d_planes);
size_t air_round_scratch_words();
// fold the least-significant variable of n_cols columns (n rows -> n/2 rows of extension planes); d_out must not alias d_in
cudaError_t air_fold_lsb(cudaStream_t stream, const uint32_t* d_in, uint32_t dim, uint64_t n, uint32_t n_cols, const uint32_t r[5],
                         uint32_t* d_out);
// ---- air_generic.cu: extension_op (table 1) and poseidon16 (table 2); table | 0x100 = BUS = false instantiation ----
bool air_table_shape(uint32_t table, uint32_t* n_cols, uint32_t* n_shift, uint32_t* degree, uint32_t* max_constraints);
// d_out[degree x 5] = evaluations at z = 0, 2, .., degree over (n_cols + n_shift) columns of 2^log_n rows; k = variables of
// the session the eq tables were built for
cudaError_t air_generic_round(cudaStream_t stream, uint32_t table, const uint32_t* d_cols, uint32_t dim, uint32_t log_n,
                              const uint32_t* d_eq_tab, uint32_t k, const uint32_t* alpha_powers, uint32_t n_alpha,
                              const uint32_t* la, uint32_t n_la, const uint32_t beta[5], uint32_t* d_scratch, uint32_t* d_out);
// columns 25..109 of the poseidon16 table from its flag_permute and input columns (column-major, n rows each)
cudaError_t poseidon16_fill_trace(cudaStream_t stream, uint32_t* d_cols, uint64_t n);
}  // namespace lm

// Internal (C++) launch interface of gkr.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "kb.cuh"

namespace lm {
// one section of the Logup table (see logup_fill_kernel in gkr.cu); all pointers are device pointers
constexpr int LOGUP_MAX_DATA = 16;
enum { LOGUP_NUM_ONE = 0, LOGUP_NUM_COL = 1, LOGUP_NUM_NEG_COL = 2, LOGUP_NUM_ZERO = 3 };
enum { LOGUP_DATA_COL = 0, LOGUP_DATA_ROW = 1, LOGUP_DATA_CONST = 2 };
struct LogupData {
  const uint32_t* col;
  uint64_t offset, stride;
  uint32_t add;   // Montgomery-form constant added to the column value (or the constant itself)
  uint32_t kind;
};
struct LogupSection {
  uint64_t n_rows;
  int num_mode;
  const uint32_t* num_col;
  int den_sign;  // +1: c + fp, -1: c - fp, 0: denominator 1 (padding)
  int n_data;
  Ef c, contrib;
  Ef alphas[LOGUP_MAX_DATA];
  LogupData data[LOGUP_MAX_DATA];
};
cudaError_t logup_fill_section(cudaStream_t stream, const LogupSection& section, uint32_t* d_nums, uint32_t* d_dens);
cudaError_t finger_print(cudaStream_t stream, const uint32_t* d_data, uint64_t n_rows, uint32_t n_data,
                         const uint32_t* d_alphas, const uint32_t c[5], uint32_t* d_out);
cudaError_t gkr_pad(cudaStream_t stream, uint32_t* d_nums, uint32_t num_dim, uint32_t* d_dens, uint64_t active, uint64_t n);
cudaError_t gkr_layer_up(cudaStream_t stream, const uint32_t* d_nums, uint32_t num_dim, const uint32_t* d_dens, uint64_t n,
                         uint32_t* d_out_nums, uint32_t* d_out_dens);
size_t gkr_round_scratch_words(uint32_t n_vars);
cudaError_t gkr_round(cudaStream_t stream, int src, uint32_t num_dim, const uint32_t* a, const uint32_t* b, uint32_t n_vars,
                      const uint32_t* d_eq_point, const uint32_t alpha[5], uint32_t* d_scratch, uint32_t* d_out10,
                      const uint32_t* eq_scale = nullptr);  // host, 5 words: constant factor of every eq weight (nullptr = 1)
cudaError_t gkr_fold(cudaStream_t stream, int src, uint32_t num_dim, const uint32_t* a, const uint32_t* b, uint32_t n_vars,
                     const uint32_t r[5], uint32_t* d_out);
}  // namespace lm

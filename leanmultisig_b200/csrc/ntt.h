// Internal (C++) launch interface of ntt.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace lm {
uint32_t two_adic_generator_monty(unsigned bits);
// Every d_tw passed to this interface holds 2^(log_n - 1) + NTT_TW_SCRATCH_WORDS words: the table, then scratch for the
// compact per-pass twiddle tables the pass kernels read (written by the launcher, stream-ordered).
constexpr size_t NTT_TW_SCRATCH_WORDS = 4 * 4096;
// d_tw[e] = g^e, e < 2^(log_n - 1), g = primitive 2^log_n-th root of unity (Montgomery form)
cudaError_t ntt_fill_twiddles(cudaStream_t stream, uint32_t* d_tw, unsigned log_n);
// in-place evals-DFT of an h x w row-major matrix, skipping the first `skip_layers` layers
cudaError_t ntt_dft_batch_by_evals(cudaStream_t stream, uint32_t* d_mat, uint64_t h, uint64_t w, int skip_layers,
                                   const uint32_t* d_tw, unsigned tw_log_n);
// gather + DFT: d_out is (2^(n_vars + log_inv_rate - folding)) x (dft_n_cols * dim)
cudaError_t ntt_reorder_and_dft(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim,
                                uint32_t folding_factor, uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_out,
                                const uint32_t* d_tw, unsigned tw_log_n);
cudaError_t ntt_reorder_and_dft_cols(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding_factor,
                                     uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t col_begin, uint32_t col_count,
                                     uint32_t* d_out, const uint32_t* d_tw, unsigned tw_log_n);
// local transform of one rank of the row-sharded commit with the all-to-all fused into the stores of its last pass
cudaError_t ntt_reorder_and_dft_scatter(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding_factor,
                                        uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_work, uint32_t* const* peers,
                                        uint32_t world, uint32_t rank, const uint32_t* d_tw, unsigned tw_log_n,
                                        uint32_t col_begin = 0, uint32_t col_count = 0);  // column range (0, 0 = all)
// layers [l_first, log_h) on the rows a rank holds after the exchange of the row-sharded commit
cudaError_t ntt_layers_mapped(cudaStream_t stream, uint32_t* d_mat, uint64_t w, unsigned log_h, unsigned l_first,
                              uint64_t n_blocks, uint64_t run, uint64_t block, uint64_t offset, const uint32_t* d_tw,
                              unsigned tw_log_n, uint64_t col_begin = 0, uint64_t col_count = 0,  // column range (0, 0 = all)
                              uint32_t* d_out = nullptr);                                          // result matrix (null = in place)
}  // namespace lm

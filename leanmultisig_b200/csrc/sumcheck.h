// Internal (C++) launch interface of sumcheck.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace lm {
// w[(selector << m) + x] += scalar * eq(point, x);  point: m x 5 words on the device
size_t weights_add_eq_scratch_words(uint32_t m);
cudaError_t weights_add_eq(cudaStream_t stream, uint32_t* d_w, uint64_t selector, const uint32_t* d_point, uint32_t m,
                           const uint32_t scalar[5], uint32_t* d_scratch);
// w[(selector << m) + x] += scalar * next_mle(point, x)
cudaError_t weights_add_next(cudaStream_t stream, uint32_t* d_w, uint64_t selector, const uint32_t* d_point, uint32_t m,
                             const uint32_t scalar[5]);
// w[x] += sum_q scalars[q] * eq(points[q], x), base-field points (n_q x m words), w has 2^m entries
size_t weights_add_base_eq_scratch_words(uint32_t m, uint32_t n_q);
cudaError_t weights_add_base_eq(cudaStream_t stream, uint32_t* d_w, uint32_t m, const uint32_t* d_points, uint32_t n_q,
                                const uint32_t* d_scalars, uint32_t* d_scratch);
// d_out10 = (c0, c2) of one product-sumcheck round over n entries; p has `dim` words per entry, entries >= live are 0
size_t prod_round_scratch_words();
cudaError_t prod_round(cudaStream_t stream, const uint32_t* d_p, uint32_t dim, uint64_t live, const uint32_t* d_w, uint64_t n,
                       uint32_t* d_scratch, uint32_t* d_out10);
// fold both tables with r (n -> n/2 EF entries; in place allowed for EF tables) and compute the next (c0, c2)
cudaError_t prod_fold_round(cudaStream_t stream, const uint32_t* d_p, uint32_t dim, uint64_t live, const uint32_t* d_w,
                            uint64_t n, const uint32_t r[5], uint32_t* d_p_out, uint32_t* d_w_out, uint32_t* d_scratch,
                            uint32_t* d_out10);
cudaError_t weights_add_strided_eq(cudaStream_t stream, uint32_t* d_w, uint64_t base, uint32_t shift, uint64_t offset,
                                   const uint32_t* d_point, uint32_t pre, const uint32_t scalar[5]);
// K <= weights_add_eq_batch_max(m) statements per call (12 on the tensor-core path, whose accumulator bounds hold for up to 240
// input bytes per row; 16 on the scalar path)
uint32_t weights_add_eq_batch_max(uint32_t m);
size_t weights_add_eq_batch_scratch_words(uint32_t m, uint32_t K);
cudaError_t weights_add_eq_batch(cudaStream_t stream, uint32_t* d_w, uint64_t selector, const uint32_t* d_points, uint32_t m,
                                 const uint32_t* scalars, uint32_t K, uint32_t* d_scratch);
// CPU model of the tensor-core statement-weights kernel (test hook): w[(x_hi << lo_vars) + x_lo] += sum_k hi_k[x_hi] lo_k[x_lo]
void weights_gemm_model_host(uint32_t* w, const uint32_t* hi, const uint32_t* lo, int K, int hi_vars, int lo_vars);
}  // namespace lm

// Internal (C++) launch interface of merkle.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace lm {
// digests[r] = sponge(row r zero-extended to full_w); rows are stored_w wide, entries >= eff_w are zero.
cudaError_t merkle_leaf_digests(cudaStream_t stream, const uint32_t* d_mat, uint64_t h, uint32_t stored_w,
                                uint32_t full_w, uint32_t eff_w, uint32_t* d_digests);
// chunk-at-a-time variant of the leaf sponge (columns hashed as soon as they exist, right to left)
bool merkle_leaf_chunked_ok(uint32_t stored_w, uint32_t full_w, uint32_t eff_w);
cudaError_t merkle_leaf_absorb_chunks(cudaStream_t stream, const uint32_t* d_mat, uint64_t h, uint32_t stored_w,
                                      uint32_t full_w, uint32_t eff_w, uint32_t chunk_hi, uint32_t count, uint32_t* d_digests);
// d_layers holds (2h - 1) digests: layer 0 (h leaf digests, already filled) then h/2, ..., 1.
cudaError_t merkle_tree_from_digests(cudaStream_t stream, uint32_t* d_layers, uint64_t h);
// rows (n x full_w, zero-extended) and sibling paths (n x log2(h) x 8, leaf level first) of n leaf indices
cudaError_t merkle_open_gather(cudaStream_t stream, const uint32_t* d_mat, const uint32_t* d_layers, uint64_t h,
                               uint32_t stored_w, uint32_t full_w, const uint64_t* d_indices, uint32_t n,
                               uint32_t* d_rows, uint32_t* d_paths);
// verifier side: ok[q] = 1 iff row q (width words) hashes, through its sibling path (log_h x 8 words, leaf level first), to root
cudaError_t merkle_verify_openings(cudaStream_t stream, const uint32_t* d_root, uint32_t log_h, const uint64_t* d_indices,
                                   uint32_t n, const uint32_t* d_rows, uint32_t width, const uint32_t* d_paths, uint32_t* d_ok);
// n explicit 16-word states: permutation (compress = 0) or permutation + feed-forward (compress = 1)
cudaError_t poseidon1_states(cudaStream_t stream, uint32_t* d_states, uint64_t n, int compress);
// smallest w >= start whose PoW check passes for the challenger state `state` (host, 16 words); d_best: device u64
cudaError_t pow_grind(cudaStream_t stream, const uint32_t state[16], uint32_t bits, uint64_t start,
                      unsigned long long* d_best, uint64_t* witness);
// one Poseidon1 permutation on the host (Fiat-Shamir transcript sponge)
void poseidon1_permute_host(uint32_t state[16]);
// the same permutation through the CPU model of the tensor-core formulation (test hook for the CPU tier)
void poseidon1_permute_umma_model_host(uint32_t state[16]);
// the B-matrix image of that formulation; returns its size (copies when `capacity` suffices)
size_t poseidon1_umma_image_host(uint8_t* out, size_t capacity);
}  // namespace lm

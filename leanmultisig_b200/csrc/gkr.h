// Internal (C++) launch interface of gkr.cu; the public C ABI is include/leanmultisig_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "devfs.cuh"
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
// d_dens: coefficient plane 0 at the section's first row; plane k is den_stride words further
cudaError_t logup_fill_section(cudaStream_t stream, const LogupSection& section, uint32_t* d_nums, uint32_t* d_dens,
                               uint64_t den_stride);
cudaError_t finger_print(cudaStream_t stream, const uint32_t* d_data, uint64_t n_rows, uint32_t n_data,
                         const uint32_t* d_alphas, const uint32_t c[5], uint32_t* d_out);

// ---- quotient GKR -------------------------------------------------------------------------------------------
// Layout of a layer with n fractions: denominators as five coefficient planes u32[5][n]; numerators u32[n] (layer 0,
// base field) or planes u32[5][n].  Working table of a layer sumcheck: planes u32[20][rows], plane 5 c + k of column
// c in (nl, nr, dl, dr).
constexpr int GKR_MAX_VARS = 40;
struct GkrDev {  // per-session state of the layer sumcheck in device memory
  Ef point[GKR_MAX_VARS];  // claim point of the current layer (k coordinates)
  Ef q[GKR_MAX_VARS];      // challenges of the current layer, round order
  Ef inv_point[GKR_MAX_VARS];  // inverses of the claim point's coordinates, computed in parallel when the layer begins
  Ef claim_num, claim_den, alpha, s, mmf, r;
  Ef inner[4];             // (nl, nr, dl, dr) at the end of the layer
  Ef out[2];               // (c0, c2) of the last round when the transcript is driven by the host
  uint32_t k;              // claim variables of the current layer
  uint32_t counter;        // ticket of the last-block reduction
};
size_t gkr_eq_table_words(uint32_t max_claim_vars);  // prefix eq tables of one layer
constexpr int GKR_MAX_BLOCKS = 148 * 4;
constexpr int GKR_TAIL_VARS = 9;  // layers / remaining rounds with <= 2^9 row pairs run in one CTA

cudaError_t gkr_pad(cudaStream_t stream, uint32_t* d_nums, uint32_t* d_dens, uint64_t active, uint64_t n);
cudaError_t gkr_aos_to_planes(cudaStream_t stream, const uint32_t* d_aos, uint64_t count, uint64_t stride, uint32_t* d_planes);
cudaError_t gkr_layer_up(cudaStream_t stream, const uint32_t* d_nums, uint32_t num_dim, const uint32_t* d_dens, uint64_t n,
                         uint32_t* d_out_nums, uint32_t* d_out_dens);

struct GkrLayerArgs {
  const uint32_t* nums;  // layer below the claim: 2^(k+1) fractions
  const uint32_t* dens;
  uint32_t num_dim;
  uint32_t k;            // claim variables = rounds of this layer
  uint32_t* w[2];        // ping-pong working tables (w[0]: 2^(k-1) rows, w[1]: 2^(k-2) rows)
  uint32_t* eq_tab;
  uint32_t* partial;     // GKR_MAX_BLOCKS x 10 words
  GkrDev* g;
  DevFs* fs;             // nullptr: transcript driven by the host
  uint32_t* tr;
};
// host-driven pieces (lm_gkr_layer_begin / round / fold / layer_end): alpha, r and the point are already in GkrDev
cudaError_t gkr_begin(cudaStream_t stream, const GkrLayerArgs& a, const uint32_t eq_scale[5], bool sample_alpha);
// round `rnd` (0-based): folds the previous table with g->r on the fly when rnd > 0; results to g->out (host-driven) or
// into the transcript (device-driven)
cudaError_t gkr_round(cudaStream_t stream, const GkrLayerArgs& a, uint32_t rnd);
cudaError_t gkr_end(cudaStream_t stream, const GkrLayerArgs& a);  // last fold -> g->inner (+ layer-end transcript step)
// device-driven: every round of the layer, the layer-end step and (unless last) the begin step of the next layer
cudaError_t gkr_layer_device(cudaStream_t stream, const GkrLayerArgs& a, const GkrLayerArgs* next);
}  // namespace lm

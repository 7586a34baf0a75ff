// Logup quotient GKR on sm_100a: fingerprints, layer-up pass, per-layer degree-3 sumcheck rounds.
//
// Device replacement for
//   crates/utils/src/multilinear.rs:76-98                               finger_print(_packed)
//   crates/sub_protocols/src/quotient_gkr/layers.rs:124-189             sum_quotients_2_by_2(_packed_br)
//   crates/sub_protocols/src/quotient_gkr/sumcheck_utils.rs:65-79       pair_coeffs
//   crates/sub_protocols/src/quotient_gkr/sumcheck_utils.rs:112-359     the three-phase layer sumcheck (round bodies)
//   crates/sub_protocols/src/quotient_gkr/sumcheck_utils.rs:384-489     compute_round_packed / fold_and_compute_round_packed
//
// Natural order instead of the reference's chunk-bit-reversed SIMD layout: a layer of 2^(K+1) fractions is the
// interleaving of its even ("l") and odd ("r") halves, the layer above is (nl dr + nr dl, dl dr), and the
// per-layer sumcheck binds the least-significant variable first, so round 0 of a layer reads four consecutive
// fractions per thread and every later round reads two consecutive 80-byte rows of the working table
// W[row] = (nl, nr, dl, dr).  Entries past the active prefix are materialised as (0, 1).
#include <cuda_runtime.h>
#include <cstdint>
#include "gkr.h"
#include "kb.cuh"
#include "launch_count.h"
#include "poly.h"
#include "reduce.cuh"

namespace lm {

constexpr int GKR_LO = 10;

// out[r] = c - sum_i alphas[i] * data[r][i]                      (multilinear.rs:76-86)
__global__ void finger_print_kernel(const uint32_t* __restrict__ data, uint64_t n_rows, int n_data,
                                    const uint32_t* __restrict__ alphas, Ef c, uint32_t* __restrict__ out) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  uint64_t acc[5] = {0, 0, 0, 0, 0};
  int terms = 0;
  for (int i = 0; i < n_data; i++) {
    const uint32_t f = __ldg(data + r * n_data + i);
    if (terms == 3) {
#pragma unroll
      for (int k = 0; k < 5; k++) acc[k] = kb_fold(acc[k]);
      terms = 0;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) acc[k] = mad_wide(f, __ldg(alphas + 5 * i + k), acc[k]);
    terms++;
  }
  Ef o;
#pragma unroll
  for (int k = 0; k < 5; k++) o.c[k] = kb_sub(c.c[k], kb_canon(kb_redc_lazy(kb_fold(acc[k]))));
  st_ef(out + 5 * r, o);
}

cudaError_t finger_print(cudaStream_t stream, const uint32_t* d_data, uint64_t n_rows, uint32_t n_data,
                         const uint32_t* d_alphas, const uint32_t c[5], uint32_t* d_out) {
  if (n_rows == 0) return cudaSuccess;
  Ef cc;
  for (int k = 0; k < 5; k++) cc.c[k] = c[k];
  finger_print_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(d_data, n_rows, (int)n_data, d_alphas, cc, d_out);
  count_launch();
  return cudaGetLastError();
}

// ---- Logup table build (prove_generic_logup, crates/sub_protocols/src/logup.rs:52-211) --------------------------
// One section = n_rows consecutive (numerator, denominator) pairs in NATURAL row order:
//   numerator   = 1 | +col[r] | -col[r] | 0
//   denominator = c + sign * (contrib + sum_i alphas[i] * data_i[r])      (finger_print_packed, multilinear.rs:87-98)
//                 or the constant 1 for padding
// data_i[r] = col[offset + r * stride] (+ add) | r | constant.  Everything a section needs travels in the kernel
// parameter (constant bank); rows are independent and the writes are coalesced (4 B + 20 B per row).
__global__ void logup_fill_kernel(const __grid_constant__ LogupSection S, uint32_t* __restrict__ nums, uint32_t* __restrict__ dens) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S.n_rows) return;
  uint32_t num;
  switch (S.num_mode) {
    case LOGUP_NUM_ONE: num = KB_R1; break;
    case LOGUP_NUM_COL: num = __ldg(S.num_col + r); break;
    case LOGUP_NUM_NEG_COL: num = kb_neg(__ldg(S.num_col + r)); break;
    default: num = 0; break;
  }
  nums[r] = num;
  Ef o;
  if (S.den_sign == 0) {
    o = ef_from_base(KB_R1);
  } else {
    uint64_t acc[5] = {0, 0, 0, 0, 0};
    int terms = 0;
    for (int i = 0; i < S.n_data; i++) {
      const LogupData& d = S.data[i];
      uint32_t f;
      if (d.kind == LOGUP_DATA_COL)
        f = kb_add(__ldg(d.col + d.offset + r * d.stride), d.add);
      else if (d.kind == LOGUP_DATA_ROW)
        f = kb_mul((uint32_t)r, KB_R2);  // Montgomery form of the row index (< 2^31 rows per section)
      else
        f = d.add;
      if (terms == 3) {
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k] = kb_fold(acc[k]);
        terms = 0;
      }
#pragma unroll
      for (int k = 0; k < 5; k++) acc[k] = mad_wide(f, S.alphas[i].c[k], acc[k]);
      terms++;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint32_t fp = kb_add(S.contrib.c[k], kb_canon(kb_redc_lazy(kb_fold(acc[k]))));
      o.c[k] = S.den_sign > 0 ? kb_add(S.c.c[k], fp) : kb_sub(S.c.c[k], fp);
    }
  }
  st_ef(dens + 5 * r, o);
}

cudaError_t logup_fill_section(cudaStream_t stream, const LogupSection& S, uint32_t* d_nums, uint32_t* d_dens) {
  if (S.n_rows == 0) return cudaSuccess;
  if (S.n_data > LOGUP_MAX_DATA || S.n_rows > ((uint64_t)1 << 31)) return cudaErrorInvalidValue;
  logup_fill_kernel<<<(unsigned)((S.n_rows + 255) / 256), 256, 0, stream>>>(S, d_nums, d_dens);
  count_launch();
  return cudaGetLastError();
}

// nums[i] = 0, dens[i] = 1 for i in [active, n)
__global__ void gkr_pad_kernel(uint32_t* nums, int num_dim, uint32_t* dens, uint64_t active, uint64_t n) {
  const uint64_t i = active + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < num_dim; k++) nums[i * num_dim + k] = 0;
  dens[5 * i] = KB_R1;
  for (int k = 1; k < 5; k++) dens[5 * i + k] = 0;
}

template <int NUM_DIM>
__global__ void gkr_layer_up_kernel(const uint32_t* __restrict__ nums, const uint32_t* __restrict__ dens, uint64_t half,
                                    uint32_t* __restrict__ out_nums, uint32_t* __restrict__ out_dens) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const Ef d0 = ld_ef(dens + 5 * (2 * i)), d1 = ld_ef(dens + 5 * (2 * i + 1));
  Ef n;
  if (NUM_DIM == 1) {
    const uint2 nn = *reinterpret_cast<const uint2*>(nums + 2 * i);
    n = ef_add(ef_mul_base(d1, nn.x), ef_mul_base(d0, nn.y));
  } else {
    n = ef_add(ef_mul(d1, ld_ef(nums + 5 * (2 * i))), ef_mul(d0, ld_ef(nums + 5 * (2 * i + 1))));
  }
  st_ef(out_nums + 5 * i, n);
  st_ef(out_dens + 5 * i, ef_mul(d0, d1));
}

cudaError_t gkr_pad(cudaStream_t stream, uint32_t* d_nums, uint32_t num_dim, uint32_t* d_dens, uint64_t active, uint64_t n) {
  if (active >= n) return cudaSuccess;
  gkr_pad_kernel<<<(unsigned)((n - active + 255) / 256), 256, 0, stream>>>(d_nums, (int)num_dim, d_dens, active, n);
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_layer_up(cudaStream_t stream, const uint32_t* d_nums, uint32_t num_dim, const uint32_t* d_dens, uint64_t n,
                         uint32_t* d_out_nums, uint32_t* d_out_dens) {
  const uint64_t half = n / 2;
  if (half == 0) return cudaErrorInvalidValue;
  if (num_dim == 1)
    gkr_layer_up_kernel<1><<<(unsigned)((half + 255) / 256), 256, 0, stream>>>(d_nums, d_dens, half, d_out_nums, d_out_dens);
  else
    gkr_layer_up_kernel<5><<<(unsigned)((half + 255) / 256), 256, 0, stream>>>(d_nums, d_dens, half, d_out_nums, d_out_dens);
  count_launch();
  return cudaGetLastError();
}

// ---- layer sumcheck -------------------------------------------------------------------------------------
struct Row4 {
  Ef nl, nr, dl, dr;
};
// row `row` of the 4 working columns; SRC 0: straight from the layer arrays (nl = nums[2 row], nr = nums[2 row + 1], ..)
template <int SRC, int NUM_DIM>
__device__ __forceinline__ Row4 ld_row(const uint32_t* a, const uint32_t* b, uint64_t row) {
  Row4 r;
  if (SRC == 0) {
    if (NUM_DIM == 1) {
      const uint2 nn = *reinterpret_cast<const uint2*>(a + 2 * row);
      r.nl = ef_from_base(nn.x);
      r.nr = ef_from_base(nn.y);
    } else {
      r.nl = ld_ef(a + 5 * (2 * row));
      r.nr = ld_ef(a + 5 * (2 * row + 1));
    }
    r.dl = ld_ef(b + 5 * (2 * row));
    r.dr = ld_ef(b + 5 * (2 * row + 1));
  } else {
    r.nl = ld_ef_rw(a + 20 * row);
    r.nr = ld_ef_rw(a + 20 * row + 5);
    r.dl = ld_ef_rw(a + 20 * row + 10);
    r.dr = ld_ef_rw(a + 20 * row + 15);
  }
  return r;
}
// G(nl, nr, dl, dr) = nl dr + nr dl + alpha dl dr
__device__ __forceinline__ Ef gkr_g(const Row4& v, const Ef& alpha) {
  return ef_add(ef_add(ef_mul(v.nl, v.dr), ef_mul(v.nr, v.dl)), ef_mul(alpha, ef_mul(v.dl, v.dr)));
}

template <int SRC, int NUM_DIM>
__global__ void __launch_bounds__(256)
gkr_round_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t half, const uint32_t* __restrict__ eq_hi,
                 const uint32_t* __restrict__ eq_lo, int lo_vars, Ef alpha, uint32_t* __restrict__ partial) {
  Ef c0 = ef_zero(), c2 = ef_zero();
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += (uint64_t)gridDim.x * blockDim.x) {
    const Row4 lo = ld_row<SRC, NUM_DIM>(a, b, 2 * j), hi = ld_row<SRC, NUM_DIM>(a, b, 2 * j + 1);
    Row4 df;
    df.nl = ef_sub(hi.nl, lo.nl), df.nr = ef_sub(hi.nr, lo.nr), df.dl = ef_sub(hi.dl, lo.dl), df.dr = ef_sub(hi.dr, lo.dr);
    const Ef eq = ef_mul(ld_ef(eq_hi + 5 * (j >> lo_vars)), ld_ef(eq_lo + 5 * (j & (((uint64_t)1 << lo_vars) - 1))));
    c0 = ef_add(c0, ef_mul(eq, gkr_g(lo, alpha)));
    c2 = ef_add(c2, ef_mul(eq, gkr_g(df, alpha)));
  }
  block_reduce_pair(c0, c2, partial);
}

template <int SRC, int NUM_DIM>
__global__ void gkr_fold_kernel(const uint32_t* a, const uint32_t* b, uint64_t half, Ef r, uint32_t* out) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= half) return;
  const Row4 lo = ld_row<SRC, NUM_DIM>(a, b, 2 * j), hi = ld_row<SRC, NUM_DIM>(a, b, 2 * j + 1);
  st_ef(out + 20 * j, ef_add(lo.nl, ef_mul(r, ef_sub(hi.nl, lo.nl))));
  st_ef(out + 20 * j + 5, ef_add(lo.nr, ef_mul(r, ef_sub(hi.nr, lo.nr))));
  st_ef(out + 20 * j + 10, ef_add(lo.dl, ef_mul(r, ef_sub(hi.dl, lo.dl))));
  st_ef(out + 20 * j + 15, ef_add(lo.dr, ef_mul(r, ef_sub(hi.dr, lo.dr))));
}

size_t gkr_round_scratch_words(uint32_t n_vars) {
  const uint32_t lv = n_vars ? n_vars - 1 : 0;
  const int lo = lv < (uint32_t)GKR_LO ? (int)lv : GKR_LO;
  return 5 * (((size_t)1 << (lv - lo)) + ((size_t)1 << lo)) + 10 * (148 * 8) + 64;
}

// One round over `n_rows` rows of the 4 working columns.  src == 0: a = layer nums, b = layer dens (2 n_rows entries);
// src == 1: a = working table W (n_rows x 20 words).  d_eq_point: log2(n_rows) - 1 EF entries.
cudaError_t gkr_round(cudaStream_t stream, int src, uint32_t num_dim, const uint32_t* a, const uint32_t* b, uint32_t n_vars,
                      const uint32_t* d_eq_point, const uint32_t alpha[5], uint32_t* d_scratch, uint32_t* d_out10,
                      const uint32_t* eq_scale) {
  if (n_vars < 1) return cudaErrorInvalidValue;
  Ef al;
  for (int k = 0; k < 5; k++) al.c[k] = alpha[k];
  const uint64_t half = (uint64_t)1 << (n_vars - 1);
  const uint32_t lv = n_vars - 1;
  const int lo_vars = lv < (uint32_t)GKR_LO ? (int)lv : GKR_LO;
  const int hi_vars = (int)lv - lo_vars;
  uint32_t* d_hi = d_scratch;
  uint32_t* d_lo = d_hi + 5 * ((size_t)1 << hi_vars);
  uint32_t* d_part = d_lo + 5 * ((size_t)1 << lo_vars);
  const uint32_t one[5] = {KB_R1, 0, 0, 0, 0};
  cudaError_t e;
  if ((e = eq_table(stream, d_eq_point, hi_vars, eq_scale ? eq_scale : one, d_hi)) != cudaSuccess) return e;
  if ((e = eq_table(stream, d_eq_point + 5 * hi_vars, lo_vars, one, d_lo)) != cudaSuccess) return e;
  uint64_t blocks = (half + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const unsigned g = (unsigned)blocks;
  if (src == 0 && num_dim == 1)
    gkr_round_kernel<0, 1><<<g, 256, 0, stream>>>(a, b, half, d_hi, d_lo, lo_vars, al, d_part);
  else if (src == 0)
    gkr_round_kernel<0, 5><<<g, 256, 0, stream>>>(a, b, half, d_hi, d_lo, lo_vars, al, d_part);
  else
    gkr_round_kernel<1, 5><<<g, 256, 0, stream>>>(a, b, half, d_hi, d_lo, lo_vars, al, d_part);
  count_launch();
  sum_pair_partials_kernel<<<1, 256, 0, stream>>>(d_part, (int)g, d_out10);
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_fold(cudaStream_t stream, int src, uint32_t num_dim, const uint32_t* a, const uint32_t* b, uint32_t n_vars,
                     const uint32_t r[5], uint32_t* d_out) {
  if (n_vars < 1) return cudaErrorInvalidValue;
  Ef rr;
  for (int k = 0; k < 5; k++) rr.c[k] = r[k];
  const uint64_t half = (uint64_t)1 << (n_vars - 1);
  const unsigned g = (unsigned)((half + 127) / 128);
  if (src == 0 && num_dim == 1)
    gkr_fold_kernel<0, 1><<<g, 128, 0, stream>>>(a, b, half, rr, d_out);
  else if (src == 0)
    gkr_fold_kernel<0, 5><<<g, 128, 0, stream>>>(a, b, half, rr, d_out);
  else
    gkr_fold_kernel<1, 5><<<g, 128, 0, stream>>>(a, b, half, rr, d_out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace lm

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
// fractions per thread.  Layers and working tables are stored as coefficient planes (u32[5][n] per EF array), so a
// thread's 16-byte loads of consecutive rows are coalesced across the warp.  Entries past the active prefix are
// materialised as (0, 1).  The challenger can run on the device (devfs.cuh): then a layer is a sequence of kernel
// launches with no host synchronisation, and everything below 2^9 row pairs is one CTA.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "gkr.h"
#include "kb.cuh"
#include "launch_count.h"
#include "poly.h"
#include "eqtab.cuh"
#include "reduce.cuh"

namespace lm {

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
// parameter (constant bank); rows are independent and the writes are coalesced (the numerator column and the five
// coefficient planes of the denominators, 4 B each per row).
__global__ void logup_fill_kernel(const __grid_constant__ LogupSection S, uint32_t* __restrict__ nums, uint32_t* __restrict__ dens,
                                  uint64_t den_stride) {
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
#pragma unroll
  for (int k = 0; k < 5; k++) dens[k * den_stride + r] = o.c[k];  // coefficient planes (the GKR layer layout, see below)
}

cudaError_t logup_fill_section(cudaStream_t stream, const LogupSection& S, uint32_t* d_nums, uint32_t* d_dens, uint64_t den_stride) {
  if (S.n_rows == 0) return cudaSuccess;
  if (S.n_data > LOGUP_MAX_DATA || S.n_rows > ((uint64_t)1 << 31)) return cudaErrorInvalidValue;
  logup_fill_kernel<<<(unsigned)((S.n_rows + 255) / 256), 256, 0, stream>>>(S, d_nums, d_dens, den_stride);
  count_launch();
  return cudaGetLastError();
}

// ---- layers ----------------------------------------------------------------------------------------------
// nums[i] = 0, dens[i] = 1 for i in [active, n) of layer 0
__global__ void gkr_pad_kernel(uint32_t* nums, uint32_t* dens, uint64_t active, uint64_t n) {
  const uint64_t i = active + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  nums[i] = 0;
  dens[i] = KB_R1;
  for (int k = 1; k < 5; k++) dens[k * n + i] = 0;
}

__global__ void gkr_aos_to_planes_kernel(const uint32_t* __restrict__ aos, uint64_t count, uint64_t stride,
                                         uint32_t* __restrict__ planes) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
#pragma unroll
  for (int k = 0; k < 5; k++) planes[k * stride + i] = __ldg(aos + 5 * i + k);
}

__device__ __forceinline__ uint2 ldg2(const uint32_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ uint4 ldg4(const uint32_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
// tables another kernel (or another CTA of this kernel, behind a barrier) has just written: L2, not the read-only path
__device__ __forceinline__ uint4 ldcg4(const uint32_t* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint2 ldcg2(const uint32_t* p) { return __ldcg(reinterpret_cast<const uint2*>(p)); }

// (n, d)' = (n0 d1 + n1 d0, d0 d1) on adjacent fractions (layers.rs:124-189); planes in, planes out
template <int NUM_DIM>
__global__ void __launch_bounds__(256)
gkr_layer_up_kernel(const uint32_t* __restrict__ nums, const uint32_t* __restrict__ dens, uint64_t n_in,
                    uint32_t* __restrict__ out_nums, uint32_t* __restrict__ out_dens) {
  const uint64_t half = n_in / 2;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  Ef d0, d1;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint2 v = ldg2(dens + k * n_in + 2 * i);
    d0.c[k] = v.x, d1.c[k] = v.y;
  }
  Ef n;
  if (NUM_DIM == 1) {
    const uint2 nn = ldg2(nums + 2 * i);
    n = ef_add(ef_mul_base(d1, nn.x), ef_mul_base(d0, nn.y));
  } else {
    Ef n0, n1;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint2 v = ldg2(nums + k * n_in + 2 * i);
      n0.c[k] = v.x, n1.c[k] = v.y;
    }
    n = ef_mul2_add(n0, d1, n1, d0);
  }
  const Ef d = ef_mul(d0, d1);
#pragma unroll
  for (int k = 0; k < 5; k++) out_nums[k * half + i] = n.c[k], out_dens[k * half + i] = d.c[k];
}

cudaError_t gkr_pad(cudaStream_t stream, uint32_t* d_nums, uint32_t* d_dens, uint64_t active, uint64_t n) {
  if (active >= n) return cudaSuccess;
  gkr_pad_kernel<<<(unsigned)((n - active + 255) / 256), 256, 0, stream>>>(d_nums, d_dens, active, n);
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_aos_to_planes(cudaStream_t stream, const uint32_t* d_aos, uint64_t count, uint64_t stride, uint32_t* d_planes) {
  if (count == 0) return cudaSuccess;
  gkr_aos_to_planes_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(d_aos, count, stride, d_planes);
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_layer_up(cudaStream_t stream, const uint32_t* d_nums, uint32_t num_dim, const uint32_t* d_dens, uint64_t n,
                         uint32_t* d_out_nums, uint32_t* d_out_dens) {
  const uint64_t half = n / 2;
  if (half == 0) return cudaErrorInvalidValue;
  if (num_dim == 1)
    gkr_layer_up_kernel<1><<<(unsigned)((half + 255) / 256), 256, 0, stream>>>(d_nums, d_dens, n, d_out_nums, d_out_dens);
  else
    gkr_layer_up_kernel<5><<<(unsigned)((half + 255) / 256), 256, 0, stream>>>(d_nums, d_dens, n, d_out_nums, d_out_dens);
  count_launch();
  return cudaGetLastError();
}

// ---- layer sumcheck -------------------------------------------------------------------------------------
// Round `rnd` of the layer with k claim variables works on 2^(k - rnd) rows of the four columns (nl, nr, dl, dr):
// pair j = rows (2j, 2j + 1), weight eq(point[0 .. m), j) with m = k - 1 - rnd.  The fold of the previous round's
// challenge is fused into the round that consumes it (fold_and_compute_round_packed, sumcheck_utils.rs:426-489):
// a thread folds four old rows into the two it needs, stores them for the next round and evaluates on them, so every
// table is read once and written at half size.  MODE 0: rows straight from the layer arrays (round 0); MODE 1: layer
// rows folded with r (round 1); MODE 2: working-table rows folded with r (round >= 2).
struct Row4 {
  Ef nl, nr, dl, dr;
};

size_t gkr_eq_table_words(uint32_t max_claim_vars) { return eqtab_words(max_claim_vars); }
__device__ __forceinline__ void gkr_build_tables(uint32_t* tab, const GkrDev* g, uint32_t k, const Ef& scale) {
  eqtab_build(tab, g->point, k, scale);
}

__device__ __forceinline__ Ef fold1(const Ef& r, uint32_t lo, uint32_t hi) {  // lo + r (hi - lo), base values
  return ef_add_base(ef_mul_base(r, kb_sub(hi, lo)), lo);
}
__device__ __forceinline__ Ef fold5(const Ef& r, const Ef& lo, const Ef& hi) { return ef_add(lo, ef_mul(r, ef_sub(hi, lo))); }

template <int MODE, int NUM_DIM>
__device__ __forceinline__ void gkr_rows(const GkrLayerArgs& A, uint32_t rnd, uint64_t j, const Ef& r, Row4& a, Row4& b) {
  const uint64_t n = (uint64_t)2 << A.k;  // fractions of the layer
  if (MODE == 0) {
    if (NUM_DIM == 1) {
      const uint4 v = ldg4(A.nums + 4 * j);
      a.nl = ef_from_base(v.x), a.nr = ef_from_base(v.y), b.nl = ef_from_base(v.z), b.nr = ef_from_base(v.w);
    } else {
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const uint4 v = ldg4(A.nums + k * n + 4 * j);
        a.nl.c[k] = v.x, a.nr.c[k] = v.y, b.nl.c[k] = v.z, b.nr.c[k] = v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint4 v = ldg4(A.dens + k * n + 4 * j);
      a.dl.c[k] = v.x, a.dr.c[k] = v.y, b.dl.c[k] = v.z, b.dr.c[k] = v.w;
    }
    return;
  }
  uint32_t* wn = A.w[(rnd - 1) & 1];
  const uint64_t rows_new = (uint64_t)1 << (A.k - rnd);
  if (MODE == 1) {
    // old rows 4j .. 4j+3 = fractions 8j .. 8j+7
    if (NUM_DIM == 1) {
      const uint4 v0 = ldg4(A.nums + 8 * j), v1 = ldg4(A.nums + 8 * j + 4);
      a.nl = fold1(r, v0.x, v0.z), a.nr = fold1(r, v0.y, v0.w);
      b.nl = fold1(r, v1.x, v1.z), b.nr = fold1(r, v1.y, v1.w);
    } else {
      Ef o[8];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const uint4 v0 = ldg4(A.nums + k * n + 8 * j), v1 = ldg4(A.nums + k * n + 8 * j + 4);
        o[0].c[k] = v0.x, o[1].c[k] = v0.y, o[2].c[k] = v0.z, o[3].c[k] = v0.w;
        o[4].c[k] = v1.x, o[5].c[k] = v1.y, o[6].c[k] = v1.z, o[7].c[k] = v1.w;
      }
      a.nl = fold5(r, o[0], o[2]), a.nr = fold5(r, o[1], o[3]);
      b.nl = fold5(r, o[4], o[6]), b.nr = fold5(r, o[5], o[7]);
    }
    {
      Ef o[8];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const uint4 v0 = ldg4(A.dens + k * n + 8 * j), v1 = ldg4(A.dens + k * n + 8 * j + 4);
        o[0].c[k] = v0.x, o[1].c[k] = v0.y, o[2].c[k] = v0.z, o[3].c[k] = v0.w;
        o[4].c[k] = v1.x, o[5].c[k] = v1.y, o[6].c[k] = v1.z, o[7].c[k] = v1.w;
      }
      a.dl = fold5(r, o[0], o[2]), a.dr = fold5(r, o[1], o[3]);
      b.dl = fold5(r, o[4], o[6]), b.dr = fold5(r, o[5], o[7]);
    }
  } else {
    const uint32_t* wo = A.w[rnd & 1];
    const uint64_t rows_old = rows_new * 2;
    Ef* dst_a[4] = {&a.nl, &a.nr, &a.dl, &a.dr};
    Ef* dst_b[4] = {&b.nl, &b.nr, &b.dl, &b.dr};
#pragma unroll
    for (int c = 0; c < 4; c++) {
      Ef o0, o1, o2, o3;
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const uint4 v = ldcg4(wo + (uint64_t)(5 * c + k) * rows_old + 4 * j);
        o0.c[k] = v.x, o1.c[k] = v.y, o2.c[k] = v.z, o3.c[k] = v.w;
      }
      *dst_a[c] = fold5(r, o0, o1);
      *dst_b[c] = fold5(r, o2, o3);
    }
  }
  // the folded rows are the next round's input
  const Ef* src_a[4] = {&a.nl, &a.nr, &a.dl, &a.dr};
  const Ef* src_b[4] = {&b.nl, &b.nr, &b.dl, &b.dr};
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int k = 0; k < 5; k++)
      *reinterpret_cast<uint2*>(wn + (uint64_t)(5 * c + k) * rows_new + 2 * j) = make_uint2(src_a[c]->c[k], src_b[c]->c[k]);
}

// G(nl, nr, dl, dr) = nl dr + nr dl + alpha dl dr
template <bool BASE_NUM>
__device__ __forceinline__ Ef gkr_g(const Row4& v, const Ef& alpha) {
  Ef t;
  if (BASE_NUM)
    t = ef_add(ef_mul_base(v.dr, v.nl.c[0]), ef_mul_base(v.dl, v.nr.c[0]));
  else
    t = ef_mul2_add(v.nl, v.dr, v.nr, v.dl);
  return ef_add(t, ef_mul(alpha, ef_mul(v.dl, v.dr)));
}

// sums of this thread's pairs: c0 += eq G(row 2j), c2 += eq G(row 2j+1 - row 2j)
template <int MODE, int NUM_DIM>
__device__ __forceinline__ void gkr_accumulate(const GkrLayerArgs& A, uint32_t rnd, uint64_t j0, uint64_t stride, Ef& c0, Ef& c2) {
  const uint32_t m = A.k - 1 - rnd;
  const uint64_t half = (uint64_t)1 << m;
  const EqView eqv(A.eq_tab, A.k, m);
  const Ef alpha = A.g->alpha;
  Ef r = ef_zero();
  if (MODE != 0) r = A.g->r;
  for (uint64_t j = j0; j < half; j += stride) {
    Row4 a, b;
    gkr_rows<MODE, NUM_DIM>(A, rnd, j, r, a, b);
    Row4 df;
    df.nl = ef_sub(b.nl, a.nl), df.nr = ef_sub(b.nr, a.nr), df.dl = ef_sub(b.dl, a.dl), df.dr = ef_sub(b.dr, a.dr);
    const Ef eq = eqv(j);
    constexpr bool BASE = MODE == 0 && NUM_DIM == 1;
    c0 = ef_add(c0, ef_mul(eq, gkr_g<BASE>(a, alpha)));
    c2 = ef_add(c2, ef_mul(eq, gkr_g<BASE>(df, alpha)));
  }
}

// CTA reduction of (c0, c2); the result is in red[0], red[1] (shared) for every thread after the trailing barrier
__device__ __forceinline__ void gkr_block_reduce(Ef c0, Ef c2, Ef* red /* 2 * 32 */) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Ef o0, o2;
#pragma unroll
    for (int c = 0; c < 5; c++) {
      o0.c[c] = __shfl_down_sync(0xffffffffu, c0.c[c], off);
      o2.c[c] = __shfl_down_sync(0xffffffffu, c2.c[c], off);
    }
    c0 = ef_add(c0, o0), c2 = ef_add(c2, o2);
  }
  const int t = threadIdx.x;
  __syncthreads();  // red may still be read from a previous use
  if ((t & 31) == 0) red[2 * (t >> 5)] = c0, red[2 * (t >> 5) + 1] = c2;
  __syncthreads();
  if (t < 32) {
    const int nw = blockDim.x >> 5;
    c0 = t < nw ? red[2 * t] : ef_zero();
    c2 = t < nw ? red[2 * t + 1] : ef_zero();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Ef o0, o2;
#pragma unroll
      for (int c = 0; c < 5; c++) {
        o0.c[c] = __shfl_down_sync(0xffffffffu, c0.c[c], off);
        o2.c[c] = __shfl_down_sync(0xffffffffu, c2.c[c], off);
      }
      c0 = ef_add(c0, o0), c2 = ef_add(c2, o2);
    }
  }
  __syncthreads();
  if (t == 0) red[0] = c0, red[1] = c2;
  __syncthreads();
}

// ---- transcript steps (warp 0 of one CTA; prove_gkr_layer, quotient_gkr/mod.rs:80-141) ---------------------------
// after round `rnd`: build_bare_from_coeffs (sumcheck_utils.rs:491-503), add_sumcheck_polynomial, sample r
__device__ void gkr_step_round(GkrDev* g, DevFs* fs, uint32_t* tr, const Ef& c0r, const Ef& c2r, uint32_t rnd, uint32_t* sbuf,
                               const uint32_t* rc_s) {
  FsWarp w;
  w.load(fs, tr, rc_s);
  const uint32_t k = g->k;
  const Ef eq_alpha = g->point[k - 1 - rnd], mmf = g->mmf, s = g->s, one = fs_ef_one();
  const Ef c0 = fs_ef_mul(c0r, mmf), c2 = fs_ef_mul(c2r, mmf);
  const Ef ainv = g->inv_point[k - 1 - rnd];  // off the critical path: inverted when the layer began (gkr_invert_point)
  const Ef h1 = fs_ef_mul(ef_sub(s, fs_ef_mul(ef_sub(one, eq_alpha), c0)), ainv);
  const Ef b1 = ef_sub(ef_sub(h1, c0), c2);
  if ((threadIdx.x & 31) == 0) st_ef(sbuf, c0), st_ef(sbuf + 5, b1), st_ef(sbuf + 10, c2);
  __syncwarp();
  w.add_sumcheck_polynomial_bare(sbuf, 3, eq_alpha);
  const Ef r = w.sample_ef();
  const Ef eq_eval = fs_eq1(eq_alpha, r);
  const Ef val = ef_add(c0, fs_ef_mul(r, ef_add(b1, fs_ef_mul(r, c2))));
  const Ef s_new = fs_ef_mul(eq_eval, val), mmf_new = fs_ef_mul(mmf, eq_eval);
  if ((threadIdx.x & 31) == 0) g->s = s_new, g->mmf = mmf_new, g->r = r, g->q[rnd] = r;
  w.store();
}
// end of the layer: send the four inner evaluations, sample beta, next claim = line through them, next point = (q reversed, beta)
__device__ void gkr_step_end(GkrDev* g, DevFs* fs, uint32_t* tr, uint32_t* sbuf, const uint32_t* rc_s) {
  FsWarp w;
  w.load(fs, tr, rc_s);
  const uint32_t k = g->k;
  const Ef nl = g->inner[0], nr = g->inner[1], dl = g->inner[2], dr = g->inner[3];
  if ((threadIdx.x & 31) == 0) st_ef(sbuf, nl), st_ef(sbuf + 5, nr), st_ef(sbuf + 10, dl), st_ef(sbuf + 15, dr);
  __syncwarp();
  w.absorb(sbuf, 20);
  w.record(sbuf, 20);
  const Ef beta = w.sample_ef();
  const Ef cn = ef_add(nl, fs_ef_mul(beta, ef_sub(nr, nl))), cd = ef_add(dl, fs_ef_mul(beta, ef_sub(dr, dl)));
  // (1 - beta) nl + beta nr, written as the reference does not matter: the value is the same field element
  const int lane = threadIdx.x & 31;
  Ef qv[2];
  for (int t = 0; t < 2; t++) {
    const uint32_t i = lane + 32 * t;
    if (i < k) qv[t] = g->q[k - 1 - i];
  }
  __syncwarp();
  for (int t = 0; t < 2; t++) {
    const uint32_t i = lane + 32 * t;
    if (i < k) g->point[i] = qv[t];
  }
  if (lane == 0) g->point[k] = beta, g->claim_num = cn, g->claim_den = cd, g->k = k + 1;
  w.store();
}
// start of a layer: alpha after a duplex (mod.rs:88-91), running sum = claim_num + alpha claim_den
__device__ void gkr_step_begin(GkrDev* g, DevFs* fs, uint32_t* tr, const uint32_t* rc_s) {
  FsWarp w;
  w.load(fs, tr, rc_s);
  w.duplex();
  const Ef alpha = w.sample_ef();
  const Ef s = ef_add(g->claim_num, fs_ef_mul(alpha, g->claim_den));
  if ((threadIdx.x & 31) == 0) g->alpha = alpha, g->s = s, g->mmf = fs_ef_one();
  w.store();
}

// inverses of the k coordinates of the claim point, one thread each (a round's p(1) = (s - (1 - a) p(0)) / a needs 1 / a; on
// the transcript warp the inversion was 5 k of the ~45 k cycles of every round)
__device__ __forceinline__ void gkr_invert_point(GkrDev* g, DevFs* fs, uint32_t k) {
  if (threadIdx.x < k) {
    Ef inv;
    if (!fs_ef_inv(g->point[threadIdx.x], &inv) && fs) atomicOr(&fs->error, (uint32_t)DEVFS_ERR_ZERO_INV);
    g->inv_point[threadIdx.x] = inv;
  }
}

__global__ void __launch_bounds__(512) gkr_begin_kernel(GkrLayerArgs A, Ef scale, int sample_alpha) {
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  if (sample_alpha && threadIdx.x < 32) {
    fs_load_rc(rc_s);
    gkr_step_begin(A.g, A.fs, A.tr, rc_s);
  }
  if (!sample_alpha && threadIdx.x == 0) A.g->mmf = Ef{{KB_R1, 0, 0, 0, 0}};
  __syncthreads();
  if (sample_alpha) gkr_invert_point(A.g, A.fs, A.k);
  gkr_build_tables(A.eq_tab, A.g, A.k, scale);
  if (threadIdx.x == 0) A.g->k = A.k, A.g->counter = 0;
}

template <int MODE, int NUM_DIM>
__global__ void __launch_bounds__(256, 2) gkr_round_kernel(GkrLayerArgs A, uint32_t rnd) {
  __shared__ Ef red[64];
  __shared__ uint32_t sbuf[64];
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  __shared__ bool is_last;
  Ef c0 = ef_zero(), c2 = ef_zero();
  gkr_accumulate<MODE, NUM_DIM>(A, rnd, (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, (uint64_t)gridDim.x * blockDim.x, c0, c2);
  gkr_block_reduce(c0, c2, red);
  if (threadIdx.x == 0) {
    st_ef(A.partial + 10 * blockIdx.x, red[0]);
    st_ef(A.partial + 10 * blockIdx.x + 5, red[1]);
    __threadfence();
    is_last = atomicAdd(&A.g->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  c0 = ef_zero(), c2 = ef_zero();
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    Ef p0, p2;
#pragma unroll
    for (int c = 0; c < 5; c++) p0.c[c] = __ldcg(A.partial + 10 * i + c), p2.c[c] = __ldcg(A.partial + 10 * i + 5 + c);
    c0 = ef_add(c0, p0), c2 = ef_add(c2, p2);
  }
  gkr_block_reduce(c0, c2, red);
  if (threadIdx.x < 32) {
    if (A.fs) {
      fs_load_rc(rc_s);
      gkr_step_round(A.g, A.fs, A.tr, red[0], red[1], rnd, sbuf, rc_s);
    } else if (threadIdx.x == 0)
      A.g->out[0] = red[0], A.g->out[1] = red[1];
    if (threadIdx.x == 0) A.g->counter = 0;
  }
}

// last fold: the two remaining rows -> (nl, nr, dl, dr); one warp
template <int NUM_DIM>
__device__ void gkr_final_fold(const GkrLayerArgs& A) {
  const Ef r = A.g->r;
  Ef in[4];
  if (A.k == 1) {  // the only round read the layer itself: fractions 0..3
    const uint64_t n = 4;
    Ef lo[4], hi[4];
    for (int c = 0; c < 2; c++) {
      if (NUM_DIM == 1) {
        lo[c] = ef_from_base(A.nums[c]), hi[c] = ef_from_base(A.nums[2 + c]);
      } else {
        for (int k = 0; k < 5; k++) lo[c].c[k] = A.nums[k * n + c], hi[c].c[k] = A.nums[k * n + 2 + c];
      }
      for (int k = 0; k < 5; k++) lo[2 + c].c[k] = A.dens[k * n + c], hi[2 + c].c[k] = A.dens[k * n + 2 + c];
    }
    for (int c = 0; c < 4; c++) in[c] = fold5(r, lo[c], hi[c]);
  } else {
    const uint32_t* w = A.w[(A.k - 2) & 1];  // written by round k - 1: 2 rows
    for (int c = 0; c < 4; c++) {
      Ef lo, hi;
      for (int k = 0; k < 5; k++) {
        const uint2 v = ldcg2(w + (uint64_t)(5 * c + k) * 2);
        lo.c[k] = v.x, hi.c[k] = v.y;
      }
      in[c] = fold5(r, lo, hi);
    }
  }
  if ((threadIdx.x & 31) == 0)
    for (int c = 0; c < 4; c++) A.g->inner[c] = in[c];
  __syncwarp();
}

template <int NUM_DIM>
__global__ void gkr_end_kernel(GkrLayerArgs A) {
  __shared__ uint32_t sbuf[64];
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  gkr_final_fold<NUM_DIM>(A);
  if (A.fs) {
    fs_load_rc(rc_s);
    gkr_step_end(A.g, A.fs, A.tr, sbuf, rc_s);
  }
}

// One CTA: rounds rnd_start .. k-1 of the layer (<= 2^GKR_TAIL_VARS pairs each), the layer-end step and the begin step +
// eq tables of the next layer.  MODE0 = mode of the first round it runs.
template <int NUM_DIM>
__global__ void __launch_bounds__(512) gkr_tail_kernel(GkrLayerArgs A, uint32_t rnd_start, GkrLayerArgs next, int has_next) {
  __shared__ Ef red[64];
  __shared__ uint32_t sbuf[64];
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  if (threadIdx.x < 32) fs_load_rc(rc_s);
  for (uint32_t rnd = rnd_start; rnd < A.k; rnd++) {
#ifdef LM_GKR_PROFILE
    long long t_a = clock64();
#endif
    Ef c0 = ef_zero(), c2 = ef_zero();
    if (rnd == 0)
      gkr_accumulate<0, NUM_DIM>(A, rnd, threadIdx.x, blockDim.x, c0, c2);
    else if (rnd == 1)
      gkr_accumulate<1, NUM_DIM>(A, rnd, threadIdx.x, blockDim.x, c0, c2);
    else
      gkr_accumulate<2, NUM_DIM>(A, rnd, threadIdx.x, blockDim.x, c0, c2);
#ifdef LM_GKR_PROFILE
    long long t_b = clock64();
#endif
    gkr_block_reduce(c0, c2, red);
#ifdef LM_GKR_PROFILE
    long long t_c = clock64();
#endif
    if (threadIdx.x < 32) gkr_step_round(A.g, A.fs, A.tr, red[0], red[1], rnd, sbuf, rc_s);
    __threadfence_block();
    __syncthreads();
#ifdef LM_GKR_PROFILE
    if (threadIdx.x == 0 && A.k == 12) printf("k %u rnd %u: accumulate %lld, reduce %lld, step+sync %lld cycles\n", A.k, rnd, t_b - t_a, t_c - t_b, clock64() - t_c);
#endif
  }
  if (threadIdx.x < 32) {
    gkr_final_fold<NUM_DIM>(A);
    gkr_step_end(A.g, A.fs, A.tr, sbuf, rc_s);
    if (has_next) gkr_step_begin(A.g, A.fs, A.tr, rc_s);
  }
  __syncthreads();
  if (has_next) {
    gkr_invert_point(A.g, A.fs, next.k);
    gkr_build_tables(next.eq_tab, A.g, next.k, Ef{{KB_R1, 0, 0, 0, 0}});
    if (threadIdx.x == 0) A.g->counter = 0;
  }
}

cudaError_t gkr_begin(cudaStream_t stream, const GkrLayerArgs& a, const uint32_t eq_scale[5], bool sample_alpha) {
  Ef sc{{KB_R1, 0, 0, 0, 0}};
  if (eq_scale)
    for (int k = 0; k < 5; k++) sc.c[k] = eq_scale[k];
  gkr_begin_kernel<<<1, 512, 0, stream>>>(a, sc, sample_alpha ? 1 : 0);
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_round(cudaStream_t stream, const GkrLayerArgs& a, uint32_t rnd) {
  if (rnd >= a.k) return cudaErrorInvalidValue;
  const uint64_t half = (uint64_t)1 << (a.k - 1 - rnd);
  uint64_t blocks = (half + 255) / 256;
  if (blocks > (uint64_t)GKR_MAX_BLOCKS) blocks = GKR_MAX_BLOCKS;
  const unsigned g = (unsigned)blocks;
  const int mode = rnd == 0 ? 0 : (rnd == 1 ? 1 : 2);
  if (a.num_dim == 1) {
    if (mode == 0) gkr_round_kernel<0, 1><<<g, 256, 0, stream>>>(a, rnd);
    else if (mode == 1) gkr_round_kernel<1, 1><<<g, 256, 0, stream>>>(a, rnd);
    else gkr_round_kernel<2, 1><<<g, 256, 0, stream>>>(a, rnd);
  } else {
    if (mode == 0) gkr_round_kernel<0, 5><<<g, 256, 0, stream>>>(a, rnd);
    else if (mode == 1) gkr_round_kernel<1, 5><<<g, 256, 0, stream>>>(a, rnd);
    else gkr_round_kernel<2, 5><<<g, 256, 0, stream>>>(a, rnd);
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_end(cudaStream_t stream, const GkrLayerArgs& a) {
  if (a.num_dim == 1)
    gkr_end_kernel<1><<<1, 32, 0, stream>>>(a);
  else
    gkr_end_kernel<5><<<1, 32, 0, stream>>>(a);
  count_launch();
  return cudaGetLastError();
}

cudaError_t gkr_layer_device(cudaStream_t stream, const GkrLayerArgs& a, const GkrLayerArgs* next) {
  if (!a.fs) return cudaErrorInvalidValue;
  const uint32_t rnd_start = a.k > (uint32_t)GKR_TAIL_VARS + 1 ? a.k - (GKR_TAIL_VARS + 1) : 0;
  for (uint32_t rnd = 0; rnd < rnd_start; rnd++) {
    const cudaError_t e = gkr_round(stream, a, rnd);
    if (e != cudaSuccess) return e;
  }
  const GkrLayerArgs nx = next ? *next : a;
  if (a.num_dim == 1)
    gkr_tail_kernel<1><<<1, 512, 0, stream>>>(a, rnd_start, nx, next ? 1 : 0);
  else
    gkr_tail_kernel<5><<<1, 512, 0, stream>>>(a, rnd_start, nx, next ? 1 : 0);
  count_launch();
  return cudaGetLastError();
}

}  // namespace lm

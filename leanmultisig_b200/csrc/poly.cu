// Multilinear-extension kernels on sm_100a: eq tables, MLE evaluation, MSB-first folding.
//
// Device replacement for
//   crates/backend/poly/src/eq_mle.rs:16-26,85-150    eval_eq / compute_eval_eq (scaled eq table)
//   crates/backend/poly/src/evals.rs:142-347           eval_multilinear_generic (sqrt-split evaluation)
//   crates/backend/poly/src/utils.rs:161-186           fold_multilinear (MSB-first)
// MLE index convention: variable x_0 is the most-significant bit of the evaluation index (evals.rs:219).
//
// mle_eval: index = (hi | lo) with |lo| = LO_VARS = 10 bits.  Thread t of a CTA owns lo = t and walks a
// contiguous range of hi, so global reads are fully coalesced 4 KiB rows (base field) and the per-row factor
// eq_hi[hi] is a warp-uniform broadcast.  Each term is 5 IMAD.WIDE into 64-bit accumulators with delayed
// reduction; the kernel is HBM-bound (4 B per element read once).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include "launch_count.h"
#include "reduce.cuh"
#include "kb.cuh"
#include "poly.h"

namespace lm {

constexpr int LO_VARS = 10;
constexpr int EVAL_THREADS = 1 << LO_VARS;

// out[b] = scalar * prod_i (b_i ? z_i : 1 - z_i), b big-endian over k variables; one thread per entry.
__global__ void eq_table_kernel(const uint32_t* __restrict__ point, int k, Ef scalar, uint32_t* __restrict__ out) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ((uint64_t)1 << k)) return;
  Ef acc = scalar;
  for (int i = 0; i < k; i++) {
    Ef z;
#pragma unroll
    for (int c = 0; c < 5; c++) z.c[c] = __ldg(point + 5 * i + c);
    if (!((b >> (k - 1 - i)) & 1)) {
      // 1 - z
#pragma unroll
      for (int c = 0; c < 5; c++) z.c[c] = kb_neg(z.c[c]);
      z.c[0] = kb_add(z.c[0], KB_R1);
    }
    acc = ef_mul(acc, z);
  }
#pragma unroll
  for (int c = 0; c < 5; c++) out[5 * b + c] = acc.c[c];
}

cudaError_t eq_table(cudaStream_t stream, const uint32_t* d_point, int k, const uint32_t scalar[5], uint32_t* d_out) {
  Ef s;
  for (int c = 0; c < 5; c++) s.c[c] = scalar[c];
  const uint64_t n = (uint64_t)1 << k;
  eq_table_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_point, k, s, d_out); count_launch();
  return cudaGetLastError();
}

__device__ __forceinline__ Ef warp_reduce_ef(Ef v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Ef o;
#pragma unroll
    for (int c = 0; c < 5; c++) o.c[c] = __shfl_down_sync(0xffffffffu, v.c[c], off);
    v = ef_add(v, o);
  }
  return v;
}

// partial[blockIdx.x] = sum over this CTA's hi range of eq_hi[hi] * sum_lo eq_lo[lo] * f[hi, lo]
// n_lo = 2^lo_vars <= 1024 entries per row; rows >= live_rows are all-zero and skipped by the launcher.
template <int DIM>
__global__ void __launch_bounds__(EVAL_THREADS)
mle_eval_kernel(const uint32_t* __restrict__ evals, int lo_vars, uint64_t live_rows, uint64_t rows_per_cta,
                const uint32_t* __restrict__ eq_hi, const uint32_t* __restrict__ eq_lo, uint32_t* __restrict__ partial) {
  __shared__ Ef red[EVAL_THREADS / 32];
  const int t = threadIdx.x;
  const uint64_t n_lo = (uint64_t)1 << lo_vars;
  const uint64_t row0 = (uint64_t)blockIdx.x * rows_per_cta;
  uint64_t row1 = row0 + rows_per_cta;
  if (row1 > live_rows) row1 = live_rows;
  Ef acc = ef_zero();
  if ((uint64_t)t < n_lo) {
    uint64_t a[5] = {0, 0, 0, 0, 0};
    int terms = 0;
    for (uint64_t row = row0; row < row1; row++) {
      Ef e;
#pragma unroll
      for (int c = 0; c < 5; c++) e.c[c] = __ldg(eq_hi + 5 * row + c);
      if (DIM == 1) {
        const uint32_t f = __ldg(evals + row * n_lo + t);
        if (terms == 3) {
#pragma unroll
          for (int c = 0; c < 5; c++) a[c] = kb_fold(a[c]);
          terms = 0;
        }
#pragma unroll
        for (int c = 0; c < 5; c++) a[c] = mad_wide(f, e.c[c], a[c]);
        terms++;
      } else {
        Ef f;
#pragma unroll
        for (int c = 0; c < 5; c++) f.c[c] = __ldg(evals + (row * n_lo + t) * 5 + c);
        acc = ef_add(acc, ef_mul(f, e));
      }
    }
    if (DIM == 1) {
#pragma unroll
      for (int c = 0; c < 5; c++) acc.c[c] = kb_canon(kb_redc_lazy(kb_fold(a[c])));
    }
    Ef l;
#pragma unroll
    for (int c = 0; c < 5; c++) l.c[c] = __ldg(eq_lo + 5 * t + c);
    acc = ef_mul(acc, l);
  }
  acc = warp_reduce_ef(acc);
  if ((t & 31) == 0) red[t >> 5] = acc;
  __syncthreads();
  if (t < 32) {
    Ef v = (t < EVAL_THREADS / 32) ? red[t] : ef_zero();
    v = warp_reduce_ef(v);
    if (t == 0) {
#pragma unroll
      for (int c = 0; c < 5; c++) partial[5 * blockIdx.x + c] = v.c[c];
    }
  }
}

// Base-field polynomials, four consecutive low indices per thread: per row ONE 16-byte load and the five (CTA-uniform) words of
// eq_hi[row] feed twenty multiply-accumulates.  mle_eval_kernel<1> issues six loads per element (the eq_hi words once per
// thread and row) and is bound by the load / L1 pipe, not by HBM: 0.31 ms for 2^27 live entries = 1.7 TB/s (round 1: "0.23 of HBM").
constexpr int EVAL_VEC_THREADS = 256;
__global__ void __launch_bounds__(EVAL_VEC_THREADS)
mle_eval_vec4_kernel(const uint32_t* __restrict__ evals, int lo_vars, uint64_t live_rows, uint64_t rows_per_cta,
                     const uint32_t* __restrict__ eq_hi, const uint32_t* __restrict__ eq_lo, uint32_t* __restrict__ partial) {
  __shared__ Ef red[EVAL_VEC_THREADS / 32];
  const int t = threadIdx.x;
  const uint64_t n_lo = (uint64_t)1 << lo_vars;
  const uint64_t row0 = (uint64_t)blockIdx.x * rows_per_cta;
  uint64_t row1 = row0 + rows_per_cta;
  if (row1 > live_rows) row1 = live_rows;
  Ef acc = ef_zero();
  if ((uint64_t)(4 * t) < n_lo) {
    uint64_t a[4][5];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int c = 0; c < 5; c++) a[q][c] = 0;
    // rows in groups of three: the three 16-byte loads are in flight together, one fold per group (three products of
    // < 0.2462 * 2^64 on a folded accumulator)
    uint64_t row = row0;
    for (; row + 3 <= row1; row += 3) {
      uint4 f4[3];
      uint32_t e[3][5];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        f4[r] = __ldg(reinterpret_cast<const uint4*>(evals + (row + r) * n_lo) + t);
#pragma unroll
        for (int c = 0; c < 5; c++) e[r][c] = __ldg(eq_hi + 5 * (row + r) + c);
      }
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int c = 0; c < 5; c++) a[q][c] = kb_fold(a[q][c]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const uint32_t f[4] = {f4[r].x, f4[r].y, f4[r].z, f4[r].w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
          for (int c = 0; c < 5; c++) a[q][c] = mad_wide(f[q], e[r][c], a[q][c]);
      }
    }
    if (row < row1) {  // up to two rows left
#pragma unroll
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int c = 0; c < 5; c++) a[q][c] = kb_fold(a[q][c]);
      for (; row < row1; row++) {
        const uint4 f4 = __ldg(reinterpret_cast<const uint4*>(evals + row * n_lo) + t);
        const uint32_t f[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int c = 0; c < 5; c++) {
          const uint32_t ec = __ldg(eq_hi + 5 * row + c);
#pragma unroll
          for (int q = 0; q < 4; q++) a[q][c] = mad_wide(f[q], ec, a[q][c]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      Ef s, l;
#pragma unroll
      for (int c = 0; c < 5; c++) s.c[c] = kb_canon(kb_redc_lazy(kb_fold(a[q][c])));
#pragma unroll
      for (int c = 0; c < 5; c++) l.c[c] = __ldg(eq_lo + 5 * (4 * t + q) + c);
      acc = ef_add(acc, ef_mul(s, l));
    }
  }
  acc = warp_reduce_ef(acc);
  if ((t & 31) == 0) red[t >> 5] = acc;
  __syncthreads();
  if (t < 32) {
    Ef v = (t < EVAL_VEC_THREADS / 32) ? red[t] : ef_zero();
    v = warp_reduce_ef(v);
    if (t == 0) {
#pragma unroll
      for (int c = 0; c < 5; c++) partial[5 * blockIdx.x + c] = v.c[c];
    }
  }
}
// base-field evaluation kernel of choice for this shape
static void launch_mle_eval_base(cudaStream_t stream, const uint32_t* d_evals, int lo_vars, uint64_t live_rows, uint64_t rows_per_cta,
                                 uint64_t n_cta, const uint32_t* d_eq_hi, const uint32_t* d_eq_lo, uint32_t* d_partial) {
  static const bool vec_ok = getenv("LM_MLE_EVAL_SCALAR") == nullptr;
  if (vec_ok && lo_vars >= 2 && ((1u << lo_vars) / 4) <= (unsigned)EVAL_VEC_THREADS && (reinterpret_cast<uintptr_t>(d_evals) & 15) == 0)
    mle_eval_vec4_kernel<<<(unsigned)n_cta, EVAL_VEC_THREADS, 0, stream>>>(d_evals, lo_vars, live_rows, rows_per_cta, d_eq_hi, d_eq_lo, d_partial);
  else
    mle_eval_kernel<1><<<(unsigned)n_cta, EVAL_THREADS, 0, stream>>>(d_evals, lo_vars, live_rows, rows_per_cta, d_eq_hi, d_eq_lo, d_partial);
}

__global__ void sum_partials_kernel(const uint32_t* __restrict__ partial, int n, uint32_t* __restrict__ out) {
  // single CTA of 256 threads
  __shared__ Ef red[8];
  Ef acc = ef_zero();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    Ef v;
#pragma unroll
    for (int c = 0; c < 5; c++) v.c[c] = partial[5 * i + c];
    acc = ef_add(acc, v);
  }
  acc = warp_reduce_ef(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    Ef v = (threadIdx.x < 8) ? red[threadIdx.x] : ef_zero();
    v = warp_reduce_ef(v);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int c = 0; c < 5; c++) out[c] = v.c[c];
    }
  }
}

size_t mle_eval_scratch_words(uint32_t n_vars) {
  const int lo_vars = n_vars < (uint32_t)LO_VARS ? (int)n_vars : LO_VARS;
  const uint64_t n_hi = (uint64_t)1 << (n_vars - lo_vars);
  return (size_t)(5 * n_hi + 5 * ((uint64_t)1 << lo_vars) + 5 * 4096 + 8);
}

cudaError_t mle_eval(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim, uint64_t live_len,
                     const uint32_t* d_point, uint32_t* d_scratch, uint32_t* d_out) {
  if (dim != 1 && dim != 5) return cudaErrorInvalidValue;
  const int lo_vars = n_vars < (uint32_t)LO_VARS ? (int)n_vars : LO_VARS;
  const int hi_vars = (int)n_vars - lo_vars;
  const uint64_t n_hi = (uint64_t)1 << hi_vars, n_lo = (uint64_t)1 << lo_vars;
  uint32_t* d_eq_hi = d_scratch;
  uint32_t* d_eq_lo = d_eq_hi + 5 * n_hi;
  uint32_t* d_partial = d_eq_lo + 5 * n_lo;
  const uint32_t one[5] = {KB_R1, 0, 0, 0, 0};
  cudaError_t e;
  if ((e = eq_table(stream, d_point, hi_vars, one, d_eq_hi)) != cudaSuccess) return e;
  if ((e = eq_table(stream, d_point + 5 * hi_vars, lo_vars, one, d_eq_lo)) != cudaSuccess) return e;
  uint64_t live_rows = (live_len + n_lo - 1) / n_lo;
  if (live_rows > n_hi) live_rows = n_hi;
  // enough CTAs to fill the machine, at most 4096 partial sums
  uint64_t n_cta = live_rows < 4096 ? live_rows : 4096;
  if (n_cta == 0) n_cta = 1;
  const uint64_t rows_per_cta = (live_rows + n_cta - 1) / n_cta;
  n_cta = rows_per_cta ? (live_rows + rows_per_cta - 1) / rows_per_cta : 1;
  if (n_cta == 0) n_cta = 1;
  if (dim == 1)
    launch_mle_eval_base(stream, d_evals, lo_vars, live_rows, rows_per_cta, n_cta, d_eq_hi, d_eq_lo, d_partial);
  else
    mle_eval_kernel<5><<<(unsigned)n_cta, EVAL_THREADS, 0, stream>>>(d_evals, lo_vars, live_rows, rows_per_cta, d_eq_hi,
                                                                     d_eq_lo, d_partial); count_launch();
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  sum_partials_kernel<<<1, 256, 0, stream>>>(d_partial, (int)n_cta, d_out); count_launch();
  return cudaGetLastError();
}

// n_cols base-field columns at ONE point: the two eq tables are built once, every column is one streaming launch plus its
// partial-sum reduction, results land in d_out[5 k ..] — no host round trip between columns (Logup evaluates ~100 columns
// of a table at the same point, logup.rs:224-305).  d_cols / live_lens are host arrays.
cudaError_t mle_eval_batch(cudaStream_t stream, const uint32_t* const* d_cols, const uint64_t* live_lens, uint32_t n_cols,
                           uint32_t n_vars, const uint32_t* d_point, uint32_t* d_scratch, uint32_t* d_out) {
  const int lo_vars = n_vars < (uint32_t)LO_VARS ? (int)n_vars : LO_VARS;
  const int hi_vars = (int)n_vars - lo_vars;
  const uint64_t n_hi = (uint64_t)1 << hi_vars, n_lo = (uint64_t)1 << lo_vars;
  uint32_t* d_eq_hi = d_scratch;
  uint32_t* d_eq_lo = d_eq_hi + 5 * n_hi;
  uint32_t* d_partial = d_eq_lo + 5 * n_lo;
  const uint32_t one[5] = {KB_R1, 0, 0, 0, 0};
  cudaError_t e;
  if ((e = eq_table(stream, d_point, hi_vars, one, d_eq_hi)) != cudaSuccess) return e;
  if ((e = eq_table(stream, d_point + 5 * hi_vars, lo_vars, one, d_eq_lo)) != cudaSuccess) return e;
  for (uint32_t k = 0; k < n_cols; k++) {
    uint64_t live_rows = (live_lens[k] + n_lo - 1) / n_lo;
    if (live_rows > n_hi) live_rows = n_hi;
    uint64_t n_cta = live_rows < 4096 ? live_rows : 4096;
    if (n_cta == 0) n_cta = 1;
    const uint64_t rows_per_cta = (live_rows + n_cta - 1) / n_cta;
    n_cta = rows_per_cta ? (live_rows + rows_per_cta - 1) / rows_per_cta : 1;
    if (n_cta == 0) n_cta = 1;
    launch_mle_eval_base(stream, d_cols[k], lo_vars, live_rows, rows_per_cta, n_cta, d_eq_hi, d_eq_lo, d_partial);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    sum_partials_kernel<<<1, 256, 0, stream>>>(d_partial, (int)n_cta, d_out + 5 * k);
    count_launch();
  }
  return cudaGetLastError();
}

// out[i] = in[i] + r * (in[i + half] - in[i]),  i < half;   EF output
template <int DIM>
__global__ void fold_msb_kernel(const uint32_t* in, uint64_t half, uint64_t live, Ef r, uint32_t* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  Ef o;
  if (DIM == 1) {
    const uint32_t a = i < live ? __ldg(in + i) : 0u, b = i + half < live ? __ldg(in + i + half) : 0u;
    o = ef_add_base(ef_mul_base(r, kb_sub(b, a)), a);
  } else {
    Ef a = ef_zero(), b = ef_zero();
    if (i < live) {
#pragma unroll
      for (int c = 0; c < 5; c++) a.c[c] = in[5 * i + c];
    }
    if (i + half < live) {
#pragma unroll
      for (int c = 0; c < 5; c++) b.c[c] = in[5 * (i + half) + c];
    }
    o = ef_add(a, ef_mul(r, ef_sub(b, a)));
  }
#pragma unroll
  for (int c = 0; c < 5; c++) out[5 * i + c] = o.c[c];
}

cudaError_t fold_msb(cudaStream_t stream, const uint32_t* d_in, uint64_t n_in, uint32_t dim, uint64_t live,
                     const uint32_t r[5], uint32_t* d_out) {
  if ((dim != 1 && dim != 5) || n_in < 2) return cudaErrorInvalidValue;
  Ef rr;
  for (int c = 0; c < 5; c++) rr.c[c] = r[c];
  const uint64_t half = n_in / 2;
  const unsigned blocks = (unsigned)((half + 255) / 256);
  if (dim == 1)
    fold_msb_kernel<1><<<blocks, 256, 0, stream>>>(d_in, half, live, rr, d_out);
  else
    fold_msb_kernel<5><<<blocks, 256, 0, stream>>>(d_in, half, live, rr, d_out); count_launch();
  return cudaGetLastError();
}

// ---- access counts (memory_acc / bytecode_acc, crates/lean_prover/src/prove_execution.rs:91-110) --------------------
// counts[addr(i) + j] += 1 for every row i of an index column (Montgomery-form addresses) and j < n_values; the reference
// does this sequentially on the host ("TODO parallelize").  *d_bad is set when an address falls outside the table.
__global__ void access_count_kernel(const uint32_t* __restrict__ idx_col, uint64_t n, uint32_t n_values, uint64_t table_len,
                                    uint32_t* __restrict__ counts, uint32_t* __restrict__ d_bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t addr = kb_canon(kb_redc_lazy((uint64_t)__ldg(idx_col + i)));  // Montgomery -> canonical integer
  if (addr + n_values > table_len) {
    *d_bad = 1;
    return;
  }
  for (uint32_t j = 0; j < n_values; j++) atomicAdd(counts + addr + j, 1u);
}
__global__ void counts_to_monty_kernel(uint32_t* __restrict__ counts, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = kb_mul(counts[i] % KB_P, KB_R2);  // integer count -> field element in Montgomery form
}
cudaError_t access_count(cudaStream_t stream, const uint32_t* d_idx_col, uint64_t n, uint32_t n_values, uint64_t table_len,
                         uint32_t* d_counts, uint32_t* d_bad) {
  if (n == 0) return cudaSuccess;
  access_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_idx_col, n, n_values, table_len, d_counts, d_bad);
  count_launch();
  return cudaGetLastError();
}
cudaError_t counts_to_monty(cudaStream_t stream, uint32_t* d_counts, uint64_t n) {
  if (n == 0) return cudaSuccess;
  counts_to_monty_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_counts, n);
  count_launch();
  return cudaGetLastError();
}

// one CTA per opened leaf: the leaf in shared memory as extension elements, folded most-significant variable first
__global__ void __launch_bounds__(256) rows_mle_eval_kernel(const uint32_t* __restrict__ rows, int dim, int k, const uint32_t* __restrict__ point,
                                                            uint32_t* __restrict__ out) {
  extern __shared__ uint32_t sm_row[];  // 2^k x 5
  const uint32_t len = 1u << k;
  const uint32_t* row = rows + (uint64_t)blockIdx.x * len * dim;
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 5; c++) sm_row[5 * i + c] = dim == 5 ? row[5 * i + c] : (c == 0 ? row[i] : 0u);
  }
  __syncthreads();
  for (int v = 0; v < k; v++) {
    const uint32_t half = len >> (v + 1);
    const Ef x = ld_ef(point + 5 * v);
    Ef res[1];
    // every thread handles at most one pair per level for leaves up to 2^9 elements; larger leaves loop
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
      const Ef lo = ld_ef_rw(sm_row + 5 * i), hi = ld_ef_rw(sm_row + 5 * (i + half));
      res[0] = ef_add(lo, ef_mul(x, ef_sub(hi, lo)));
      st_ef(sm_row + 5 * i, res[0]);  // in place: slot i is only read by its own thread at this level
    }
    __syncthreads();
  }
  if (threadIdx.x < 5) out[5 * blockIdx.x + threadIdx.x] = sm_row[threadIdx.x];
}

cudaError_t rows_mle_eval(cudaStream_t stream, const uint32_t* d_rows, uint32_t n_rows, uint32_t dim, uint32_t k, const uint32_t* d_point,
                          uint32_t* d_out) {
  if (n_rows == 0) return cudaSuccess;
  if ((dim != 1 && dim != 5) || k > 12) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)5 << k) * sizeof(uint32_t);
  if (smem > 48 * 1024) cudaFuncSetAttribute(rows_mle_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rows_mle_eval_kernel<<<n_rows, 256, smem, stream>>>(d_rows, (int)dim, (int)k, d_point, d_out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace lm

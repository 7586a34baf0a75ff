// AIR ("SuperSpartan") sumcheck rounds for the lean_vm execution table on sm_100a.
//
// Device replacement for
//   crates/sub_protocols/src/air_sumcheck.rs:225-287   compute_bare_round_poly / process_challenge (the heavy bodies)
//   crates/sub_protocols/src/air_sumcheck.rs:560-634   compute_raw_poly_impl: evaluations at z = 0, 2, .., d
//   crates/sub_protocols/src/air_sumcheck.rs:683-694   compute_shifted_columns
//   crates/backend/poly/src/utils.rs:117-160           fold_multilinear_at_bit
//   crates/lean_vm/src/tables/execution/air.rs:56-130  ExecutionTable::eval (13 constraints, degree 5, 20 + 2 columns)
//   crates/lean_vm/src/tables/utils.rs:5-21            eval_virtual_bus_column
//   crates/backend/air/src/constraint_folder/normal.rs:49-62  accumulator += alpha^k * constraint_k
//
// Layout.  The reference bit-reverses every column inside 2^12-row chunks so that AVX lanes stay full while it
// folds "right to left"; on the GPU the natural row order already gives the best access pattern for that fold
// order: round r pairs rows (2j, 2j+1), thread j reads 8 B (base) / 40 B (EF) contiguous per column and writes
// folded row j.  Columns are SoA: base rounds u32[c][n], EF rounds u32[c][n][5].  The eq factor of the free
// variables is split as eq_hi[j >> 10] * eq_lo[j & 1023] (the reference's SplitEq, split_eq.rs:5-103).
// Round evaluations are reduced warp -> CTA -> partials -> one final CTA; the host does p(1), the Lagrange
// interpolation and the transcript (air_sumcheck.rs:250-266).
#include <cuda_runtime.h>
#include <cstdint>
#include "air.h"
#include "air_values.cuh"
#include "kb.cuh"
#include "launch_count.h"
#include "poly.h"

namespace lm {

constexpr int EXEC_COLS = 20, EXEC_SHIFT = 2, EXEC_ALL = 22, EXEC_DEG = 5;

// sum_k alpha^k * constraint_k(point), point = 20 flat + 2 shift values (execution/air.rs:56-130)
template <class T>
__device__ __forceinline__ Ef exec_air_eval(const T* pt, const AirExtra& X) {
  const T pc = pt[0], fp = pt[1], addr_a = pt[2], addr_b = pt[3], addr_c = pt[4];
  const T val_a = pt[5], val_b = pt[6], val_c = pt[7], op_a = pt[8], op_b = pt[9], op_c = pt[10];
  const T flag_a = pt[11], flag_b = pt[12], flag_c = pt[13], flag_c_fp = pt[14], flag_ab_fp = pt[15];
  const T mul = pt[16], jump = pt[17], aux = pt[18], pdata = pt[19];
  const T pc_shift = pt[20], fp_shift = pt[21];

  const T om_a = -sub_one(flag_a + flag_ab_fp);
  const T om_b = -sub_one(flag_b + flag_ab_fp);
  const T om_c = -sub_one(flag_c + flag_c_fp);
  const T fp_op_a = fp + op_a, fp_op_b = fp + op_b, fp_op_c = fp + op_c;
  const T nu_a = flag_a * op_a + om_a * val_a + flag_ab_fp * fp_op_a;
  const T nu_b = flag_b * op_b + om_b * val_b + flag_ab_fp * fp_op_b;
  const T nu_c = flag_c * op_c + om_c * val_c + flag_c_fp * fp_op_c;
  const T add = dbl(aux) - aux * aux;
  const T deref = halve(aux * sub_one(aux));
  const T is_precompile = -sub_one(add + mul + deref + jump);

  // bus column: (sum_i la[i] data[i] + la_last * DOMAINSEP(=1)) * beta + flag
  Ef bus = scale(X.la[0], pdata) + scale(X.la[1], nu_a) + scale(X.la[2], nu_b) + scale(X.la[3], nu_c) + X.la_last;
  bus = add_val(ef_mul(bus, X.beta), is_precompile);
  Ef acc = ef_mul(X.alpha[0], bus);
  acc = acc + scale(X.alpha[1], om_a * (addr_a - fp_op_a));
  acc = acc + scale(X.alpha[2], om_b * (addr_b - fp_op_b));
  acc = acc + scale(X.alpha[3], om_c * (addr_c - fp_op_c));
  acc = acc + scale(X.alpha[4], add * (nu_b - (nu_a + nu_c)));
  acc = acc + scale(X.alpha[5], mul * (nu_b - nu_a * nu_c));
  acc = acc + scale(X.alpha[6], deref * (addr_b - (val_a + op_b)));
  acc = acc + scale(X.alpha[7], deref * (val_b - nu_c));
  const T jc = jump * nu_a;
  acc = acc + scale(X.alpha[8], jc * sub_one(nu_a));
  acc = acc + scale(X.alpha[9], jc * (pc_shift - nu_b));
  acc = acc + scale(X.alpha[10], jc * (fp_shift - nu_c));
  const T njc = -sub_one(jc);
  acc = acc + scale(X.alpha[11], njc * (pc_shift - add_one(pc)));
  acc = acc + scale(X.alpha[12], njc * (fp_shift - fp));
  return acc;
}

template <class T>
__device__ __forceinline__ T ld_val(const uint32_t* col, uint64_t row);
template <>
__device__ __forceinline__ Fb ld_val<Fb>(const uint32_t* col, uint64_t row) {
  return Fb{__ldg(col + row)};
}
template <>
__device__ __forceinline__ Ef ld_val<Ef>(const uint32_t* col, uint64_t row) {
  Ef v;
#pragma unroll
  for (int c = 0; c < 5; c++) v.c[c] = __ldg(col + 5 * row + c);
  return v;
}

// partial[blockIdx.x][z] = sum over this CTA's j of eq(j) * C(col(2j) + z (col(2j+1) - col(2j))), z = 0,2,3,4,5
// EF rounds keep the 22 per-column differences in shared memory (one 4-byte-interleaved slot per thread, so the
// accesses are conflict free): holding point and difference in registers at once is 220 words and spilled 1.4 KiB
// per thread in the first version of this kernel.
constexpr int AIR_THREADS = 128;
template <class T, int DIM>
__global__ void __launch_bounds__(AIR_THREADS)
air_exec_round_kernel(const uint32_t* __restrict__ cols, uint64_t n, uint64_t half, const uint32_t* __restrict__ eq_hi,
                      const uint32_t* __restrict__ eq_lo, int lo_vars, AirExtra X, uint32_t* __restrict__ partial) {
  __shared__ Ef red[EXEC_DEG][AIR_THREADS / 32];
  extern __shared__ uint32_t diff_s[];  // DIM == 5: [22 * 5][AIR_THREADS]
  Ef acc[EXEC_DEG];
#pragma unroll
  for (int z = 0; z < EXEC_DEG; z++) acc[z] = ef_zero();
  const int tid = threadIdx.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + tid; j < half; j += (uint64_t)gridDim.x * blockDim.x) {
    T pt[EXEC_ALL];
    T diff_r[DIM == 1 ? EXEC_ALL : 1];
#pragma unroll
    for (int c = 0; c < EXEC_ALL; c++) {
      const uint32_t* col = cols + (uint64_t)c * n * DIM;
      const T lo = ld_val<T>(col, 2 * j), hi = ld_val<T>(col, 2 * j + 1);
      pt[c] = lo;
      const T d = hi - lo;
      if constexpr (DIM == 1) {
        diff_r[c] = d;
      } else {
#pragma unroll
        for (int k = 0; k < 5; k++) diff_s[(c * 5 + k) * AIR_THREADS + tid] = d.c[k];
      }
    }
    Ef eh, el;
#pragma unroll
    for (int c = 0; c < 5; c++) {
      eh.c[c] = __ldg(eq_hi + 5 * (j >> lo_vars) + c);
      el.c[c] = __ldg(eq_lo + 5 * (j & (((uint64_t)1 << lo_vars) - 1)) + c);
    }
    const Ef eq = ef_mul(eh, el);
#pragma unroll 1
    for (int zi = 0; zi < EXEC_DEG; zi++) {
      if (zi >= 1) {
        const int steps = zi == 1 ? 2 : 1;  // z: 0 -> 2 -> 3 -> 4 -> 5
        for (int s2 = 0; s2 < steps; s2++) {
#pragma unroll
          for (int c = 0; c < EXEC_ALL; c++) {
            if constexpr (DIM == 1) {
              pt[c] = pt[c] + diff_r[c];
            } else {
              Ef d;
#pragma unroll
              for (int k = 0; k < 5; k++) d.c[k] = diff_s[(c * 5 + k) * AIR_THREADS + tid];
              pt[c] = pt[c] + d;
            }
          }
        }
      }
      const Ef v = ef_mul(exec_air_eval<T>(pt, X), eq);
      // acc[zi] += v with a register-indexed accumulator (zi is a runtime loop counter)
#pragma unroll
      for (int z = 0; z < EXEC_DEG; z++)
        if (z == zi) acc[z] = ef_add(acc[z], v);
    }
  }
  // warp then CTA reduction
#pragma unroll
  for (int z = 0; z < EXEC_DEG; z++) {
    Ef v = acc[z];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Ef o;
#pragma unroll
      for (int c = 0; c < 5; c++) o.c[c] = __shfl_down_sync(0xffffffffu, v.c[c], off);
      v = ef_add(v, o);
    }
    if ((threadIdx.x & 31) == 0) red[z][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < EXEC_DEG) {
    Ef v = red[threadIdx.x][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = ef_add(v, red[threadIdx.x][w]);
#pragma unroll
    for (int c = 0; c < 5; c++) partial[(uint64_t)blockIdx.x * EXEC_DEG * 5 + threadIdx.x * 5 + c] = v.c[c];
  }
}

__global__ void air_sum_partials_kernel(const uint32_t* __restrict__ partial, int n_part, int n_vals, uint32_t* __restrict__ out) {
  // one thread per output EF value (n_vals <= 32): sums are tiny (<= 1184 partials)
  const int z = threadIdx.x;
  if (z >= n_vals) return;
  Ef v = ef_zero();
  for (int k = 0; k < n_part; k++) {
    Ef o;
#pragma unroll
    for (int c = 0; c < 5; c++) o.c[c] = partial[((uint64_t)k * n_vals + z) * 5 + c];
    v = ef_add(v, o);
  }
#pragma unroll
  for (int c = 0; c < 5; c++) out[5 * z + c] = v.c[c];
}

// fold the least-significant variable of every column: out[c][j] = in[c][2j] + r (in[c][2j+1] - in[c][2j])
template <int DIM>
__global__ void air_fold_lsb_kernel(const uint32_t* in, uint64_t n, int n_cols, Ef r, uint32_t* out) {
  const uint64_t half = n / 2;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= half * n_cols) return;
  const uint64_t c = idx / half, j = idx % half;
  Ef o;
  if (DIM == 1) {
    const uint2 ab = *reinterpret_cast<const uint2*>(in + c * n + 2 * j);
    o = ef_add_base(ef_mul_base(r, kb_sub(ab.y, ab.x)), ab.x);
  } else {
    Ef a, b;
#pragma unroll
    for (int k = 0; k < 5; k++) a.c[k] = in[(c * n + 2 * j) * 5 + k], b.c[k] = in[(c * n + 2 * j + 1) * 5 + k];
    o = ef_add(a, ef_mul(r, ef_sub(b, a)));
  }
#pragma unroll
  for (int k = 0; k < 5; k++) out[(c * half + j) * 5 + k] = o.c[k];
}

// shifted[i] = col[i + 1], last row repeated (air_sumcheck.rs:683-694)
__global__ void air_shift_kernel(const uint32_t* __restrict__ col, uint64_t n, uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __ldg(col + (i + 1 < n ? i + 1 : n - 1));
}

cudaError_t air_shift_column(cudaStream_t stream, const uint32_t* d_col, uint64_t n, uint32_t* d_out) {
  air_shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_col, n, d_out);
  count_launch();
  return cudaGetLastError();
}

size_t air_round_scratch_words(uint32_t log_n) {
  const uint32_t lv = log_n ? log_n - 1 : 0;
  const int lo = lv < (uint32_t)AIR_LO ? (int)lv : AIR_LO;
  // partial sums: 148 * 8 CTAs x 5 values (execution table) or 148 * 4 CTAs x 10 values (air_generic.cu)
  return 5 * (((size_t)1 << (lv - lo)) + ((size_t)1 << lo)) + (size_t)(148 * 8) * 10 * 5 + 64;
}

cudaError_t air_exec_round(cudaStream_t stream, const uint32_t* d_cols, uint32_t dim, uint32_t log_n, const uint32_t* d_eq_point,
                           const uint32_t* alpha_powers, const uint32_t* la, uint32_t n_la, const uint32_t beta[5],
                           uint32_t* d_scratch, uint32_t* d_out, const uint32_t* eq_scale) {
  if (log_n < 1 || (dim != 1 && dim != 5) || n_la < 5 || n_la > 64) return cudaErrorInvalidValue;
  AirExtra X;
  for (int k = 0; k < 13; k++)
    for (int c = 0; c < 5; c++) X.alpha[k].c[c] = alpha_powers[5 * k + c];
  for (int k = 13; k < 16; k++) X.alpha[k] = Ef{{0, 0, 0, 0, 0}};
  for (int k = 0; k < 4; k++)
    for (int c = 0; c < 5; c++) X.la[k].c[c] = la[5 * k + c];
  for (int k = 4; k < 8; k++) X.la[k] = Ef{{0, 0, 0, 0, 0}};
  // la_last * DOMAINSEP with LOGUP_PRECOMPILE_DOMAINSEP = 1 (lean_vm/src/core/constants.rs:5)
  for (int c = 0; c < 5; c++) X.la_last.c[c] = la[5 * (n_la - 1) + c], X.beta.c[c] = beta[c];

  const uint64_t n = (uint64_t)1 << log_n, half = n / 2;
  const uint32_t lv = log_n - 1;
  const int lo_vars = lv < (uint32_t)AIR_LO ? (int)lv : AIR_LO;
  const int hi_vars = (int)lv - lo_vars;
  uint32_t* d_hi = d_scratch;
  uint32_t* d_lo = d_hi + 5 * ((size_t)1 << hi_vars);
  uint32_t* d_part = d_lo + 5 * ((size_t)1 << lo_vars);
  const uint32_t one[5] = {KB_R1, 0, 0, 0, 0};
  cudaError_t e;
  if ((e = eq_table(stream, d_eq_point, hi_vars, eq_scale ? eq_scale : one, d_hi)) != cudaSuccess) return e;
  if ((e = eq_table(stream, d_eq_point + 5 * hi_vars, lo_vars, one, d_lo)) != cudaSuccess) return e;
  uint64_t blocks = (half + AIR_THREADS - 1) / AIR_THREADS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dim == 1) {
    air_exec_round_kernel<Fb, 1><<<(unsigned)blocks, AIR_THREADS, 0, stream>>>(d_cols, n, half, d_hi, d_lo, lo_vars, X, d_part);
  } else {
    const size_t smem = (size_t)EXEC_ALL * 5 * AIR_THREADS * sizeof(uint32_t);  // 55 KiB
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(air_exec_round_kernel<Ef, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_set = true;
    }
    air_exec_round_kernel<Ef, 5><<<(unsigned)blocks, AIR_THREADS, smem, stream>>>(d_cols, n, half, d_hi, d_lo, lo_vars, X, d_part);
  }
  count_launch();
  air_sum_partials_kernel<<<1, 32, 0, stream>>>(d_part, (int)blocks, EXEC_DEG, d_out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t air_fold_lsb(cudaStream_t stream, const uint32_t* d_in, uint32_t dim, uint64_t n, uint32_t n_cols, const uint32_t r[5],
                         uint32_t* d_out) {
  if (n < 2 || (dim != 1 && dim != 5)) return cudaErrorInvalidValue;
  Ef rr;
  for (int c = 0; c < 5; c++) rr.c[c] = r[c];
  const uint64_t total = (n / 2) * n_cols;
  if (dim == 1)
    air_fold_lsb_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_in, n, (int)n_cols, rr, d_out);
  else
    air_fold_lsb_kernel<5><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_in, n, (int)n_cols, rr, d_out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace lm

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
// order: round r pairs rows (2j, 2j+1).  Base columns are u32[c][n], folded columns coefficient planes u32[c][5][n]
// (air.h), so the 8 / 16-byte loads of a warp are contiguous.
//
// One pass over the table per round: the fold of a challenge is applied by the round that next reads the table
// (air.h, AirExecMode).  A CTA stages 32 row pairs (row 2j and the difference to row 2j+1) in shared memory and its five
// warps evaluate them at z = 0, 2, 3, 4, 5, one point per warp; the constraint code reads column values from shared
// memory, so the register file holds only the evaluator's temporaries.  The weighted sum
// sum_k alpha^k C_k (+ the bus column) is ONE delayed-reduction accumulation per coefficient against weights the host
// pre-multiplied (alpha^0 beta la_i for the bus data), not 13 reduced products.  The eq factor of the free variables is
// eq_hi[j >> s] * eq_lo[j & mask] from the session's prefix tables (eqtab.cuh; the reference's SplitEq,
// split_eq.rs:5-103).  Round sums are reduced warp -> CTA -> last CTA; the host does p(1), the Lagrange interpolation
// and the transcript (air_sumcheck.rs:250-266).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include "air.h"
#include "air_values.cuh"
#include "eqtab.cuh"
#include "kb.cuh"
#include "launch_count.h"
#include "poly.h"

namespace lm {

constexpr int EXEC_COLS = 20, EXEC_ALL = 22, EXEC_DEG = 5;

// weights of the 17 terms of the accumulator: 0..3 = alpha^0 beta la_i for the bus data (precompile data, nu_a, nu_b, nu_c),
// 4 = alpha^0 (bus flag), 5..16 = alpha^1..alpha^12; k0 = alpha^0 beta la_last * DOMAINSEP
struct AirExecConsts {
  EfRows w[17];
  Ef k0;
};

// sum_t weight_t * value_t, five 64-bit accumulators, a fold every four products (kb.cuh KbDot); `pending` is a
// compile-time constant after inlining (the call sequence is straight-line code)
struct AlphaAcc {
  uint64_t a[5];
  int pending;
  __device__ __forceinline__ AlphaAcc() : a{0, 0, 0, 0, 0}, pending(0) {}
  __device__ __forceinline__ void step() {
    if (pending == 4) {
#pragma unroll
      for (int i = 0; i < 5; i++) a[i] = kb_fold(a[i]);
      pending = 0;
    }
    pending++;
  }
  __device__ __forceinline__ void mac(const EfRows& w, Fb v) {
    step();
    a[0] = mad_wide(v.v, w.b0, a[0]);
    a[1] = mad_wide(v.v, w.b1, a[1]);
    a[2] = mad_wide(v.v, w.b2, a[2]);
    a[3] = mad_wide(v.v, w.b3, a[3]);
    a[4] = mad_wide(v.v, w.b4, a[4]);
  }
  __device__ __forceinline__ void mac(const EfRows& w, const Ef& v) {
    const uint32_t rows[5][5] = {{w.b0, w.b4, w.b3, w.b2, w.b1m4},
                                 {w.b1, w.b0, w.b4, w.b3, w.b2},
                                 {w.b2, w.b1m4, w.b0m3, w.b4m2, w.b3m14},
                                 {w.b3, w.b2, w.b1m4, w.b0m3, w.b4m2},
                                 {w.b4, w.b3, w.b2, w.b1m4, w.b0m3}};
#pragma unroll
    for (int k = 0; k < 5; k++) {
      step();
#pragma unroll
      for (int i = 0; i < 5; i++) a[i] = mad_wide(v.c[k], rows[i][k], a[i]);
    }
  }
  __device__ __forceinline__ Ef finish(const Ef& k0) const {
    Ef r;
#pragma unroll
    for (int i = 0; i < 5; i++) r.c[i] = kb_add(kb_canon(kb_redc_lazy(kb_fold(a[i]))), k0.c[i]);
    return r;
  }
};

// Extension-field products of the constraint code are calls of two out-of-line functions: the evaluator shrinks from
// ~1.6 k to ~0.9 k IMAD.WIDE (the kernels stall on instruction fetch, ncu no_instruction 1.6 warps per issue), measured
// 8 % faster on the 2^22-row table than fully inlined (-DLM_AIR_INLINE_MUL).
#ifndef LM_AIR_INLINE_MUL
static __device__ __noinline__ Ef air_ef_mul(Ef a, Ef b) { return ef_mul(a, b); }
static __device__ __noinline__ Ef air_ef_mul2_add(Ef a, Ef b, Ef c, Ef d) { return ef_mul2_add(a, b, c, d); }
#else
__device__ __forceinline__ Ef air_ef_mul(const Ef& a, const Ef& b) { return ef_mul(a, b); }
__device__ __forceinline__ Ef air_ef_mul2_add(const Ef& a, const Ef& b, const Ef& c, const Ef& d) { return ef_mul2_add(a, b, c, d); }
#endif
__device__ __forceinline__ Fb mul1(Fb a, Fb b) { return a * b; }
__device__ __forceinline__ Ef mul1(const Ef& a, const Ef& b) { return air_ef_mul(a, b); }
__device__ __forceinline__ Fb mul2_add(Fb a, Fb b, Fb c, Fb d) { return a * b + c * d; }
__device__ __forceinline__ Ef mul2_add(const Ef& a, const Ef& b, const Ef& c, const Ef& d) { return air_ef_mul2_add(a, b, c, d); }

// sum_k alpha^k * constraint_k(point), point = 20 flat + 2 shift values (execution/air.rs:56-130).  The value-selection
// terms are written with one product less than the reference's text, nu = val + flag (op - val) + flag_fp (fp + op - val),
// which is the same polynomial as flag op + (1 - flag - flag_fp) val + flag_fp (fp + op).
template <class T, class View>
__device__ __forceinline__ Ef exec_air_eval(const View& col, const AirExecConsts& X) {
  AlphaAcc acc;
  const T fp = col(1);
  const T op_a = col(8), op_b = col(9), op_c = col(10);
  const T val_a = col(5), val_b = col(6), val_c = col(7);
  const T flag_a = col(11), flag_b = col(12), flag_c = col(13), flag_c_fp = col(14), flag_ab_fp = col(15);
  const T fp_op_a = fp + op_a, fp_op_b = fp + op_b, fp_op_c = fp + op_c;
  const T nu_a = val_a + mul2_add(flag_a, op_a - val_a, flag_ab_fp, fp_op_a - val_a);
  const T nu_b = val_b + mul2_add(flag_b, op_b - val_b, flag_ab_fp, fp_op_b - val_b);
  const T nu_c = val_c + mul2_add(flag_c, op_c - val_c, flag_c_fp, fp_op_c - val_c);
  const T aux = col(18), mul = col(16), jump = col(17);
  const T aux2 = mul1(aux, aux);
  const T add = dbl(aux) - aux2;
  const T deref = halve(aux2 - aux);
  const T is_precompile = -sub_one(add + mul + deref + jump);
  // bus column: (sum_i la[i] data[i] + la_last * DOMAINSEP(=1)) * beta + flag, times alpha^0
  acc.mac(X.w[0], col(19));
  acc.mac(X.w[1], nu_a);
  acc.mac(X.w[2], nu_b);
  acc.mac(X.w[3], nu_c);
  acc.mac(X.w[4], is_precompile);
  const T om_a = -sub_one(flag_a + flag_ab_fp);
  const T om_b = -sub_one(flag_b + flag_ab_fp);
  const T om_c = -sub_one(flag_c + flag_c_fp);
  const T addr_b = col(3);
  acc.mac(X.w[5], mul1(om_a, col(2) - fp_op_a));
  acc.mac(X.w[6], mul1(om_b, addr_b - fp_op_b));
  acc.mac(X.w[7], mul1(om_c, col(4) - fp_op_c));
  acc.mac(X.w[8], mul1(add, nu_b - (nu_a + nu_c)));
  acc.mac(X.w[9], mul1(mul, nu_b - mul1(nu_a, nu_c)));
  acc.mac(X.w[10], mul1(deref, addr_b - (val_a + op_b)));
  acc.mac(X.w[11], mul1(deref, val_b - nu_c));
  const T jc = mul1(jump, nu_a);
  acc.mac(X.w[12], mul1(jc, sub_one(nu_a)));
  const T pc_shift = col(20), fp_shift = col(21);
  acc.mac(X.w[13], mul1(jc, pc_shift - nu_b));
  acc.mac(X.w[14], mul1(jc, fp_shift - nu_c));
  const T njc = -sub_one(jc);
  acc.mac(X.w[15], mul1(njc, pc_shift - add_one(col(0))));
  acc.mac(X.w[16], mul1(njc, fp_shift - fp));
  return acc.finish(X.k0);
}

// A CTA is five warps, one per evaluation point z = 0, 2, 3, 4, 5, and works on 32 row pairs at a time: all 160 threads
// stage the pairs in shared memory - for every word of every column the five values row(2j) + z (row(2j+1) - row(2j)),
// five additions, word-major so that lane = pair is conflict free - then warp z evaluates the constraints of its 32 pairs
// reading plain words.
constexpr int EXEC_WARPS = EXEC_DEG, EXEC_THREADS = 32 * EXEC_WARPS, EXEC_PAIRS = 32;

// column values of pair `lane` at this warp's z: pt points at [z][0][lane] of the staged points
template <class T>
struct SmView;
template <>
struct SmView<Fb> {
  const uint32_t* pt;
  __device__ __forceinline__ Fb operator()(int c) const { return Fb{pt[c * EXEC_PAIRS]}; }
};
template <>
struct SmView<Ef> {
  const uint32_t* pt;
  __device__ __forceinline__ Ef operator()(int c) const {
    Ef r;
#pragma unroll
    for (int k = 0; k < 5; k++) r.c[k] = pt[(c * 5 + k) * EXEC_PAIRS];
    return r;
  }
};

__device__ __forceinline__ uint2 a_ldg2(const uint32_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ uint4 a_ldg4(const uint32_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint2 a_ldcg2(const uint32_t* p) { return __ldcg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ uint4 a_ldcg4(const uint32_t* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ Ef a_fold1(const Ef& r, uint32_t lo, uint32_t hi) { return ef_add_base(ef_mul_base(r, kb_sub(hi, lo)), lo); }
__device__ __forceinline__ Ef a_fold5(const Ef& r, const Ef& lo, const Ef& hi) { return ef_add(lo, ef_mul(r, ef_sub(hi, lo))); }

// the five points a + z (b - a), z = 0, 2, 3, 4, 5 of one word; pt points at the pair's slot, zs = words per point
__device__ __forceinline__ void stage_word(uint32_t* pt, int zs, int w, uint32_t a, uint32_t b) {
  const uint32_t d = kb_sub(b, a);
  uint32_t v = kb_add(b, d);  // z = 2
  pt[w * EXEC_PAIRS] = a;
  pt[zs + w * EXEC_PAIRS] = v;
#pragma unroll
  for (int z = 2; z < EXEC_DEG; z++) {
    v = kb_add(v, d);
    pt[z * zs + w * EXEC_PAIRS] = v;
  }
}
__device__ __forceinline__ void stage_put(uint32_t* pt, int zs, int c, Fb a, Fb b) { stage_word(pt, zs, c, a.v, b.v); }
__device__ __forceinline__ void stage_put(uint32_t* pt, int zs, int c, const Ef& a, const Ef& b) {
#pragma unroll
  for (int k = 0; k < 5; k++) stage_word(pt, zs, c * 5 + k, a.c[k], b.c[k]);
}
__device__ __forceinline__ void plane_put(uint32_t* dst, uint64_t rows, int c, uint64_t j, const Ef& a, const Ef& b) {
#pragma unroll
  for (int k = 0; k < 5; k++)
    *reinterpret_cast<uint2*>(dst + (uint64_t)(c * 5 + k) * rows + 2 * j) = make_uint2(a.c[k], b.c[k]);
}

// rows (2j, 2j+1) of the current table, column c (c < 20 in the base modes, which also produce the shifted column
// behind columns 0 and 1; c < 22 in the extension modes) -> shared memory (and, in the modes that materialise, -> dst)
template <int MODE>
__device__ __forceinline__ void exec_stage(const AirExecArgs& A, int c, uint64_t j, const Ef& r_new, const Ef& r_old, uint32_t* pt,
                                           int df) {
  const uint64_t n = A.n_base;
  const uint64_t rows_dst = (uint64_t)2 << A.m;
  if (MODE == AIR_B0) {
    const uint2 v = a_ldg2(A.base + (uint64_t)c * n + 2 * j);
    stage_put(pt, df, c, Fb{v.x}, Fb{v.y});
    if (c < 2) {
      const uint32_t nx = 2 * j + 2 < n ? __ldg(A.base + (uint64_t)c * n + 2 * j + 2) : A.halo[c];
      stage_put(pt, df, EXEC_COLS + c, Fb{v.y}, Fb{nx});
    }
  } else if (MODE == AIR_B1) {
    const uint4 v = a_ldg4(A.base + (uint64_t)c * n + 4 * j);
    stage_put(pt, df, c, a_fold1(r_new, v.x, v.y), a_fold1(r_new, v.z, v.w));
    if (c < 2) {
      const uint32_t nx = 4 * j + 4 < n ? __ldg(A.base + (uint64_t)c * n + 4 * j + 4) : A.halo[c];
      stage_put(pt, df, EXEC_COLS + c, a_fold1(r_new, v.y, v.z), a_fold1(r_new, v.w, nx));
    }
  } else if (MODE == AIR_B2) {
    const uint4 v0 = a_ldg4(A.base + (uint64_t)c * n + 8 * j), v1 = a_ldg4(A.base + (uint64_t)c * n + 8 * j + 4);
    const Ef a = a_fold5(r_new, a_fold1(r_old, v0.x, v0.y), a_fold1(r_old, v0.z, v0.w));
    const Ef b = a_fold5(r_new, a_fold1(r_old, v1.x, v1.y), a_fold1(r_old, v1.z, v1.w));
    stage_put(pt, df, c, a, b);
    plane_put(A.dst, rows_dst, c, j, a, b);
    if (c < 2) {
      const uint32_t nx = 8 * j + 8 < n ? __ldg(A.base + (uint64_t)c * n + 8 * j + 8) : A.halo[c];
      const Ef sa = a_fold5(r_new, a_fold1(r_old, v0.y, v0.z), a_fold1(r_old, v0.w, v1.x));
      const Ef sb = a_fold5(r_new, a_fold1(r_old, v1.y, v1.z), a_fold1(r_old, v1.w, nx));
      stage_put(pt, df, EXEC_COLS + c, sa, sb);
      plane_put(A.dst, rows_dst, EXEC_COLS + c, j, sa, sb);
    }
  } else if (MODE == AIR_E0) {
    const uint64_t rows_src = (uint64_t)2 << A.m;
    Ef a, b;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint2 v = a_ldcg2(A.src + (uint64_t)(c * 5 + k) * rows_src + 2 * j);
      a.c[k] = v.x, b.c[k] = v.y;
    }
    stage_put(pt, df, c, a, b);
  } else {
    const uint64_t rows_src = (uint64_t)4 << A.m;
    Ef l0, h0, l1, h1;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint4 v = a_ldcg4(A.src + (uint64_t)(c * 5 + k) * rows_src + 4 * j);
      l0.c[k] = v.x, h0.c[k] = v.y, l1.c[k] = v.z, h1.c[k] = v.w;
    }
    const Ef a = a_fold5(r_new, l0, h0), b = a_fold5(r_new, l1, h1);
    stage_put(pt, df, c, a, b);
    plane_put(A.dst, rows_dst, c, j, a, b);
  }
}

// out[z] = sum over pairs j of eq(j) * C(row(2j) + z (row(2j+1) - row(2j))), z = 0, 2, 3, 4, 5
template <int MODE>
__global__ void __launch_bounds__(EXEC_THREADS, 3)
air_exec_round_kernel(const __grid_constant__ AirExecArgs A, const __grid_constant__ AirExecConsts X) {
  using T = typename std::conditional<MODE == AIR_B0, Fb, Ef>::type;
  constexpr int W = EXEC_ALL * (MODE == AIR_B0 ? 1 : 5);
  constexpr int STAGE_COLS = (MODE == AIR_E0 || MODE == AIR_E1) ? EXEC_ALL : EXEC_COLS;
  constexpr int ZS = W * EXEC_PAIRS;  // words per evaluation point
  extern __shared__ uint32_t sm_pt[];  // [EXEC_DEG][W][EXEC_PAIRS]
  __shared__ uint32_t sm_eq[5 * EXEC_PAIRS];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ef acc = ef_zero();
  Ef r_new = ef_zero(), r_old = ef_zero();
  if (MODE == AIR_B1 || MODE == AIR_B2 || MODE == AIR_E1) r_new = A.d->r[0];
  if (MODE == AIR_B2) r_old = A.d->r[1];
  const uint64_t pairs = (uint64_t)1 << A.m;
  const EqView eqv(A.eq_tab, A.k, A.m);
  for (uint64_t j0 = (uint64_t)blockIdx.x * EXEC_PAIRS; j0 < pairs; j0 += (uint64_t)gridDim.x * EXEC_PAIRS) {
    // stage: item = (column, pair), pair fastest so that a warp reads one column of 32 consecutive pairs
    for (int it = tid; it < (STAGE_COLS + 1) * EXEC_PAIRS; it += EXEC_THREADS) {
      const int c = it >> 5, p = it & 31;
      const uint64_t j = j0 + p;
      if (j >= pairs) continue;
      if (c < STAGE_COLS) {
        exec_stage<MODE>(A, c, j, r_new, r_old, sm_pt + p, ZS);
      } else {
        const Ef e = eqv(j);
#pragma unroll
        for (int k = 0; k < 5; k++) sm_eq[k * EXEC_PAIRS + p] = e.c[k];
      }
    }
    __syncthreads();
    if (j0 + lane < pairs) {
      Ef eq;
#pragma unroll
      for (int k = 0; k < 5; k++) eq.c[k] = sm_eq[k * EXEC_PAIRS + lane];
      acc = ef_add(acc, ef_mul(exec_air_eval<T>(SmView<T>{sm_pt + warp * ZS + lane}, X), eq));
    }
    __syncthreads();
  }
  // warp reduction; warp z owns output z
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Ef o;
#pragma unroll
    for (int c = 0; c < 5; c++) o.c[c] = __shfl_down_sync(0xffffffffu, acc.c[c], off);
    acc = ef_add(acc, o);
  }
  if (lane == 0) st_ef(A.partial + ((uint64_t)blockIdx.x * EXEC_DEG + warp) * 5, acc);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&A.d->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid < EXEC_DEG * 5) {  // word tid of the 25 output words
    uint32_t s = 0;
    for (uint32_t b = 0; b < gridDim.x; b++) s = kb_add(s, __ldcg(A.partial + (uint64_t)b * EXEC_DEG * 5 + tid));
    A.d->out[tid / 5].c[tid % 5] = s;
  }
  if (tid == 0) A.d->counter = 0;
}


// ---- round 0 from registers --------------------------------------------------------------------------------------------
// The first round reads base-field rows: a pair is 2 x 22 words and its five evaluation points are five additions per word, so
// the staging through shared memory (two barriers per 32 pairs, one warp per point) costs more than the constraints it feeds.
// Here a thread owns a pair, keeps row(2j) + z (row(2j+1) - row(2j)) in registers and walks z = 0, 2, 3, 4, 5 itself; the eq
// weight is computed once per pair.  Same sums, same partial-sum layout and last-CTA reduction as air_exec_round_kernel.
struct RegViewB0 {
  const uint32_t* cur;
  __device__ __forceinline__ Fb operator()(int c) const { return Fb{cur[c]}; }
};
constexpr int EXEC_B0_THREADS = 128;
__global__ void __launch_bounds__(EXEC_B0_THREADS, 4)
air_exec_round_b0_reg_kernel(const __grid_constant__ AirExecArgs A, const __grid_constant__ AirExecConsts X) {
  __shared__ uint32_t sm_red[EXEC_B0_THREADS / 32][EXEC_DEG * 5];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ef acc[EXEC_DEG];
#pragma unroll
  for (int z = 0; z < EXEC_DEG; z++) acc[z] = ef_zero();
  const uint64_t pairs = (uint64_t)1 << A.m, n = A.n_base;
  const EqView eqv(A.eq_tab, A.k, A.m);
  for (uint64_t j = (uint64_t)blockIdx.x * EXEC_B0_THREADS + tid; j < pairs; j += (uint64_t)gridDim.x * EXEC_B0_THREADS) {
    uint32_t cur[EXEC_ALL], d[EXEC_ALL];
#pragma unroll
    for (int c = 0; c < EXEC_COLS; c++) {
      const uint2 v = a_ldg2(A.base + (uint64_t)c * n + 2 * j);
      cur[c] = v.x;
      d[c] = kb_sub(v.y, v.x);
      if (c < 2) {  // the shifted columns: rows i + 1 of columns 0 and 1
        const uint32_t nx = 2 * j + 2 < n ? __ldg(A.base + (uint64_t)c * n + 2 * j + 2) : A.halo[c];
        cur[EXEC_COLS + c] = v.y;
        d[EXEC_COLS + c] = kb_sub(nx, v.y);
      }
    }
    const Ef eq = eqv(j);
#pragma unroll
    for (int zi = 0; zi < EXEC_DEG; zi++) {
      if (zi == 1) {
#pragma unroll
        for (int c = 0; c < EXEC_ALL; c++) cur[c] = kb_add(kb_add(cur[c], d[c]), d[c]);  // z = 2
      } else if (zi > 1) {
#pragma unroll
        for (int c = 0; c < EXEC_ALL; c++) cur[c] = kb_add(cur[c], d[c]);
      }
      acc[zi] = ef_add(acc[zi], ef_mul(exec_air_eval<Fb>(RegViewB0{cur}, X), eq));
    }
  }
#pragma unroll
  for (int zi = 0; zi < EXEC_DEG; zi++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Ef o;
#pragma unroll
      for (int c = 0; c < 5; c++) o.c[c] = __shfl_down_sync(0xffffffffu, acc[zi].c[c], off);
      acc[zi] = ef_add(acc[zi], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 5; c++) sm_red[warp][zi * 5 + c] = acc[zi].c[c];
    }
  }
  __syncthreads();
  if (tid < EXEC_DEG * 5) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < EXEC_B0_THREADS / 32; w++) s = kb_add(s, sm_red[w][tid]);
    A.partial[(uint64_t)blockIdx.x * EXEC_DEG * 5 + tid] = s;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&A.d->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid < EXEC_DEG * 5) {
    uint32_t s = 0;
    for (uint32_t b = 0; b < gridDim.x; b++) s = kb_add(s, __ldcg(A.partial + (uint64_t)b * EXEC_DEG * 5 + tid));
    A.d->out[tid / 5].c[tid % 5] = s;
  }
  if (tid == 0) A.d->counter = 0;
}

// ---- round 1 in the base field --------------------------------------------------------------------------------------
// After ONE fold the table's values are a + r0 b with a, b in the base field.  At an evaluation point z every column is
// P(z) + r0 Q(z) with P, Q base-field rows, so every constraint is a polynomial in r0 of degree <= 5 with BASE-FIELD
// coefficients: the 21 extension products of the constraint code become small polynomial products over the base field
// (140 base multiplications with delayed reduction instead of 21 x 25), and sum_k alpha^k C_k = sum_{k,i} (alpha^k r0^i)
// c_{k,i} is the same delayed accumulation as before against weights the host multiplied by the powers of r0.  Same field
// elements as evaluating in the extension field (exact arithmetic); ~1.3 k instead of ~5.6 k instructions per (pair, z) for
// the round that is half of all extension work of a session.
template <int N>
struct Pl {  // c[0] + c[1] t + ... + c[N-1] t^(N-1)
  uint32_t c[N];
};
template <int N, int M>
__device__ __forceinline__ Pl<(N > M ? N : M)> operator+(const Pl<N>& a, const Pl<M>& b) {
  Pl<(N > M ? N : M)> r;
#pragma unroll
  for (int i = 0; i < (N > M ? N : M); i++) r.c[i] = i < N ? (i < M ? kb_add(a.c[i], b.c[i]) : a.c[i]) : b.c[i];
  return r;
}
template <int N, int M>
__device__ __forceinline__ Pl<(N > M ? N : M)> operator-(const Pl<N>& a, const Pl<M>& b) {
  Pl<(N > M ? N : M)> r;
#pragma unroll
  for (int i = 0; i < (N > M ? N : M); i++) r.c[i] = i < N ? (i < M ? kb_sub(a.c[i], b.c[i]) : a.c[i]) : kb_neg(b.c[i]);
  return r;
}
template <int N>
__device__ __forceinline__ Pl<N> operator-(const Pl<N>& a) {
  Pl<N> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.c[i] = kb_neg(a.c[i]);
  return r;
}
// product with one reduction per output coefficient
template <int N, int M>
__device__ __forceinline__ Pl<N + M - 1> operator*(const Pl<N>& a, const Pl<M>& b) {
  Pl<N + M - 1> r;
#pragma unroll
  for (int k = 0; k < N + M - 1; k++) {
    uint64_t acc = 0;
    int terms = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j < 0 || j >= M) continue;
      if (terms == 4) acc = kb_fold(acc), terms = 0;
      acc = mad_wide(a.c[i], b.c[j], acc);
      terms++;
    }
    r.c[k] = kb_canon(kb_redc_lazy(kb_fold(acc)));
  }
  return r;
}
template <int N>
__device__ __forceinline__ Pl<N> pl_sub_one(Pl<N> a) {
  a.c[0] = kb_sub(a.c[0], KB_R1);
  return a;
}
template <int N>
__device__ __forceinline__ Pl<N> pl_add_one(Pl<N> a) {
  a.c[0] = kb_add(a.c[0], KB_R1);
  return a;
}
template <int N>
__device__ __forceinline__ Pl<N> pl_dbl(const Pl<N>& a) { return a + a; }
template <int N>
__device__ __forceinline__ Pl<N> pl_halve(Pl<N> a) {
#pragma unroll
  for (int i = 0; i < N; i++) a.c[i] = kb_halve(a.c[i]);
  return a;
}

struct AirExecConstsB1 {
  EfRows w[17][6];  // weight of term t (AirExecConsts::w) times r0^i
  Ef k0;
};
template <int N>
__device__ __forceinline__ void mac_poly(AlphaAcc& acc, const EfRows (&w)[6], const Pl<N>& v) {
#pragma unroll
  for (int i = 0; i < N; i++) acc.mac(w[i], Fb{v.c[i]});
}

// the thread's staged point: word 2c = P_c(z), word 2c + 1 = Q_c(z)
struct SmViewB1 {
  const uint32_t* pt;
  __device__ __forceinline__ Pl<2> operator()(int c) const { return Pl<2>{{pt[(2 * c) * EXEC_PAIRS], pt[(2 * c + 1) * EXEC_PAIRS]}}; }
};

// ExecutionTable::eval (execution/air.rs:56-130) on columns that are degree-1 polynomials in the first challenge
__device__ __forceinline__ Ef exec_air_eval_b1(const SmViewB1& col, const AirExecConstsB1& X) {
  AlphaAcc acc;
  const Pl<2> fp = col(1);
  const Pl<2> op_a = col(8), op_b = col(9), op_c = col(10);
  const Pl<2> val_a = col(5), val_b = col(6), val_c = col(7);
  const Pl<2> flag_a = col(11), flag_b = col(12), flag_c = col(13), flag_c_fp = col(14), flag_ab_fp = col(15);
  const Pl<2> fp_op_a = fp + op_a, fp_op_b = fp + op_b, fp_op_c = fp + op_c;
  const Pl<3> nu_a = val_a + flag_a * (op_a - val_a) + flag_ab_fp * (fp_op_a - val_a);
  const Pl<3> nu_b = val_b + flag_b * (op_b - val_b) + flag_ab_fp * (fp_op_b - val_b);
  const Pl<3> nu_c = val_c + flag_c * (op_c - val_c) + flag_c_fp * (fp_op_c - val_c);
  const Pl<2> aux = col(18), mul = col(16), jump = col(17);
  const Pl<3> aux2 = aux * aux;
  const Pl<3> add = pl_dbl(aux) - aux2;
  const Pl<3> deref = pl_halve(aux2 - aux);
  const Pl<3> is_precompile = -pl_sub_one(add + mul + deref + jump);
  mac_poly(acc, X.w[0], col(19));
  mac_poly(acc, X.w[1], nu_a);
  mac_poly(acc, X.w[2], nu_b);
  mac_poly(acc, X.w[3], nu_c);
  mac_poly(acc, X.w[4], is_precompile);
  const Pl<2> om_a = -pl_sub_one(flag_a + flag_ab_fp);
  const Pl<2> om_b = -pl_sub_one(flag_b + flag_ab_fp);
  const Pl<2> om_c = -pl_sub_one(flag_c + flag_c_fp);
  const Pl<2> addr_b = col(3);
  mac_poly(acc, X.w[5], om_a * (col(2) - fp_op_a));
  mac_poly(acc, X.w[6], om_b * (addr_b - fp_op_b));
  mac_poly(acc, X.w[7], om_c * (col(4) - fp_op_c));
  mac_poly(acc, X.w[8], add * (nu_b - (nu_a + nu_c)));
  mac_poly(acc, X.w[9], mul * (nu_b - nu_a * nu_c));
  mac_poly(acc, X.w[10], deref * (addr_b - (val_a + op_b)));
  mac_poly(acc, X.w[11], deref * (val_b - nu_c));
  const Pl<4> jc = jump * nu_a;
  mac_poly(acc, X.w[12], jc * pl_sub_one(nu_a));
  const Pl<2> pc_shift = col(20), fp_shift = col(21);
  mac_poly(acc, X.w[13], jc * (pc_shift - nu_b));
  mac_poly(acc, X.w[14], jc * (fp_shift - nu_c));
  const Pl<4> njc = -pl_sub_one(jc);
  mac_poly(acc, X.w[15], njc * (pc_shift - pl_add_one(col(0))));
  mac_poly(acc, X.w[16], njc * (fp_shift - fp));
  return acc.finish(X.k0);
}

// round 1 (one pending challenge on the base table): out[z] as in air_exec_round_kernel, columns staged as (P(z), Q(z))
__global__ void __launch_bounds__(EXEC_THREADS, 3)
air_exec_round_b1_kernel(const __grid_constant__ AirExecArgs A, const __grid_constant__ AirExecConstsB1 X) {
  constexpr int W = 2 * EXEC_ALL;
  constexpr int ZS = W * EXEC_PAIRS;
  __shared__ uint32_t sm_pt[EXEC_DEG * ZS];  // 27.5 KiB
  __shared__ uint32_t sm_eq[5 * EXEC_PAIRS];
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ef acc = ef_zero();
  const uint64_t pairs = (uint64_t)1 << A.m, n = A.n_base;
  const EqView eqv(A.eq_tab, A.k, A.m);
  // The evaluation is short (~1.3 k instructions), so the global loads of the NEXT 32 pairs are issued before the current ones
  // are evaluated and land in registers meanwhile: item it = tid + 160 s, s < 5, is (column it >> 5, pair it & 31)
  constexpr int SLOTS = ((EXEC_COLS + 1) * EXEC_PAIRS + EXEC_THREADS - 1) / EXEC_THREADS;
  uint4 pre[SLOTS];
  uint32_t pre_nx[SLOTS];
  auto prefetch = [&](uint64_t j0) {
#pragma unroll
    for (int sl = 0; sl < SLOTS; sl++) {
      const int it = tid + sl * EXEC_THREADS;
      const int c = it >> 5;
      const uint64_t j = j0 + (it & 31);
      pre[sl] = make_uint4(0, 0, 0, 0);
      pre_nx[sl] = 0;
      if (c < EXEC_COLS && j < pairs) {
        pre[sl] = a_ldg4(A.base + (uint64_t)c * n + 4 * j);
        if (c < 2) pre_nx[sl] = 4 * j + 4 < n ? __ldg(A.base + (uint64_t)c * n + 4 * j + 4) : A.halo[c];
      }
    }
  };
  const uint64_t stride = (uint64_t)gridDim.x * EXEC_PAIRS;
  uint64_t j0 = (uint64_t)blockIdx.x * EXEC_PAIRS;
  if (j0 < pairs) prefetch(j0);
  for (; j0 < pairs; j0 += stride) {
#pragma unroll
    for (int sl = 0; sl < SLOTS; sl++) {
      const int it = tid + sl * EXEC_THREADS;
      const int c = it >> 5, p = it & 31;
      const uint64_t j = j0 + p;
      if (it >= (EXEC_COLS + 1) * EXEC_PAIRS || j >= pairs) continue;
      uint32_t* pt = sm_pt + p;
      if (c < EXEC_COLS) {
        // base rows 4j .. 4j+3 = (A0, A1, B0, B1): P(z) = A0 + z (B0 - A0), Q(z) = (A1 - A0) + z ((B1 - B0) - (A1 - A0))
        const uint4 v = pre[sl];
        stage_word(pt, ZS, 2 * c, v.x, v.z);
        stage_word(pt, ZS, 2 * c + 1, kb_sub(v.y, v.x), kb_sub(v.w, v.z));
        if (c < 2) {  // shifted column: rows 4j+1 .. 4j+4
          const uint32_t nx = pre_nx[sl];
          stage_word(pt, ZS, 2 * (EXEC_COLS + c), v.y, v.w);
          stage_word(pt, ZS, 2 * (EXEC_COLS + c) + 1, kb_sub(v.z, v.y), kb_sub(nx, v.w));
        }
      } else {
        const Ef e = eqv(j);
#pragma unroll
        for (int k = 0; k < 5; k++) sm_eq[k * EXEC_PAIRS + p] = e.c[k];
      }
    }
    __syncthreads();
    if (j0 + stride < pairs) prefetch(j0 + stride);
    if (j0 + lane < pairs) {
      Ef eq;
#pragma unroll
      for (int k = 0; k < 5; k++) eq.c[k] = sm_eq[k * EXEC_PAIRS + lane];
      acc = ef_add(acc, ef_mul(exec_air_eval_b1(SmViewB1{sm_pt + warp * ZS + lane}, X), eq));
    }
    __syncthreads();
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Ef o;
#pragma unroll
    for (int c = 0; c < 5; c++) o.c[c] = __shfl_down_sync(0xffffffffu, acc.c[c], off);
    acc = ef_add(acc, o);
  }
  if (lane == 0) st_ef(A.partial + ((uint64_t)blockIdx.x * EXEC_DEG + warp) * 5, acc);
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(&A.d->counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid < EXEC_DEG * 5) {
    uint32_t s = 0;
    for (uint32_t b = 0; b < gridDim.x; b++) s = kb_add(s, __ldcg(A.partial + (uint64_t)b * EXEC_DEG * 5 + tid));
    A.d->out[tid / 5].c[tid % 5] = s;
  }
  if (tid == 0) A.d->counter = 0;
}

__global__ void air_exec_final_kernel(int mode, AirExecArgs A, uint32_t* out) {
  const int c = threadIdx.x;
  if (c >= EXEC_ALL) return;
  const Ef r_new = A.d->r[0], r_old = A.d->r[1];
  const uint64_t n = A.n_base;
  const int bc = c < EXEC_COLS ? c : c - EXEC_COLS;  // base column behind column c
  const int sh = c < EXEC_COLS ? 0 : 1;              // shifted columns read one row further
  auto base_at = [&](uint64_t i) { return i < n ? A.base[(uint64_t)bc * n + i] : A.halo[bc]; };
  Ef v;
  if (mode == AIR_B0) {
    v = ef_from_base(base_at(sh));
  } else if (mode == AIR_B1) {
    v = a_fold1(r_new, base_at(sh), base_at(1 + sh));
  } else if (mode == AIR_B2) {
    v = a_fold5(r_new, a_fold1(r_old, base_at(sh), base_at(1 + sh)), a_fold1(r_old, base_at(2 + sh), base_at(3 + sh)));
  } else if (mode == AIR_E0) {
    for (int k = 0; k < 5; k++) v.c[k] = A.src[c * 5 + k];
  } else {
    Ef lo, hi;
    for (int k = 0; k < 5; k++) lo.c[k] = A.src[(c * 5 + k) * 2], hi.c[k] = A.src[(c * 5 + k) * 2 + 1];
    v = a_fold5(r_new, lo, hi);
  }
  st_ef(out + 5 * c, v);
}

__global__ void __launch_bounds__(512) air_eq_tables_kernel(const uint32_t* __restrict__ point, uint32_t k, Ef scale, uint32_t* tab) {
  eqtab_build(tab, reinterpret_cast<const Ef*>(point), k, scale);
}

cudaError_t air_build_eq_tables(cudaStream_t stream, const uint32_t* d_eq_point, uint32_t n_vars, const uint32_t eq_scale[5],
                                uint32_t* d_tab) {
  Ef sc{{KB_R1, 0, 0, 0, 0}};
  if (eq_scale)
    for (int k = 0; k < 5; k++) sc.c[k] = eq_scale[k];
  air_eq_tables_kernel<<<1, 512, 0, stream>>>(d_eq_point, n_vars, sc, d_tab);
  count_launch();
  return cudaGetLastError();
}

static Ef host_ef(const uint32_t* p) {
  Ef e;
  for (int c = 0; c < 5; c++) e.c[c] = p[c];
  return e;
}

cudaError_t air_exec_round(cudaStream_t stream, int mode, const AirExecArgs& a, const uint32_t* alpha_powers, const uint32_t* la,
                           uint32_t n_la, const uint32_t beta[5], const uint32_t* r0_host) {
  if (n_la < 5 || mode < 0 || mode > AIR_E1) return cudaErrorInvalidValue;
  AirExecConsts X;
  const Ef a0 = host_ef(alpha_powers), be = host_ef(beta);
  const Ef a0b = ef_mul(a0, be);
  for (int i = 0; i < 4; i++) X.w[i] = ef_rows(ef_mul(a0b, host_ef(la + 5 * i)));
  X.w[4] = ef_rows(a0);
  for (int k = 1; k < 13; k++) X.w[4 + k] = ef_rows(host_ef(alpha_powers + 5 * k));
  // la_last * DOMAINSEP with LOGUP_PRECOMPILE_DOMAINSEP = 1 (lean_vm/src/core/constants.rs:5)
  X.k0 = ef_mul(a0b, host_ef(la + 5 * (n_la - 1)));

  const uint64_t pairs = (uint64_t)1 << a.m;
  uint64_t blocks = (pairs + EXEC_PAIRS - 1) / EXEC_PAIRS;
  if (blocks > (uint64_t)AIR_MAX_BLOCKS) blocks = AIR_MAX_BLOCKS;
  const unsigned g = (unsigned)blocks;
  if (mode == AIR_B1 && r0_host && getenv("LM_AIR_B1_EXT") == nullptr) {
    // round 1 with base-field polynomial arithmetic: weights times the powers of the first challenge
    static AirExecConstsB1 Y;  // 3.7 KiB: keep it off the stack
    const Ef r0 = host_ef(r0_host);
    Ef pw[6];
    pw[0] = Ef{{KB_R1, 0, 0, 0, 0}};
    for (int i = 1; i < 6; i++) pw[i] = ef_mul(pw[i - 1], r0);
    Ef wts[17];
    for (int i = 0; i < 4; i++) wts[i] = ef_mul(a0b, host_ef(la + 5 * i));
    wts[4] = a0;
    for (int k = 1; k < 13; k++) wts[4 + k] = host_ef(alpha_powers + 5 * k);
    for (int t = 0; t < 17; t++)
      for (int i = 0; i < 6; i++) Y.w[t][i] = ef_rows(ef_mul(wts[t], pw[i]));
    Y.k0 = X.k0;
    air_exec_round_b1_kernel<<<g, EXEC_THREADS, 0, stream>>>(a, Y);
    count_launch();
    return cudaGetLastError();
  }
  const size_t smem_b = (size_t)EXEC_DEG * EXEC_ALL * EXEC_PAIRS * sizeof(uint32_t);  // 13.75 KiB
  const size_t smem_ef = 5 * smem_b;                                                    // 68.75 KiB: three CTAs per SM
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(air_exec_round_kernel<AIR_B1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ef);
    cudaFuncSetAttribute(air_exec_round_kernel<AIR_B2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ef);
    cudaFuncSetAttribute(air_exec_round_kernel<AIR_E0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ef);
    cudaFuncSetAttribute(air_exec_round_kernel<AIR_E1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ef);
    attr_set = true;
  }
  static const bool b0_smem = getenv("LM_AIR_B0_SMEM") != nullptr;  // the staged round-0 kernel, for cross-checks
  if (mode == AIR_B0 && !b0_smem) {
    uint64_t nb = (pairs + EXEC_B0_THREADS - 1) / EXEC_B0_THREADS;
    if (nb > (uint64_t)AIR_MAX_BLOCKS) nb = AIR_MAX_BLOCKS;
    air_exec_round_b0_reg_kernel<<<(unsigned)nb, EXEC_B0_THREADS, 0, stream>>>(a, X);
    count_launch();
    return cudaGetLastError();
  }
  switch (mode) {
    case AIR_B0: air_exec_round_kernel<AIR_B0><<<g, EXEC_THREADS, smem_b, stream>>>(a, X); break;
    case AIR_B1: air_exec_round_kernel<AIR_B1><<<g, EXEC_THREADS, smem_ef, stream>>>(a, X); break;
    case AIR_B2: air_exec_round_kernel<AIR_B2><<<g, EXEC_THREADS, smem_ef, stream>>>(a, X); break;
    case AIR_E0: air_exec_round_kernel<AIR_E0><<<g, EXEC_THREADS, smem_ef, stream>>>(a, X); break;
    default: air_exec_round_kernel<AIR_E1><<<g, EXEC_THREADS, smem_ef, stream>>>(a, X); break;
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t air_exec_final(cudaStream_t stream, int mode, const AirExecArgs& a, uint32_t* d_out) {
  air_exec_final_kernel<<<1, 32, 0, stream>>>(mode, a, d_out);
  count_launch();
  return cudaGetLastError();
}

// ---- generic pieces ------------------------------------------------------------------------------------------------
// fold the least-significant variable of every column: out[c][.][j] = in[c][2j] + r (in[c][2j+1] - in[c][2j])
template <int DIM>
__global__ void air_fold_lsb_kernel(const uint32_t* in, uint64_t n, int n_cols, Ef r, uint32_t* out) {
  const uint64_t half = n / 2;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= half * n_cols) return;
  const uint64_t c = idx / half, j = idx % half;
  Ef o;
  if (DIM == 1) {
    const uint2 ab = *reinterpret_cast<const uint2*>(in + c * n + 2 * j);
    o = a_fold1(r, ab.x, ab.y);
  } else {
    Ef a, b;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint2 v = *reinterpret_cast<const uint2*>(in + (c * 5 + k) * n + 2 * j);
      a.c[k] = v.x, b.c[k] = v.y;
    }
    o = a_fold5(r, a, b);
  }
#pragma unroll
  for (int k = 0; k < 5; k++) out[(c * 5 + k) * half + j] = o.c[k];
}

// shifted[i] = col[i + 1], last row repeated (air_sumcheck.rs:683-694)
__global__ void air_shift_kernel(const uint32_t* __restrict__ col, uint64_t n, uint32_t* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __ldg(col + (i + 1 < n ? i + 1 : n - 1));
}

__global__ void air_aos_to_planes_kernel(const uint32_t* __restrict__ aos, uint64_t total, uint64_t n, uint32_t* __restrict__ planes) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (column, row)
  if (idx >= total) return;
  const uint64_t c = idx / n, i = idx % n;
#pragma unroll
  for (int k = 0; k < 5; k++) planes[(c * 5 + k) * n + i] = __ldg(aos + idx * 5 + k);
}

cudaError_t air_shift_column(cudaStream_t stream, const uint32_t* d_col, uint64_t n, uint32_t* d_out) {
  air_shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_col, n, d_out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t air_aos_to_planes(cudaStream_t stream, const uint32_t* d_aos, uint32_t n_cols, uint64_t n, uint32_t* d_planes) {
  const uint64_t total = (uint64_t)n_cols * n;
  air_aos_to_planes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_aos, total, n, d_planes);
  count_launch();
  return cudaGetLastError();
}

size_t air_round_scratch_words() {
  // partial sums: AIR_MAX_BLOCKS CTAs x up to 10 values (air_generic.cu) or x 5 values (execution table)
  return (size_t)AIR_MAX_BLOCKS * 10 * 5 + 64;
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

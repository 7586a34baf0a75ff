// AIR sumcheck rounds for the extension_op and poseidon16 tables, and the Poseidon16 trace generator, on sm_100a.
//
// Device replacement for
//   crates/lean_vm/src/tables/extension_op/air.rs:44-163    ExtensionOpPrecompile::eval (29 + 13 shifted columns, degree 6)
//   crates/lean_vm/src/tables/poseidon_16/mod.rs:294-548    Poseidon16Precompile::eval + eval_poseidon1_16 (109 columns, degree 10)
//   crates/lean_vm/src/tables/poseidon_16/trace_gen.rs:10-165  fill_trace_poseidon_16
//   crates/sub_protocols/src/air_sumcheck.rs:403-634        compute_raw_poly_impl / compute_raw_poly_degree_split
//
// These tables are too wide to hold a row pair in registers (poseidon16 in extension rounds: 109 x 5 x 2 words), so
// the constraint code STREAMS: it asks a column view for "column c at z" exactly where the reference reads flat[c],
// and the view forms lo + z (hi - lo) from the two adjacent rows on the fly (one 8-byte load in the base round, five
// in extension rounds; repeated z values hit L1/L2).  Live state is the 16-lane Poseidon state (or the handful of
// extension_op values), not the row.  The reference evaluates the degree-3 partial-round constraints of poseidon16 on
// fewer points and extrapolates (compute_raw_poly_degree_split); the round polynomial is the same field elements, and
// here every constraint is simply evaluated at all z = 0, 2, .., d.
#include <cuda_runtime.h>
#include <cstdint>
#include "air.h"
#include "air_values.cuh"
#include "eqtab.cuh"
#include "kb.cuh"
#include "launch_count.h"
#include "poly.h"
#include "poseidon1.cuh"

namespace lm {

struct P16Sparse {
  uint32_t RC_FULL[8][16];
  uint32_t FIRST_RC[16];
  uint32_t M_I[16][16];
  uint32_t FIRST_ROW[20][16];
  uint32_t V[20][16];
  uint32_t SCALAR_RC[20];
};
static __constant__ P16Sparse c_p16 =
#include "poseidon1_sparse_tables.inc"
    ;

// alpha powers for up to 100 constraints + the bus constants; passed as a __grid_constant__ kernel parameter
// (constant bank, uniform reads)
constexpr int AIRG_MAX_ALPHA = 100;
struct AirExtraBig {
  Ef alpha[AIRG_MAX_ALPHA];
  Ef la[4];
  Ef la_last;
  Ef beta;
  int bus;
};

// ---- more generic value helpers ---------------------------------------------------------------------------
__device__ __forceinline__ Fb add_c(Fb a, uint32_t c) { return Fb{kb_add(a.v, c)}; }
__device__ __forceinline__ Ef add_c(Ef a, uint32_t c) { return ef_add_base(a, c); }
__device__ __forceinline__ Fb mul_c(Fb a, uint32_t c) { return Fb{kb_mul(a.v, c)}; }
__device__ __forceinline__ Ef mul_c(const Ef& a, uint32_t c) { return ef_mul_base(a, c); }
__device__ __forceinline__ Fb zero_of(Fb) { return Fb{0}; }
__device__ __forceinline__ Ef zero_of(const Ef&) { return ef_zero(); }
template <class T>
__device__ __forceinline__ T cube(const T& x) {
  return x * x * x;
}
template <class T>
__device__ __forceinline__ T bool_check(const T& x) {
  return (-sub_one(x)) * x;  // (1 - x) x   (field.rs:207)
}
template <uint32_t V>
struct MontySmall {
  static constexpr uint32_t value = (uint32_t)(((uint64_t)V << 32) % KB_P);
};

// sum_j s[j] * row[j] with Montgomery-form constants, delayed reduction (fold every 4 products)
__device__ __forceinline__ Fb dot16c(const Fb s[16], const uint32_t* row) {
  uint64_t acc = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    if (j && (j & 3) == 0) acc = kb_fold(acc);
    acc = mad_wide(s[j].v, row[j], acc);
  }
  return Fb{kb_canon(kb_redc_lazy(kb_fold(acc)))};
}
__device__ __forceinline__ Ef dot16c(const Ef s[16], const uint32_t* row) {
  Ef r;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    uint64_t acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (j && (j & 3) == 0) acc = kb_fold(acc);
      acc = mad_wide(s[j].c[k], row[j], acc);
    }
    r.c[k] = kb_canon(kb_redc_lazy(kb_fold(acc)));
  }
  return r;
}

// circulant MDS on 16 lanes through the exact FP64 evaluation of poseidon1.cuh: p1_mds_redc returns
// 4 (C x) 2^-32 lazily reduced; multiplying by 2^64 / 4 in Montgomery form restores C x.
constexpr uint32_t MDS_UNSCALE = (uint32_t)((((unsigned __int128)1 << 62)) % KB_P);
__device__ __forceinline__ void mds16(Fb s[16]) {
  uint32_t a[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = s[i].v;
  p1_mds_redc<16>(a, nullptr, nullptr, o);
#pragma unroll
  for (int i = 0; i < 16; i++) s[i].v = kb_mul(o[i], MDS_UNSCALE);
}
__device__ __forceinline__ void mds16(Ef s[16]) {
#pragma unroll 1
  for (int k = 0; k < 5; k++) {
    uint32_t a[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      // select coordinate k without dynamic register indexing
      uint32_t v = s[i].c[0];
      v = k == 1 ? s[i].c[1] : v;
      v = k == 2 ? s[i].c[2] : v;
      v = k == 3 ? s[i].c[3] : v;
      v = k == 4 ? s[i].c[4] : v;
      a[i] = v;
    }
    p1_mds_redc<16>(a, nullptr, nullptr, o);
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t r = kb_mul(o[i], MDS_UNSCALE);
      if (k == 0) s[i].c[0] = r;
      if (k == 1) s[i].c[1] = r;
      if (k == 2) s[i].c[2] = r;
      if (k == 3) s[i].c[3] = r;
      if (k == 4) s[i].c[4] = r;
    }
  }
}

// ---- column view: the row pair (2j, 2j+1) of every column, evaluated at z ---------------------------------
template <class T, int DIM>
struct ColView;
template <>
struct ColView<Fb, 1> {
  const uint32_t* cols;
  uint64_t n, j;
  uint32_t zm;  // Montgomery form of z
  __device__ __forceinline__ Fb operator()(int c) const {
    const uint2 ab = __ldg(reinterpret_cast<const uint2*>(cols + (uint64_t)c * n + 2 * j));
    return Fb{kb_add(ab.x, kb_mul(zm, kb_sub(ab.y, ab.x)))};
  }
};
template <>
struct ColView<Ef, 5> {
  const uint32_t* cols;  // coefficient planes u32[c][5][n] (air.h)
  uint64_t n, j;
  uint32_t zm;
  __device__ __forceinline__ Ef operator()(int c) const {
    Ef r;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(cols + (uint64_t)(c * 5 + k) * n + 2 * j));
      r.c[k] = kb_add(v.x, kb_mul(zm, kb_sub(v.y, v.x)));
    }
    return r;
  }
};

// accumulator of sum_k alpha^k constraint_k with the reference's running constraint index (normal.rs:49-62)
template <class T>
struct Folder {
  const AirExtraBig& X;
  Ef acc;
  int idx;
  __device__ __forceinline__ explicit Folder(const AirExtraBig& x) : X(x), acc(ef_zero()), idx(0) {}
  __device__ __forceinline__ void assert_zero(const T& v) {
    acc = acc + scale(X.alpha[idx], v);
    idx++;
  }
  // eval_virtual_bus_column (tables/utils.rs:5-21); BUS = false instantiations only declare the values
  __device__ __forceinline__ void bus(const T& flag, const T& d0, const T& d1, const T& d2, const T& d3) {
    if (!X.bus) return;
    Ef s = scale(X.la[0], d0) + scale(X.la[1], d1) + scale(X.la[2], d2) + scale(X.la[3], d3) + X.la_last;
    acc = acc + ef_mul(X.alpha[idx], add_val(ef_mul(s, X.beta), flag));
    idx++;
  }
};

// ---- extension_op -----------------------------------------------------------------------------------------
struct ExtOpAir {
  static constexpr int COLS = 29, SHIFT = 13, DEG = 6;
  enum { IS_BE, START, LEN, FLAG_ADD, FLAG_MUL, FLAG_POLY_EQ, IDX_A, IDX_B, COMP, IDX_RES = 13, VA = 14, VB = 19, VRES = 24 };

  // quintic_mul on AIR values (extension.rs:531-548 via quintic_mul_air)
  template <class T>
  static __device__ __forceinline__ void qmul(const T a[5], const T b[5], T out[5]) {
    T d[9];
#pragma unroll
    for (int k = 0; k < 9; k++) d[k] = zero_of(a[0]);
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
      for (int j = 0; j < 5; j++) d[i + j] = d[i + j] + a[i] * b[j];
    out[0] = d[0] + d[5] - d[8];
    out[1] = d[1] + d[6];
    out[2] = d[2] + d[7] - d[5] + d[8];
    out[3] = d[3] - d[6] + d[8];
    out[4] = d[4] - d[7];
  }

  template <class T, class View>
  static __device__ __forceinline__ Ef eval(const View& col, const AirExtraBig& X) {
    Folder<T> f(X);
    const T is_be = col(IS_BE), start = col(START), len = col(LEN);
    const T flag_add = col(FLAG_ADD), flag_mul = col(FLAG_MUL), flag_poly_eq = col(FLAG_POLY_EQ);
    const T idx_a = col(IDX_A), idx_b = col(IDX_B);
    const T start_shift = col(COLS + START);
    {
      const T active = flag_add + flag_mul + flag_poly_eq;
      const T aux = mul_c(is_be, MontySmall<4>::value) + mul_c(flag_add, MontySmall<8>::value) + mul_c(flag_mul, MontySmall<16>::value) +
                    mul_c(flag_poly_eq, MontySmall<32>::value) + mul_c(len, MontySmall<64>::value);
      f.bus(start * active, aux, idx_a, idx_b, col(IDX_RES));
    }
    const T is_ee = -sub_one(is_be);
    const T nss = -sub_one(start_shift);  // not_start_shift
    T va[5], vb[5], comp[5], tail[5], cshift[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      va[k] = k == 0 ? col(VA) : col(VA + k) * is_ee;
      vb[k] = col(VB + k);
      comp[k] = col(COMP + k);
      cshift[k] = col(COLS + COMP + k);
      tail[k] = cshift[k] * nss;
    }
    f.assert_zero(bool_check(is_be));
    f.assert_zero(bool_check(start));
    f.assert_zero(bool_check(flag_add));
    f.assert_zero(bool_check(flag_mul));
    f.assert_zero(bool_check(flag_poly_eq));
#pragma unroll
    for (int k = 0; k < 5; k++) f.assert_zero((comp[k] - (va[k] + vb[k] + tail[k])) * flag_add);
    T prod[5];
    qmul(va, vb, prod);
#pragma unroll
    for (int k = 0; k < 5; k++) f.assert_zero((comp[k] - (prod[k] + tail[k])) * flag_mul);
    {
      T pe[5], cso[5], per[5];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        pe[k] = dbl(prod[k]) - va[k] - vb[k];
        cso[k] = tail[k];
      }
      pe[0] = add_one(pe[0]);
      cso[0] = cso[0] + start_shift;
      qmul(pe, cso, per);
#pragma unroll
      for (int k = 0; k < 5; k++) f.assert_zero((comp[k] - per[k]) * flag_poly_eq);
    }
#pragma unroll
    for (int k = 0; k < 5; k++) f.assert_zero((comp[k] - col(VRES + k)) * start);
    f.assert_zero(nss * sub_one(len - col(COLS + LEN)));
    f.assert_zero(nss * (is_be - col(COLS + IS_BE)));
    f.assert_zero(nss * (flag_add - col(COLS + FLAG_ADD)));
    f.assert_zero(nss * (flag_mul - col(COLS + FLAG_MUL)));
    f.assert_zero(nss * (flag_poly_eq - col(COLS + FLAG_POLY_EQ)));
    const T a_inc = is_be + mul_c(is_ee, MontySmall<5>::value);
    f.assert_zero(nss * (col(COLS + IDX_A) - idx_a - a_inc));
    f.assert_zero(nss * add_c(col(COLS + IDX_B) - idx_b, KB_P - MontySmall<5>::value));
    f.assert_zero(start_shift * sub_one(len));
    return f.acc;
  }
};

// ---- poseidon16 -------------------------------------------------------------------------------------------
struct Poseidon16Air {
  static constexpr int COLS = 109, SHIFT = 0, DEG = 10;
  enum { FLAG, INDEX_B, INDEX_RES, FLAG_HALF, FLAG_HARD, OFFSET_HARD, EFF_FIRST, EFF_SECOND, FLAG_PERMUTE, INPUTS = 9,
         BEGIN = 25, PARTIAL = 57, END = 77, OUT_LEFT = 93, OUT_RIGHT = 101 };

  template <class T>
  static __device__ __forceinline__ void two_full_rounds(T s[16], int r) {
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int i = 0; i < 16; i++) s[i] = cube(add_c(s[i], c_p16.RC_FULL[r + h][i]));
      mds16(s);
    }
  }

  template <class T, class View>
  static __device__ __forceinline__ Ef eval(const View& col, const AirExtraBig& X) {
    Folder<T> f(X);
    const T flag_half = col(FLAG_HALF), flag_perm = col(FLAG_PERMUTE);
    {
      const T flag = col(FLAG), flag_hard = col(FLAG_HARD), offset = col(OFFSET_HARD), eff_first = col(EFF_FIRST);
      const T pdata = add_one(mul_c(flag_half, MontySmall<4>::value) + mul_c(flag_hard, MontySmall<8>::value) +
                              mul_c(flag_hard * offset, MontySmall<16>::value) + mul_c(flag_perm, MontySmall<2>::value));
      const T om_hard = -sub_one(flag_hard);
      const T index_a = col(EFF_SECOND) - mul_c(om_hard, MontySmall<4>::value);
      f.bus(flag, pdata, index_a, col(INDEX_B), col(INDEX_RES));
      f.assert_zero(bool_check(flag));
      f.assert_zero(bool_check(flag_half));
      f.assert_zero(bool_check(flag_hard));
      f.assert_zero(bool_check(flag_perm));
      f.assert_zero(flag_perm * (flag_half + flag_hard));
      f.assert_zero(flag_hard * (offset - eff_first));
      f.assert_zero(om_hard * (index_a - eff_first));
    }
    T s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = col(INPUTS + i);
#pragma unroll 1
    for (int r = 0; r < 2; r++) {
      two_full_rounds(s, 2 * r);
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const T post = col(BEGIN + 16 * r + i);
        f.assert_zero(s[i] - post);
        s[i] = post;
      }
    }
    // sparse partial rounds (mod.rs:384-420)
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = add_c(s[i], c_p16.FIRST_RC[i]);
    {
      T t[16];
#pragma unroll
      for (int i = 0; i < 16; i++) t[i] = dot16c(s, c_p16.M_I[i]);
#pragma unroll
      for (int i = 0; i < 16; i++) s[i] = t[i];
    }
#pragma unroll 1
    for (int r = 0; r < 20; r++) {
      const T post = col(PARTIAL + r);
      f.assert_zero(cube(s[0]) - post);
      s[0] = r < 19 ? add_c(post, c_p16.SCALAR_RC[r]) : post;
      const T old = s[0];
      const T dot = dot16c(s, c_p16.FIRST_ROW[r]);
#pragma unroll
      for (int i = 1; i < 16; i++) s[i] = s[i] + mul_c(old, c_p16.V[r][i - 1]);
      s[0] = dot;
    }
    two_full_rounds(s, 4);
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const T post = col(END + i);
      f.assert_zero(s[i] - post);
      s[i] = post;
    }
    two_full_rounds(s, 6);
    const T not_perm = -sub_one(flag_perm);
    const T last4 = not_perm - flag_half;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const T out_l = col(OUT_LEFT + i);
      f.assert_zero((i < 4 ? not_perm : last4) * (s[i] + col(INPUTS + i) - out_l));
      f.assert_zero(flag_perm * (s[i] - out_l));
      f.assert_zero(flag_perm * (s[i + 8] - col(OUT_RIGHT + i)));
    }
    return f.acc;
  }
};

// ---- the streaming round kernel -----------------------------------------------------------------------------
// partial[blockIdx.x][zi] = sum over this CTA's row pairs j of eq(j) * C(row pair j at z), z = 0, 2, 3, .., DEG.
// One warp per evaluation point: a CTA is DEG warps working on the same 32 row pairs, so a round over a handful of pairs
// costs one constraint evaluation of latency instead of DEG of them in sequence (the last ~10 rounds of every session
// are such rounds, and the poseidon16 body is ~50 k instructions), and the warps of a CTA share their column loads in L1.
template <class Air, class T, int DIM>
__global__ void __launch_bounds__(32 * Air::DEG)
air_stream_round_kernel(const uint32_t* __restrict__ cols, uint64_t n, uint64_t half, const uint32_t* __restrict__ eq_tab, uint32_t k_vars,
                        uint32_t m_vars, const __grid_constant__ AirExtraBig X, uint32_t* __restrict__ partial) {
  const int lane = threadIdx.x & 31, zi = threadIdx.x >> 5;
  const uint32_t z = zi == 0 ? 0u : (uint32_t)zi + 1u;
  const uint32_t zm = kb_mul(z, KB_R2);
  Ef acc = ef_zero();
  const EqView eqv(eq_tab, k_vars, m_vars);
  for (uint64_t j = (uint64_t)blockIdx.x * 32 + lane; j < half; j += (uint64_t)gridDim.x * 32) {
    const ColView<T, DIM> view{cols, n, j, zm};
    acc = ef_add(acc, ef_mul(Air::template eval<T>(view, X), eqv(j)));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Ef o;
#pragma unroll
    for (int c = 0; c < 5; c++) o.c[c] = __shfl_down_sync(0xffffffffu, acc.c[c], off);
    acc = ef_add(acc, o);
  }
  if (lane == 0) st_ef(partial + ((uint64_t)blockIdx.x * Air::DEG + zi) * 5, acc);
}

__global__ void airg_sum_partials_kernel(const uint32_t* __restrict__ partial, int n_part, int n_words, uint32_t* __restrict__ out) {
  const int w = threadIdx.x;
  if (w >= n_words) return;
  uint32_t s = 0;
  for (int k = 0; k < n_part; k++) s = kb_add(s, partial[(uint64_t)k * n_words + w]);
  out[w] = s;
}

bool air_table_shape(uint32_t table, uint32_t* n_cols, uint32_t* n_shift, uint32_t* degree, uint32_t* max_constraints) {
  switch (table & 0xffu) {
    case 0: *n_cols = 20, *n_shift = 2, *degree = 5, *max_constraints = 13; return true;
    case 1: *n_cols = ExtOpAir::COLS, *n_shift = ExtOpAir::SHIFT, *degree = ExtOpAir::DEG, *max_constraints = 34; return true;
    case 2: *n_cols = Poseidon16Air::COLS, *n_shift = 0, *degree = Poseidon16Air::DEG, *max_constraints = 100; return true;
  }
  return false;
}

template <class Air>
static cudaError_t launch_round(cudaStream_t stream, const uint32_t* d_cols, uint32_t dim, uint64_t n, uint64_t half,
                                const uint32_t* d_eq_tab, uint32_t k_vars, uint32_t m_vars, const AirExtraBig& X, uint32_t* d_part,
                                uint32_t* d_out) {
  uint64_t blocks = (half + 31) / 32;
  if (blocks > (uint64_t)AIR_MAX_BLOCKS) blocks = AIR_MAX_BLOCKS;
  constexpr int AIRG_THREADS = 32 * Air::DEG;
  if (dim == 1)
    air_stream_round_kernel<Air, Fb, 1><<<(unsigned)blocks, AIRG_THREADS, 0, stream>>>(d_cols, n, half, d_eq_tab, k_vars, m_vars, X, d_part);
  else
    air_stream_round_kernel<Air, Ef, 5><<<(unsigned)blocks, AIRG_THREADS, 0, stream>>>(d_cols, n, half, d_eq_tab, k_vars, m_vars, X, d_part);
  count_launch();
  airg_sum_partials_kernel<<<1, 64, 0, stream>>>(d_part, (int)blocks, Air::DEG * 5, d_out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t air_generic_round(cudaStream_t stream, uint32_t table, const uint32_t* d_cols, uint32_t dim, uint32_t log_n,
                              const uint32_t* d_eq_tab, uint32_t k_vars, const uint32_t* alpha_powers, uint32_t n_alpha,
                              const uint32_t* la, uint32_t n_la, const uint32_t beta[5], uint32_t* d_scratch, uint32_t* d_out) {
  uint32_t nc, ns, deg, maxc;
  if (!air_table_shape(table, &nc, &ns, &deg, &maxc) || (table & 0xffu) == 0) return cudaErrorInvalidValue;
  const int bus = (table & 0x100u) ? 0 : 1;
  if (log_n < 1 || (dim != 1 && dim != 5) || n_alpha < maxc - (bus ? 0 : 1) || (bus && n_la < 5)) return cudaErrorInvalidValue;
  static AirExtraBig X;  // 2.1 KiB: keep it off the stack
  const uint32_t used = maxc < n_alpha ? maxc : n_alpha;
  for (uint32_t k = 0; k < (uint32_t)AIRG_MAX_ALPHA; k++)
    for (int c = 0; c < 5; c++) X.alpha[k].c[c] = k < used ? alpha_powers[5 * k + c] : 0;
  for (int k = 0; k < 4; k++)
    for (int c = 0; c < 5; c++) X.la[k].c[c] = bus ? la[5 * k + c] : 0;
  // la_last * DOMAINSEP with LOGUP_PRECOMPILE_DOMAINSEP = 1 (lean_vm/src/core/constants.rs:5)
  for (int c = 0; c < 5; c++) X.la_last.c[c] = bus ? la[5 * (n_la - 1) + c] : 0, X.beta.c[c] = beta[c];
  X.bus = bus;

  const uint64_t n = (uint64_t)1 << log_n, half = n / 2;
  uint32_t* d_part = d_scratch;
  if ((table & 0xffu) == 1) return launch_round<ExtOpAir>(stream, d_cols, dim, n, half, d_eq_tab, k_vars, log_n - 1, X, d_part, d_out);
  return launch_round<Poseidon16Air>(stream, d_cols, dim, n, half, d_eq_tab, k_vars, log_n - 1, X, d_part, d_out);
}

// ---- fill_trace_poseidon_16 (trace_gen.rs:10-165): one row per thread, columns 25..109 from columns 8..25 ------
__global__ void __launch_bounds__(128) poseidon16_fill_trace_kernel(uint32_t* __restrict__ cols, uint64_t n) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  using A = Poseidon16Air;
  Fb s[16], in[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = in[i] = Fb{cols[(uint64_t)(A::INPUTS + i) * n + row]};
#pragma unroll 1
  for (int r = 0; r < 2; r++) {
    A::two_full_rounds(s, 2 * r);
#pragma unroll
    for (int i = 0; i < 16; i++) cols[(uint64_t)(A::BEGIN + 16 * r + i) * n + row] = s[i].v;
  }
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = add_c(s[i], c_p16.FIRST_RC[i]);
  {
    Fb t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = dot16c(s, c_p16.M_I[i]);
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = t[i];
  }
#pragma unroll 1
  for (int r = 0; r < 20; r++) {
    s[0] = cube(s[0]);
    cols[(uint64_t)(A::PARTIAL + r) * n + row] = s[0].v;
    if (r < 19) s[0] = add_c(s[0], c_p16.SCALAR_RC[r]);
    const Fb old = s[0];
    const Fb dot = dot16c(s, c_p16.FIRST_ROW[r]);
#pragma unroll
    for (int i = 1; i < 16; i++) s[i] = s[i] + mul_c(old, c_p16.V[r][i - 1]);
    s[0] = dot;
  }
  A::two_full_rounds(s, 4);
#pragma unroll
  for (int i = 0; i < 16; i++) cols[(uint64_t)(A::END + i) * n + row] = s[i].v;
  A::two_full_rounds(s, 6);
  const Fb fp = Fb{cols[(uint64_t)A::FLAG_PERMUTE * n + row]};
  const Fb nfp = -sub_one(fp);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const Fb comp = s[i] + in[i];
    cols[(uint64_t)(A::OUT_LEFT + i) * n + row] = (nfp * comp + fp * s[i]).v;
    cols[(uint64_t)(A::OUT_RIGHT + i) * n + row] = (fp * s[i + 8]).v;
  }
}

cudaError_t poseidon16_fill_trace(cudaStream_t stream, uint32_t* d_cols, uint64_t n) {
  if (n == 0) return cudaSuccess;
  poseidon16_fill_trace_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_cols, n);
  count_launch();
  return cudaGetLastError();
}

}  // namespace lm

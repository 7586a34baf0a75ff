// Poseidon1-KoalaBear width-16 permutation / 2-to-1 compression for sm_100a, one state per thread.
//
// Computes exactly what the reference's Poseidon1KoalaBear16::permute / compress_in_place compute
//   crates/backend/koala-bear/src/poseidon1_koalabear_16.rs:873-912 (permute_generic)
//   .../poseidon1_koalabear_16.rs:934-1016 (permute_simd), :1020-1030 (compress_in_place)
// but is organised for a GPU integer pipe instead of 16-lane AVX-512:
//
//  * Full rounds.  State lanes are held as v * R^e (R = 2^32) with a per-round exponent e that is allowed
//    to drift: S-box = two lazy Montgomery products (v^3 R^(3e-2)); the circulant MDS has entries <= 101, so
//    it is 16 IMAD.WIDE per output lane on the un-reduced S-box outputs (sum < 2^41) started from the next
//    round's constant, followed by ONE Montgomery reduction (2 instructions) that leaves v' R^(3e-3) in
//    [0, p + 2^9).  Every step is homogeneous, so pre-scaling the round constants by the right power of R
//    (tools/gen_poseidon1_consts.py) makes the drift free; it is undone once at the end (FIX).
//  * Partial rounds.  Only lane 0 is non-linear, so lanes 1..15 are never updated round by round.  With
//    z_k = (lane0 at round k)^3 the whole partial section is   s0_{r+1} = FR0[r] z_r + D_r + sum_{k<r} GTRI[r][k] z_k,
//    D = G x' (one 21x16 matrix-vector product on the state entering the section) and the lanes leaving the
//    section are  MI x' + V z + const.  All of it is long dot products with delayed reduction (KbDot):
//    ~1.1k IMAD.WIDE and ~60 reductions instead of 300 reduced rank-1 updates.
//
// Inputs/outputs are canonical Montgomery-form residues in [0, p).
#pragma once
#include <utility>
#include "kb.cuh"

namespace lm {

struct P1Tables {
  uint32_t RC0[16];
  uint32_t RC_INIT[4][16];
  uint32_t G[21][16];
  uint32_t G_CONST[21];
  uint32_t FR0[20];
  uint32_t GTRI[20][20];
  uint32_t MI[15][16];
  uint32_t V[15][20];
  uint32_t LANE_CONST[15];
  uint32_t RC_TERM[3][16];
  uint32_t FIX;
  double RC_INIT_D[4][16];  // RC_INIT / RC_TERM again as doubles (exact integers < p)
  double RC_TERM_D[3][16];
};

// v^3 * R^-2 for a lane a < 1.43 p; result < 1.75 p
LM_HD uint32_t p1_sbox_lazy(uint32_t a) { return kb_mul_lazy(kb_mul_lazy(a, a), a); }

// out[i] = redc(init[i] + 4 * sum_j C[(i - j) mod 16] * a3[j]),  C = first column of the circulant MDS.
// N_OUT < 16 computes only the first N_OUT lanes (the digest half of a compression).  The factor 4 is the
// un-normalised even/odd splitting below; it is part of the scale drift the generated constants account for.
//
// Device version: the sum is accumulated EXACTLY on the FP64 pipe.  Measured on B200
// (profiles/r01_int_pipes2.txt): IMAD.WIDE issues once per 4 cycles per SM sub-partition (6 with a 64-bit
// addend), DFMA once per ~2.  u32 -> f64 is the 2^52 trick (one DADD), f64 -> (lo, hi) another DADD; every
// intermediate is an integer of magnitude below 2^53, so the result is bit-identical to the integer evaluation
// of the host path.  The circulant product y(X) = c(X) x(X) mod X^16 - 1 is split by the CRT
//   X^16 - 1 = (X^8 - 1)(X^8 + 1),  X^8 - 1 = (X^4 - 1)(X^4 + 1):
// u = x_lo + x_hi, v = x_lo - x_hi; 2 y_lo = P + N, 2 y_hi = P - N with P = cu (*) u cyclic, N = cv (*) v negacyclic,
// once more on P, and the halvings dropped: 96 DFMA + 48 DADD instead of 256 DFMA.
template <int N_OUT>
LM_HD void p1_mds_redc(const uint32_t a3[16], const uint32_t* init, const double* init_d, uint32_t out[16]) {
#ifdef __CUDA_ARCH__
  (void)init;
  constexpr double TWO52 = 4503599627370496.0;
  // c = {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1}
  // level 1: cu[k] = c[k] + c[k+8], cv[k] = 2 (c[k] - c[k+8])  (doubled: P below carries a factor 2)
  constexpr double CV[8] = {-200, 4, 22, 10, 112, 2, -72, 124};
  // level 2 on cu = {102, 4, 15, 39, 78, 3, 66, 64}: cuu[k] = cu[k] + cu[k+4], cuv[k] = cu[k] - cu[k+4]
  constexpr double CUU[4] = {180, 7, 81, 103};
  constexpr double CUV[4] = {24, 1, -51, -25};
  double x[16];
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = __hiloint2double(0x43300000, (int)a3[j]) - TWO52;
  double u[8], v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) u[k] = x[k] + x[k + 8], v[k] = x[k] - x[k + 8];
  // N = cv (*) v negacyclic of length 8
  double n[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int k = (i - j) & 7;
      acc = fma(j <= i ? CV[k] : -CV[k], v[j], acc);
    }
    n[i] = acc;
  }
  // 2 P = (cuu (*) uu cyclic4) +- (cuv (*) uv negacyclic4)
  double uu[4], uv[4];
#pragma unroll
  for (int k = 0; k < 4; k++) uu[k] = u[k] + u[k + 4], uv[k] = u[k] - u[k + 4];
  double pp[4], pn[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k = (i - j) & 3;
      a = fma(CUU[k], uu[j], a);
      b = fma(j <= i ? CUV[k] : -CUV[k], uv[j], b);
    }
    pp[i] = a;
    pn[i] = b;
  }
  double p2[8];
#pragma unroll
  for (int i = 0; i < 4; i++) p2[i] = pp[i] + pn[i], p2[i + 4] = pp[i] - pn[i];
#pragma unroll
  for (int i = 0; i < N_OUT; i++) {
    double y = i < 8 ? p2[i] + n[i] : p2[i - 8] - n[i - 8];  // 4 * (C x)_i, a non-negative integer < 2^43
    if (init_d) y += init_d[i];
    y += TWO52;
    const uint32_t lo = (uint32_t)__double2loint(y);
    const uint32_t hi = (uint32_t)__double2hiint(y) - 0x43300000u;
    out[i] = kb_redc_lazy(((uint64_t)hi << 32) | lo);
  }
#else
  (void)init_d;
  constexpr uint32_t C[16] = {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1};
#pragma unroll
  for (int i = 0; i < N_OUT; i++) {
    uint64_t acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) acc += (uint64_t)C[(16 + i - j) & 15] * a3[j];
    out[i] = kb_redc_lazy(4 * acc + (init ? (uint64_t)init[i] : 0ull));
  }
#endif
}

// dot(x[0..16), row) on top of `init`
#ifdef LM_P1_ACC96
LM_HD KbAcc96 p1_dot16(const uint32_t x[16], const uint32_t* row, uint64_t init) {
  KbAcc96 d(init);
#pragma unroll
  for (int j = 0; j < 16; j++) d.mac(x[j], row[j]);
  return d;
}
LM_HD KbDot p1_dot16_folded(const uint32_t x[16], const uint32_t* row, uint64_t init) {
#else
LM_HD KbDot p1_dot16(const uint32_t x[16], const uint32_t* row, uint64_t init) {
#endif
  KbDot d(init);
  d.mac<0>(x[0], row[0]);
  d.mac<1>(x[1], row[1]);
  d.mac<2>(x[2], row[2]);
  d.mac<3>(x[3], row[3]);
  d.mac<4>(x[4], row[4]);
  d.mac<5>(x[5], row[5]);
  d.mac<6>(x[6], row[6]);
  d.mac<7>(x[7], row[7]);
  d.mac<8>(x[8], row[8]);
  d.mac<9>(x[9], row[9]);
  d.mac<10>(x[10], row[10]);
  d.mac<11>(x[11], row[11]);
  d.mac<12>(x[12], row[12]);
  d.mac<13>(x[13], row[13]);
  d.mac<14>(x[14], row[14]);
  d.mac<15>(x[15], row[15]);
  return d;
}

// ---- partial section with the matrix entries as IMMEDIATE operands (device) ------------------------------------------
// Read from constant memory, the ~1.1 k matrix entries of the partial section cost ~390 LDCU (mostly 128-bit) and ~500
// UMOV per permutation on top of the IMAD.WIDE that consume them — 12 % of the issue slots of a kernel that is limited
// by issue slots and the multiplier pipe jointly.  As compile-time constants (template indices into a constexpr copy of
// the generated tables) they become the 32-bit immediate of the IMAD.WIDE itself.  -DLM_P1_NO_IMMEDIATES restores the
// constant-memory form.
LM_HD constexpr P1Tables p1_tables_constexpr() {
  return P1Tables
#include "poseidon1_tables.inc"
      ;
}
#if defined(__CUDA_ARCH__) && !defined(LM_P1_NO_IMMEDIATES)
#define LM_P1_IMM 1
template <int N>
using p1_seq = std::make_integer_sequence<int, N>;

// WHICH = 0: row R of G on top of G_CONST[R] (row 0: on top of 0); WHICH = 1: row R of MI on top of LANE_CONST[R]
template <int WHICH, int R, int... J>
__device__ __forceinline__ uint32_t p1_dot16_imm(const uint32_t x[16], std::integer_sequence<int, J...>) {
  constexpr P1Tables t = p1_tables_constexpr();
  constexpr uint32_t init = WHICH == 1 ? t.LANE_CONST[WHICH == 1 ? R : 0] : (R == 0 ? 0u : t.G_CONST[WHICH == 0 ? R : 0]);
#ifdef LM_P1_ACC96
  KbAcc96 d(init);
  (([&] {
     constexpr uint32_t c = WHICH == 1 ? t.MI[WHICH == 1 ? R : 0][J] : t.G[WHICH == 0 ? R : 0][J];
     d.mac(x[J], c);
   }()),
   ...);
#else
  KbDot d(init);
  (([&] {
     constexpr uint32_t c = WHICH == 1 ? t.MI[WHICH == 1 ? R : 0][J] : t.G[WHICH == 0 ? R : 0][J];
     d.template mac<J>(x[J], c);
   }()),
   ...);
#endif
  return d.finish_lazy();
}
template <int... R>
__device__ __forceinline__ void p1_all_d_imm(const uint32_t x[16], uint32_t d[20], std::integer_sequence<int, R...>) {
  ((d[R] = p1_dot16_imm<0, R + 1>(x, p1_seq<16>{})), ...);
}
template <int... I>
__device__ __forceinline__ void p1_all_lane_lin_imm(const uint32_t x[16], uint32_t lane_lin[15], std::integer_sequence<int, I...>) {
  ((lane_lin[I] = p1_dot16_imm<1, I>(x, p1_seq<16>{})), ...);
}
// s0_{R+1} = D_R + FR0[R] z_R + sum_{K<R} GTRI[R][K] z_K
template <int R, int... K>
__device__ __forceinline__ uint32_t p1_tri_imm(uint32_t d_r, const uint32_t z[20], std::integer_sequence<int, K...>) {
  constexpr P1Tables t = p1_tables_constexpr();
#ifdef LM_P1_ACC96
  KbAcc96 acc(mul_wide(d_r, KB_R1));
  {
    constexpr uint32_t c = t.FR0[R];
    acc.mac(z[R], c);
  }
  (([&] {
     constexpr uint32_t c = t.GTRI[R][K];
     acc.mac(z[K], c);
   }()),
   ...);
  return acc.finish_lazy();
#else
  uint64_t acc = mul_wide(d_r, KB_R1);
  {
    constexpr uint32_t c = t.FR0[R];
    acc = mad_wide(z[R], c, acc);
  }
  (([&] {
     if ((1 + K) % 4 == 0) acc = kb_fold(acc);
     constexpr uint32_t c = t.GTRI[R][K];
     acc = mad_wide(z[K], c, acc);
   }()),
   ...);
  return kb_redc_lazy(kb_fold(acc));
#endif
}
template <bool SYNC, int... R>
__device__ __forceinline__ void p1_partial_rounds_imm(uint32_t& s0, const uint32_t d[20], uint32_t z[20],
                                                      std::integer_sequence<int, R...>) {
  (([&] {
     z[R] = kb_canon(p1_sbox_lazy(s0));
     s0 = p1_tri_imm<R>(d[R], z, p1_seq<R>{});
     if (SYNC && R % 4 == 3) __syncthreads();
   }()),
   ...);
}
// lane I + 1 leaving the section: lane_lin[I] + sum_K V[I][K] z_K
template <int I, int... K>
__device__ __forceinline__ uint32_t p1_lane_imm(uint32_t lin, const uint32_t z[20], std::integer_sequence<int, K...>) {
  constexpr P1Tables t = p1_tables_constexpr();
#ifdef LM_P1_ACC96
  KbAcc96 acc(mul_wide(lin, KB_R1));
  (([&] {
     constexpr uint32_t c = t.V[I][K];
     acc.mac(z[K], c);
   }()),
   ...);
  return acc.finish_lazy();
#else
  uint64_t acc = mul_wide(lin, KB_R1);
  (([&] {
     if (K > 0 && K % 4 == 0) acc = kb_fold(acc);
     constexpr uint32_t c = t.V[I][K];
     acc = mad_wide(z[K], c, acc);
   }()),
   ...);
  return kb_redc_lazy(kb_fold(acc));
#endif
}
template <bool SYNC, int... I>
__device__ __forceinline__ void p1_all_lanes_imm(const uint32_t lane_lin[15], const uint32_t z[20], uint32_t a[16],
                                                 std::integer_sequence<int, I...>) {
  (([&] {
     a[I + 1] = p1_lane_imm<I>(lane_lin[I], z, p1_seq<20>{});
     if (SYNC && I % 5 == 4) __syncthreads();
   }()),
   ...);
}
#endif

// Permutation; N_OUT = 16 for the full permutation, 8 when only the digest half is needed.
// s: canonical in, canonical out (lanes >= N_OUT are left unspecified).
// SYNC: the device code places a CTA-wide barrier after every full round and a few times inside the partial
// section.  The warps of a CTA then walk the ~100 KiB instruction stream together and share its fetches; with two
// 256-thread CTAs per SM (<= 128 registers) this measured 8 % faster than three free-running 128-thread CTAs
// (profiles/r01_leaf_barrier_sweep.txt).  Only for kernels in which every thread of the CTA runs the permutation
// the same number of times.
#if defined(__CUDA_ARCH__)
#define LM_P1_BARRIER() do { if (SYNC) __syncthreads(); } while (0)
#else
#define LM_P1_BARRIER() do { } while (0)
#endif
// Partial section (20 partial rounds and the linear maps around them) on one state held by one thread:
// x = state entering the section (first_rc already added, lanes < p + 2^9), a = state leaving it (first terminal
// round constant included).
template <class Tab, bool SYNC>
LM_HD void p1_partial_section(const uint32_t x[16], uint32_t a[16], const Tab& T) {
#ifdef LM_P1_IMM
  {
    uint32_t d[20], lane_lin[15], z[20];
    p1_all_d_imm(x, d, p1_seq<20>{});
    uint32_t s0 = p1_dot16_imm<0, 0>(x, p1_seq<16>{});
    LM_P1_BARRIER();
    p1_all_lane_lin_imm(x, lane_lin, p1_seq<15>{});
    LM_P1_BARRIER();
    p1_partial_rounds_imm<SYNC>(s0, d, z, p1_seq<20>{});
    a[0] = s0;
    p1_all_lanes_imm<SYNC>(lane_lin, z, a, p1_seq<15>{});
  }
#else
  uint32_t d[20];  // D_r, canonical-ish (< p + 2^25), held at R^1
#pragma unroll
  for (int r = 0; r < 20; r++) d[r] = p1_dot16(x, T.G[r + 1], T.G_CONST[r + 1]).finish_lazy();
  uint32_t s0 = p1_dot16(x, T.G[0], 0).finish_lazy();
  LM_P1_BARRIER();

  // lanes 1..15 leaving the section: start their accumulators with MI x' + const now, then x is dead
  uint32_t lane_lin[15];
#pragma unroll
  for (int i = 0; i < 15; i++) lane_lin[i] = p1_dot16(x, T.MI[i], T.LANE_CONST[i]).finish_lazy();
  LM_P1_BARRIER();

  uint32_t z[20];
#pragma unroll
  for (int r = 0; r < 20; r++) {
    z[r] = kb_canon(p1_sbox_lazy(s0));
    // s0_{r+1} = D_r + FR0[r] z_r + sum_{k<r} GTRI[r][k] z_k ;  D_r enters as D_r * 2^32 == D_r * R
#ifdef LM_P1_ACC96
    KbAcc96 acc(mul_wide(d[r], KB_R1));
    acc.mac(z[r], T.FR0[r]);
#pragma unroll
    for (int k = 0; k < r; k++) acc.mac(z[k], T.GTRI[r][k]);
    s0 = acc.finish_lazy();
#else
    uint64_t acc = mul_wide(d[r], KB_R1);
    acc = mad_wide(z[r], T.FR0[r], acc);
    int terms = 1;  // products on top of the (small) D_r term
#pragma unroll
    for (int k = 0; k < r; k++) {
      if (terms % 4 == 0) acc = kb_fold(acc);
      acc = mad_wide(z[k], T.GTRI[r][k], acc);
      terms++;
    }
    s0 = kb_redc_lazy(kb_fold(acc));
#endif
    if (r % 4 == 3) LM_P1_BARRIER();
  }

  a[0] = s0;
#pragma unroll
  for (int i = 0; i < 15; i++) {
#ifdef LM_P1_ACC96
    KbAcc96 acc(mul_wide(lane_lin[i], KB_R1));
#pragma unroll
    for (int k = 0; k < 20; k++) acc.mac(z[k], T.V[i][k]);
    a[i + 1] = acc.finish_lazy();
#else
    uint64_t acc = mul_wide(lane_lin[i], KB_R1);
    int terms = 0;  // canonical z (< p) times constants (< p): four products per fold
#pragma unroll
    for (int k = 0; k < 20; k++) {
      if (terms > 0 && terms % 4 == 0) acc = kb_fold(acc);
      acc = mad_wide(z[k], T.V[i][k], acc);
      terms++;
    }
    a[i + 1] = kb_redc_lazy(kb_fold(acc));
#endif
    if (i % 5 == 4) LM_P1_BARRIER();
  }

#endif

}

template <int N_OUT, class Tab, bool SYNC = false>
LM_HD void p1_permute(uint32_t s[16], const Tab& T) {
  uint32_t a[16], x[16];

  // ---- 4 initial full rounds.  Kept as a real loop on the device (constants indexed by the round): the fully
  // unrolled permutation is ~140 KiB of SASS and was starved by instruction fetch (ncu: 19 % of warp cycles in
  // stall_no_instruction, I-cache hit rate 68 %).
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = kb_add(s[i], T.RC0[i]);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = p1_sbox_lazy(a[i]);
    p1_mds_redc<16>(a, T.RC_INIT[r], T.RC_INIT_D[r], x);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = x[i];
    LM_P1_BARRIER();
  }
  // x = x' (state entering the partial section, first_rc already added), held at R^-39, lanes < p + 2^9

  p1_partial_section<Tab, SYNC>(x, a, T);

  // ---- 4 terminal full rounds (first round constant already inside a[]); three looped, the last one only
  // produces the lanes that are kept
#ifdef __CUDA_ARCH__
#pragma unroll 1
#else
#pragma unroll
#endif
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = p1_sbox_lazy(a[i]);
    p1_mds_redc<16>(a, T.RC_TERM[r], T.RC_TERM_D[r], x);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = x[i];
    LM_P1_BARRIER();
  }
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = p1_sbox_lazy(a[i]);
  p1_mds_redc<N_OUT>(a, nullptr, nullptr, x);
#pragma unroll
  for (int i = 0; i < N_OUT; i++) s[i] = kb_mul(x[i], T.FIX);
}

// compress_in_place: state <- permute(state) + state, only the first N_OUT lanes are produced.
template <int N_OUT, class Tab, bool SYNC = false>
LM_HD void p1_compress(uint32_t s[16], const Tab& T) {
  uint32_t in[N_OUT];
#pragma unroll
  for (int i = 0; i < N_OUT; i++) in[i] = s[i];
  p1_permute<N_OUT, Tab, SYNC>(s, T);
#pragma unroll
  for (int i = 0; i < N_OUT; i++) s[i] = kb_add(s[i], in[i]);
}

}  // namespace lm

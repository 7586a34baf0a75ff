// Device-resident Fiat-Shamir challenger: the Poseidon1 duplex sponge of the reference's transcript, run by ONE warp
// inside the sumcheck kernels so that a round does not cost a host round trip.
//
// Device restatement of
//   crates/backend/fiat-shamir/src/challenger.rs:8-76    Challenger: observe / duplex / sample_many (rate = state[8..16])
//   crates/backend/fiat-shamir/src/prover.rs:74-128      add_*_scalars, add_sumcheck_polynomial (absorb the full
//                                                        polynomial, send coefficients 1..), sample
//   crates/backend/fiat-shamir/src/utils.rs:30-41        expand_bare_to_full
//   crates/backend/koala-bear/src/quintic_extension/extension.rs:585-613  inverse through the Frobenius conjugates
// The host mirror (spine.cu `lm_fs`) hands its sponge state to `DevFs` before a device-driven phase and takes it back,
// together with the transcript words the device appended, afterwards: both sides stay one transcript.
//
// Permutation: textbook form (add round constants, S-box, circulant MDS) with the 16 lanes of the state spread over
// the lanes of a warp (both half-warps carry the state); an MDS row is split between the two half-warps — 8 shuffles + 8
// multiply-accumulates each, one 64-bit exchange to join the halves.  Latency, not throughput, is what matters here (one
// sponge per proof).
#pragma once
#include <cstdint>
#include "kb.cuh"

namespace lm {

struct DevFs {
  uint32_t state[16];
  uint32_t rate_fresh;
  uint32_t n_words;    // transcript words appended by the device in this phase
  uint32_t cap_words;  // capacity of the transcript buffer
  uint32_t error;      // DEVFS_ERR_* bits, sticky
};
enum { DEVFS_ERR_STALE = 1, DEVFS_ERR_OVERFLOW = 2, DEVFS_ERR_ZERO_INV = 4 };

constexpr uint32_t devfs_to_monty(uint64_t canonical) { return (uint32_t)(((canonical % KB_P) << 32) % KB_P); }

struct DevFsTables {
  uint32_t rc[28][16];  // Montgomery form
  uint32_t mds[16];     // Montgomery form of the first MDS column
  uint32_t frob[4][5];  // X^(p i), i = 1..4, Montgomery form (quintic_extension/mod.rs:19-48)
};
constexpr DevFsTables devfs_make_tables() {
  constexpr uint32_t rc_canon[28 * 16] = {
#include "poseidon1_rc.inc"
  };
  constexpr uint32_t mds_col[16] = {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1};
  constexpr uint32_t frob[4][5] = {
      {1576402667, 1173144480, 1567662457, 1206866823, 2428146},
      {1680345488, 1381986, 615237464, 1380104858, 295431824},
      {441230756, 323126830, 704986542, 1445620072, 503505220},
      {1364444097, 1144738982, 2008416047, 143367062, 1027410849},
  };
  DevFsTables t{};
  for (int r = 0; r < 28; r++)
    for (int i = 0; i < 16; i++) t.rc[r][i] = devfs_to_monty(rc_canon[r * 16 + i]);
  for (int i = 0; i < 16; i++) t.mds[i] = devfs_to_monty(mds_col[i]);
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 5; k++) t.frob[i][k] = devfs_to_monty(frob[i][k]);
  return t;
}
#ifdef __CUDACC__
static __device__ const DevFsTables d_fs_tables = devfs_make_tables();
static __constant__ uint32_t c_fs_mds[16] = {
    devfs_to_monty(1),  devfs_to_monty(3),  devfs_to_monty(13),  devfs_to_monty(22), devfs_to_monty(67), devfs_to_monty(2),
    devfs_to_monty(15), devfs_to_monty(63), devfs_to_monty(101), devfs_to_monty(1),  devfs_to_monty(2),  devfs_to_monty(17),
    devfs_to_monty(11), devfs_to_monty(1),  devfs_to_monty(51),  devfs_to_monty(1)};

// ---- small EF helpers on uniform (warp-replicated) values -------------------------------------------------
__device__ __forceinline__ Ef fs_ef_one() { return Ef{{KB_R1, 0, 0, 0, 0}}; }
__device__ __forceinline__ uint32_t fs_kb_inv(uint32_t a) {
  // a^(p - 2), p - 2 = 0x7effffff
  uint32_t r = KB_R1;
  uint32_t e = KB_P - 2;
#pragma unroll 1
  for (int i = 0; i < 31; i++) {
    if (e & 1) r = kb_mul(r, a);
    a = kb_mul(a, a);
    e >>= 1;
  }
  return r;
}
__device__ __forceinline__ Ef fs_ef_mul(const Ef& a, const Ef& b) { return ef_mul(a, b); }
__device__ __forceinline__ Ef fs_ef_frobenius(const Ef& a) {
  Ef out = {{a.c[0], 0, 0, 0, 0}};
#pragma unroll
  for (int i = 1; i < 5; i++)
#pragma unroll
    for (int k = 0; k < 5; k++) out.c[k] = kb_add(out.c[k], kb_mul(a.c[i], d_fs_tables.frob[i - 1][k]));
  return out;
}
// a^-1 = (a^p a^(p^2) a^(p^3) a^(p^4)) / Norm(a); returns false (and zero) when a = 0
static __device__ __noinline__ bool fs_ef_inv(const Ef& a, Ef* out) {
  const Ef f1 = fs_ef_frobenius(a);
  const Ef f12 = fs_ef_frobenius(fs_ef_mul(a, f1));
  const Ef conj = fs_ef_mul(f12, fs_ef_frobenius(fs_ef_frobenius(f12)));
  const Ef norm = fs_ef_mul(a, conj);
  if (norm.c[0] == 0) {
    *out = ef_zero();
    return false;
  }
  *out = ef_mul_base(conj, fs_kb_inv(norm.c[0]));
  return true;
}
// eq(alpha, r) = (1 - alpha)(1 - r) + alpha r
__device__ __forceinline__ Ef fs_eq1(const Ef& alpha, const Ef& r) {
  const Ef one = fs_ef_one();
  return ef_add(fs_ef_mul(ef_sub(one, alpha), ef_sub(one, r)), fs_ef_mul(alpha, r));
}

// round constants -> shared memory, by the warp that will run the sponge (once per kernel)
constexpr int DEVFS_RC_WORDS = 28 * 16;
__device__ __forceinline__ void fs_load_rc(uint32_t* rc_shared) {
  const uint32_t* src = &d_fs_tables.rc[0][0];
  for (int t = threadIdx.x & 31; t < DEVFS_RC_WORDS; t += 32) rc_shared[t] = src[t];
  __syncwarp();
}

// The permutation on one warp: lane l holds state[l & 15] (both half-warps carry the same values); rc_s = shared-memory copy of
// the round constants (fs_load_rc), mds_h = this half-warp's eight MDS coefficients mds[8 h + k], h = lane >> 4.
// One warp runs straight-line code at the speed of its instruction fetches, so the round loop is NOT unrolled (the body
// stays in the instruction cache) and the round constants come from shared memory, fetched one round ahead: measured 36 us
// per transcript step with the constants loaded from global memory inside the loop, 30 us fully unrolled (45 KiB of code
// fetched per permutation), a few us this way.
__device__ __forceinline__ void fs_load_mds_half(uint32_t (&mds_h)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) mds_h[k] = c_fs_mds[8 * ((threadIdx.x >> 4) & 1) + k];
}
static __device__ __noinline__ uint32_t fs_warp_permute(uint32_t v, const uint32_t* rc_s, const uint32_t (&mds_h)[8]) {
  const int i = threadIdx.x & 15, kh = (threadIdx.x >> 1) & 8;  // kh = 8 h
  uint32_t rc = rc_s[i];
#pragma unroll 1
  for (int r = 0; r < 28; r++) {
    v = kb_add(v, rc);
    rc = rc_s[((r + 1) % 28) * 16 + i];
    const bool full = r < 4 || r >= 24;
    if (full || i == 0) v = kb_mul(kb_mul(v, v), v);
    // y_i = sum_k mds[k] x_{(i - k) mod 16}: half-warp h sums k = 8 h .. 8 h + 7 (canonical x, constants < p: four products
    // per 64-bit accumulator), the halves meet through one 64-bit exchange
    uint64_t a0 = 0, a1 = 0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      a0 = mad_wide(__shfl_sync(0xffffffffu, v, (i - kh - k) & 15), mds_h[k], a0);
      a1 = mad_wide(__shfl_sync(0xffffffffu, v, (i - kh - k - 1) & 15), mds_h[k + 1], a1);
    }
    const uint64_t mine = kb_fold(a0) + kb_fold(a1);
    v = kb_canon(kb_redc_lazy(mine + __shfl_xor_sync(0xffffffffu, mine, 16)));
  }
  return v;
}

// ---- the sponge, one warp; lane l holds state[l & 15] (both half-warps carry the same values) -----------------
struct FsWarp {
  uint32_t x;
  bool fresh;
  DevFs* fs;
  uint32_t* tr;  // transcript buffer (device), nullptr = do not record
  uint32_t n_words;
  const uint32_t* rc_s;  // shared-memory copy of the round constants (DEVFS_RC_WORDS words, fs_load_rc)
  uint32_t mds_h[8];     // this half-warp's eight MDS coefficients: mds[8 h + k], h = lane >> 4

  __device__ __forceinline__ void load(DevFs* f, uint32_t* transcript, const uint32_t* rc_shared) {
    fs = f;
    tr = transcript;
    rc_s = rc_shared;
    x = f->state[threadIdx.x & 15];
    fresh = f->rate_fresh != 0;
    n_words = f->n_words;
    fs_load_mds_half(mds_h);
  }
  __device__ __forceinline__ void store() {
    __syncwarp();
    if ((threadIdx.x & 31) < 16) fs->state[threadIdx.x & 15] = x;
    if ((threadIdx.x & 31) == 0) {
      fs->rate_fresh = fresh ? 1u : 0u;
      fs->n_words = n_words;
    }
  }
  __device__ __forceinline__ void flag(uint32_t bits) {
    if ((threadIdx.x & 31) == 0) atomicOr(&fs->error, bits);
  }

  __device__ __noinline__ void permute() { x = fs_warp_permute(x, rc_s, mds_h); }
  // state[8..16] = chunk, permute (challenger.rs observe); `w` is this lane's chunk word for lanes with (l & 15) >= 8
  __device__ __forceinline__ void observe8(uint32_t w) {
    if ((threadIdx.x & 15) >= 8) x = w;
    permute();
    fresh = true;
  }
  __device__ __forceinline__ void duplex() { observe8(0); }
  // absorb n words (uniform pointer, shared or global memory), zero padded to a multiple of 8
  __device__ __forceinline__ void absorb(const uint32_t* words, int n) {
    const int i = threadIdx.x & 15;
    for (int off = 0; off < n; off += 8) {
      const int idx = off + i - 8;
      observe8((i >= 8 && idx < n) ? words[idx] : 0u);
    }
  }
  __device__ __forceinline__ void record(const uint32_t* words, int n) {
    if (!tr) return;
    if (n_words + (uint32_t)n > fs->cap_words) {
      flag(DEVFS_ERR_OVERFLOW);
      return;
    }
    for (int t = threadIdx.x & 31; t < n; t += 32) tr[n_words + t] = words[t];
    n_words += (uint32_t)n;
  }
  // sample_vec(1): 5 words of the rate; a stale rate is the caller's protocol error (prover.rs asserts it)
  __device__ __forceinline__ Ef sample_ef() {
    if (!fresh) flag(DEVFS_ERR_STALE);
    fresh = false;
    Ef e;
#pragma unroll
    for (int k = 0; k < 5; k++) e.c[k] = __shfl_sync(0xffffffffu, x, 8 + k);
    return e;
  }
  // add_sumcheck_polynomial(coeffs, Some(eq_alpha)) for a bare polynomial of degree n_bare - 1 held in `buf`
  // (shared memory, 5 (n_bare + 1) words of scratch behind it): absorb ((1 - a) + (2a - 1) X) bare(X), send bare[1..]
  __device__ __forceinline__ void add_sumcheck_polynomial_bare(uint32_t* buf, int n_bare, const Ef& eq_alpha) {
    const Ef one = fs_ef_one();
    const Ef c0 = ef_sub(one, eq_alpha), c1 = ef_sub(ef_add(eq_alpha, eq_alpha), one);
    uint32_t* full = buf + 5 * n_bare;
    Ef prev = ef_zero();
    for (int i = 0; i <= n_bare; i++) {
      Ef cur = ef_zero(), acc = ef_zero();
      if (i < n_bare) {
#pragma unroll
        for (int k = 0; k < 5; k++) cur.c[k] = buf[5 * i + k];
        acc = fs_ef_mul(c0, cur);
      }
      if (i > 0) acc = ef_add(acc, fs_ef_mul(c1, prev));
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) full[5 * i + k] = acc.c[k];
      }
      prev = cur;
    }
    __syncwarp();
    absorb(full, 5 * (n_bare + 1));
    record(buf + 5, 5 * (n_bare - 1));
  }
};
#endif  // __CUDACC__

}  // namespace lm

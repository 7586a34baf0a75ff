// Poseidon1-KoalaBear width-16 for sm_100a with every "constant matrix x state" product on the 5th-generation tensor cores
// (tcgen05.mma.kind::i8, accumulators in tensor memory), one state per thread, 128 states per MMA.  DESIGN.md 2.1b.
//
// Same function as poseidon1.cuh (poseidon1_koalabear_16.rs:873-912 permute_generic / :1020-1030 compress_in_place of the
// reference), bit for bit on the canonical outputs.  What moves off the multiplier pipe (IMAD.WIDE 4 cycles, IMAD 2, DFMA ~2.2
// per warp instruction and sub-partition; the one-state-per-thread kernels keep it 79 % busy):
//   * the circulant MDS of the 8 full rounds, 16 x 16 (was 96 DFMA + 48 DADD per state and round) AND the second Montgomery
//     reduction of every full-round S-box (the MDS takes the unreduced 64-bit product, its constant is 4 C R^-1)
//   * D = G x' and the first lane entering the partial section, 21 x 16                  (was 336 IMAD.WIDE + folds)
//   * lanes 1..15 leaving it, MI x' + V z + const, 15 x 37                               (was 540 IMAD.WIDE + folds)
//   * the strictly lower triangle GTRI z of the partial rounds across blocks of 8 rounds (128 of its 190 IMAD.WIDE)
// What stays: the S-boxes, the serial chain of the 20 partial rounds with the triangle terms of its own block, and one
// Montgomery reduction per produced lane.
//
// Why tcgen05 and not mma.sync: on B200 the legacy warp-level IMMA.16832.U8.U8 does NOT overlap with the multiplier pipe
// (tools/microbench/imma_mix.cu, profiles/r02_p1_umma.txt: 16 IMMA + 16 IMAD.WIDE take 248 cycles, 128 + 82 apart); the
// mma.sync version of this file (poseidon1_mma.cuh) is bit-exact and 18 % SLOWER than the scalar kernel.  tcgen05.mma is
// asynchronous: one thread issues it for 128 states, the product runs in the tensor unit while the other warps of the SM keep
// the integer pipes busy.
//
// Mapping.  M = 128 states = the 4 warps of a "group" (thread i of the group = row i = TMEM lane i, so tcgen05.ld.32x32b hands
// every thread the accumulators of ITS state: the one-state-per-thread layout of poseidon1.cuh is kept, no transposition).
// A (shared memory, K-major, no swizzle: 8 rows x 16 bytes core matrices, LBO = 128, SBO = 1280): a row of ten 16-byte chunks
// written by its thread with ordinary stores — the bytes of a word are its u8 limbs.  Full rounds: chunks 0..7 = the sixteen
// 64-bit S-box products (k = 8 e + i).  Partial section: chunks 0..3 = the 16 words of x' (k = 4 e + i), chunks 4..8 = the 20
// S-box outputs z, byte 144 = the constant 1.
// B (shared memory, N x K, K-major; image built on the host by p1u_build_b_image): for a 31-bit constant M, limb i of the input
// is multiplied by M_i = M 2^(8 i) mod p, whose four bytes j go to columns 4 o + j:
//         T_j = sum_(e,i) limb_i(x_e) byte_j(M_i[o][e]) < 2^24,   sum_e M[o][e] x_e  ==  T0 + 2^8 T1 + 2^16 T2 + 2^24 T3  (mod p),
// four accumulator columns and ONE reduction per output lane instead of 16-37 multiply-accumulates with folds.
// The recombination is shifts and adds on the ALU pipe (T0 + 2^8 T1 + const < 2^32 by construction), the reduction is
// kb_redc_lazy.  Intermediate lanes are congruent to those of poseidon1.cuh (not always the same lazy representative); the
// bounds the S-boxes need (< 1.43 p) hold with room: every reduction here sees less than 2^48, i.e. returns less than p + 2^16.
//
// Cost per permutation and group: 12 round trips (store row -> fence.proxy.async -> 128-thread barrier -> MMA issue by one
// elected lane -> commit -> mbarrier -> tcgen05.ld), ~450 cycles each when nothing else runs
// (tools/microbench/umma_i8_probe.cu); four 128-thread CTAs per SM interleave, so a group's round trip is covered by the
// S-boxes of the other three.
#pragma once
#include "poseidon1.cuh"
#include "umma.cuh"

namespace lm {

// image of the three B matrices in their shared-memory (canonical K-major) layout
constexpr int P1U_B_MDS = 0;              // 64 x 128
constexpr int P1U_B_G = 8192;             // 96 x 64 (84 live columns)
constexpr int P1U_B_MV = 8192 + 6144;     // 64 x 160 (60 live columns, 145 live k)
constexpr int P1U_B_T0 = 8192 + 6144 + 10240;  // 48 x 32: z_0..7 into D_8..19
constexpr int P1U_B_T1 = P1U_B_T0 + 1536;      // 16 x 32: z_8..15 into D_16..19
constexpr int P1U_B_BYTES = P1U_B_T1 + 512;
constexpr int P1U_A_BYTES = 128 * 160;    // one group's A rows
constexpr int P1U_TMEM_COLS_PER_GROUP = 96;   // the widest product (G) has 96 columns
LM_HD constexpr uint32_t p1u_tmem_alloc_cols(int groups) {  // allocations are powers of two >= 32
  return groups * 96 <= 128 ? 128u : groups * 96 <= 256 ? 256u : 512u;
}

// host: fill `img` (P1U_B_BYTES) from the generated tables
inline void p1u_build_b_image(const P1Tables& T, uint8_t* img) {
  for (int i = 0; i < P1U_B_BYTES; i++) img[i] = 0;
  auto at = [&](int base, int kchunks, int n, int k) -> uint8_t& {
    return img[base + (n / 8) * (kchunks * 128) + (k / 16) * 128 + (n % 8) * 16 + (k % 16)];
  };
  // limb i of input word `kword` (words of LIMBS bytes) times the constant m: byte j of m 2^(8 i) mod p goes to column 4 o + j
  auto put = [&](int base, int kchunks, int o, int kword, uint32_t m, int limbs = 4) {
    uint64_t mi = m % KB_P;
    for (int i = 0; i < limbs; i++) {
      for (int j = 0; j < 4; j++) at(base, kchunks, 4 * o + j, limbs * kword + i) = (uint8_t)(mi >> (8 * j));
      mi = (mi << 8) % KB_P;
    }
  };
  // MDS: the inputs are the UNREDUCED 64-bit S-box products w = a^2 R^-1 * a (8 limbs), the constant is 4 C R^-1 mod p
  uint64_t rinv = 1, base = (1ull << 32) % KB_P;
  for (uint32_t e = KB_P - 2; e; e >>= 1, base = base * base % KB_P)
    if (e & 1) rinv = rinv * base % KB_P;
  const uint32_t C[16] = {1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1};
  for (int o = 0; o < 16; o++)
    for (int e = 0; e < 16; e++) put(P1U_B_MDS, 8, o, e, (uint32_t)(4 * C[(o - e) & 15] * rinv % KB_P), 8);
  // G: output r < 20 is D_r (row r + 1 of G), output 20 the lane entering round 0 (row 0)
  for (int r = 0; r < 21; r++)
    for (int e = 0; e < 16; e++) put(P1U_B_G, 4, r, e, T.G[r < 20 ? r + 1 : 0][e]);
  // the strictly lower triangle of the partial rounds, by blocks of 8 rounds: once z_0..7 (z_8..15) are known their
  // contribution to every LATER block's D_r is one more product accumulated into the same columns
  for (int r = 8; r < 20; r++)
    for (int k = 0; k < 8; k++) put(P1U_B_T0, 2, r - 8, k, T.GTRI[r][k]);
  for (int r = 16; r < 20; r++)
    for (int k = 8; k < 16; k++) put(P1U_B_T1, 2, r - 16, k - 8, T.GTRI[r][k]);
  for (int o = 0; o < 15; o++) {
    for (int e = 0; e < 16; e++) put(P1U_B_MV, 10, o, e, T.MI[o][e]);
    for (int q = 0; q < 20; q++) put(P1U_B_MV, 10, o, 16 + q, T.V[o][q]);
    // the additive constant rides on a byte of the A rows that is always 1 (k = 144, chunk 9)
    for (int j = 0; j < 4; j++) at(P1U_B_MV, 10, 4 * o + j, 144) = (uint8_t)(T.LANE_CONST[o] >> (8 * j));
  }
}

// ---- CPU model of the formulation -------------------------------------------------------------------------------------------
// The permutation exactly as the kernels below run it — same B image, same row layout, same recombination and block structure —
// with every tcgen05.mma replaced by the integer dot products it stands for.  Host test infrastructure for the CPU tier
// (lm_host_poseidon1_umma_model): pins the image builder, the pre-shifted constants and the no-carry bounds against the oracle
// without a GPU.
struct P1uModel {
  const uint8_t* img;
  uint8_t row[160];
  uint32_t col[128];
  void store_words(int chunk, const uint32_t* w, int n) {
    for (int i = 0; i < n; i++)
      for (int b = 0; b < 4; b++) row[16 * chunk + 4 * i + b] = (uint8_t)(w[i] >> (8 * b));
  }
  void product(int n, int k_steps, int b_off, int b_kchunks, int a_chunk0 = 0, int d_col0 = 0, bool acc = false) {
    for (int j = 0; j < n; j++) {
      uint32_t sum = acc ? col[d_col0 + j] : 0u;
      for (int k = 0; k < 32 * k_steps; k++)
        sum += (uint32_t)row[16 * a_chunk0 + k] * img[b_off + (j / 8) * (b_kchunks * 128) + (k / 16) * 128 + (j % 8) * 16 + (k % 16)];
      col[d_col0 + j] = sum;
    }
  }
};
inline uint32_t p1u_model_redc(uint64_t t) {  // kb_redc_lazy in the form the device uses
  const uint32_t m = (uint32_t)t * 0x81000001u;
  return (uint32_t)(t >> 32) - (uint32_t)(((uint64_t)m * KB_P) >> 32) + KB_P;
}
inline void p1u_model_full_round(P1uModel& M, const uint32_t a[16], const uint32_t* rc, uint32_t out[16]) {
  for (int k = 0; k < 16; k++) {
    const uint64_t w = (uint64_t)p1u_model_redc((uint64_t)a[k] * a[k]) * a[k];
    const uint32_t words[2] = {(uint32_t)w, (uint32_t)(w >> 32)};
    for (int i = 0; i < 2; i++)
      for (int b = 0; b < 4; b++) M.row[8 * k + 4 * i + b] = (uint8_t)(words[i] >> (8 * b));
  }
  M.product(64, 4, P1U_B_MDS, 8);
  for (int i = 0; i < 16; i++)
    out[i] = p1u_combine_redc<0>(M.col[4 * i], M.col[4 * i + 1], M.col[4 * i + 2], M.col[4 * i + 3], rc ? rc[i] : 0u);
}
inline void p1u_model_permute(const P1Tables& T, const uint8_t* img, uint32_t s[16]) {
  P1uModel M;
  M.img = img;
  for (int i = 0; i < 160; i++) M.row[i] = 0;
  M.row[144] = 1;
  uint32_t a[16], x[16];
  for (int i = 0; i < 16; i++) a[i] = kb_add(s[i], T.RC0[i]);
  for (int r = 0; r < 4; r++) {
    p1u_model_full_round(M, a, T.RC_INIT[r], x);
    for (int i = 0; i < 16; i++) a[i] = x[i];
  }
  // partial section
  M.store_words(0, x, 16);
  M.product(96, 2, P1U_B_G, 4);
  uint32_t z[20];
  uint32_t s0 = p1u_combine_redc<0>(M.col[80], M.col[81], M.col[82], M.col[83], 0u);
  const int block_begin[4] = {0, 8, 16, 20};
  for (int b = 0; b < 3; b++) {
    const int k0 = block_begin[b];
    for (int r = k0; r < block_begin[b + 1]; r++) {
      z[r] = kb_canon(p1u_model_redc((uint64_t)p1u_model_redc((uint64_t)s0 * s0) * s0));
      uint64_t acc = p1u_combine64(M.col[4 * r], M.col[4 * r + 1], M.col[4 * r + 2], M.col[4 * r + 3], T.G_CONST[1 + r]);
      acc += (uint64_t)z[r] * T.FR0[r];
      for (int i = 0; k0 + i < r; i++) {
        if ((1 + i) % 4 == 0) acc = kb_fold(acc);
        acc += (uint64_t)z[k0 + i] * T.GTRI[r][k0 + i];
      }
      s0 = p1u_model_redc(kb_fold(acc));
    }
    M.store_words(4 + k0 / 4, z + k0, block_begin[b + 1] - k0);
    if (b == 0) M.product(48, 1, P1U_B_T0, 2, 4, 32, true);
    if (b == 1) M.product(16, 1, P1U_B_T1, 2, 6, 64, true);
  }
  a[0] = s0;
  M.product(64, 5, P1U_B_MV, 10);
  for (int o = 0; o < 15; o++) a[o + 1] = p1u_combine_redc<0>(M.col[4 * o], M.col[4 * o + 1], M.col[4 * o + 2], M.col[4 * o + 3], 0u);
  for (int r = 0; r < 4; r++) {
    p1u_model_full_round(M, a, r < 3 ? T.RC_TERM[r] : nullptr, x);
    for (int i = 0; i < 16; i++) a[i] = x[i];
  }
  for (int i = 0; i < 16; i++) s[i] = kb_mul(x[i], T.FIX);
}

#ifdef __CUDACC__

// Per-thread view of the group's tensor-core plumbing.  Kernels: dynamic shared memory of p1u_smem_bytes(groups), blockDim.x =
// 128 * groups, every thread of the CTA calls p1u_setup once, the permutation the same number of times, p1u_teardown once.
struct P1uCtx {
  uint32_t a_row;    // shared address of this thread's row, chunk 0
  uint32_t a_base;   // shared address of the group's A rows
  uint32_t b_base;   // shared address of the B image
  uint32_t bar;      // the group's mbarrier
  uint32_t tmem;     // TMEM address of this warp's lanes, first column of the group
  uint32_t tmem_d;   // TMEM address of the group's accumulator (lane 0, first column of the group)
  uint32_t tmem_alloc;
  uint32_t parity;
  uint32_t group;
  bool issuer_warp;  // warp 0 of the group (warp-uniform): one elected lane of it issues the MMAs
};
LM_HD constexpr int p1u_smem_bytes(int groups) { return P1U_B_BYTES + groups * P1U_A_BYTES + 64; }

__device__ __forceinline__ void p1u_store_chunk(const P1uCtx& c, int chunk, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(c.a_row + chunk * 128), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
}

__device__ __forceinline__ P1uCtx p1u_setup(uint8_t* smem /* 1024-byte aligned */, const uint8_t* __restrict__ b_image, int groups) {
  const int tid = threadIdx.x, warp = tid >> 5, g = tid >> 7, r = tid & 127;
  uint8_t* sb = smem;
  uint8_t* sa = smem + P1U_B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P1U_B_BYTES + groups * P1U_A_BYTES);
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(bars + 6);
  for (int i = tid; i < P1U_B_BYTES / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sb)[i] = __ldg(reinterpret_cast<const uint4*>(b_image) + i);
  const uint32_t cols = p1u_tmem_alloc_cols(groups);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p1u_smem_u32(tm_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < groups) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p1u_smem_u32(bars + tid)), "r"(1) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  P1uCtx c;
  c.a_base = p1u_smem_u32(sa + g * P1U_A_BYTES);
  c.a_row = c.a_base + (r >> 3) * 1280 + (r & 7) * 16;
  c.b_base = p1u_smem_u32(sb);
  c.bar = p1u_smem_u32(bars + g);
  c.tmem_alloc = *tm_slot;
  c.tmem_d = c.tmem_alloc + g * P1U_TMEM_COLS_PER_GROUP;
  c.tmem = c.tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
  c.parity = 0;
  c.group = g;
  c.issuer_warp = __shfl_sync(0xffffffffu, r >> 5, 0) == 0;
  p1u_store_chunk(c, 9, 1u, 0u, 0u, 0u);  // k = 144 is the constant 1 (LANE_CONST column of the MI | V product), the rest padding
  return c;
}
__device__ __forceinline__ void p1u_teardown(const P1uCtx& c, int groups) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem_alloc), "r"(p1u_tmem_alloc_cols(groups)) : "memory");
}

// rows are written: run K_STEPS MMAs of N columns against the B matrix at b_off and wait for the accumulators
// A_CHUNK0: first 16-byte chunk of the rows that takes part; D_COL0: first accumulator column; ACC: add to what the columns hold
template <int N, int K_STEPS, int B_OFF, int B_KCHUNKS, int A_CHUNK0 = 0, int D_COL0 = 0, bool ACC = false>
__device__ __forceinline__ void p1u_product(P1uCtx& c) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (blockDim.x == 128) {
    asm volatile("bar.sync 1, 128;" ::: "memory");
  } else {
    switch (c.group) {  // immediate barrier ids: a register operand makes ptxas reserve all 16
      case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
      case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
      case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
      case 3: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
      default: asm volatile("bar.sync 5, 128;" ::: "memory"); break;
    }
  }
  if (c.issuer_warp) {  // warp-uniform branch, then one elected lane: the other warps skip the issue code altogether
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p1u_elect_one()) {
      const uint64_t da = p1u_desc(c.a_base + A_CHUNK0 * 128, 1280), db = p1u_desc(c.b_base + B_OFF, B_KCHUNKS * 128);
#pragma unroll
      for (int k = 0; k < K_STEPS; k++)
        p1u_mma(c.tmem_d + D_COL0, da + (uint64_t)(k * 256 >> 4), db + (uint64_t)(k * 256 >> 4), p1u_idesc(N), ACC || k > 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c.bar) : "memory");
    }
    __syncwarp();
  }
  for (uint32_t spins = 0; !p1u_try_wait(c.bar, c.parity);)
    if (++spins > (1u << 26)) __trap();  // a lost MMA completion must not hang the device
  c.parity ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// S-boxes and MDS of one full round: out[i] == (rc[i] + 4 sum_j C[(i - j) mod 16] a[j]^3 R^-2 R^-1) R^-1, i < N_OUT — what
// p1_sbox_lazy + p1_mds_redc compute, with the second reduction of every S-box folded into the matrix: the product
// w = (a^2 R^-1) a is stored UNREDUCED (8 limbs per lane, K = 128) and B carries 4 C R^-1 2^(8 i) mod p for limb i.
template <int N_OUT>
__device__ __forceinline__ void p1u_sbox_mds_redc(P1uCtx& c, const uint32_t a[16], const uint32_t* rc, uint32_t out[16]) {
#pragma unroll
  for (int k = 0; k < 16; k++) {
    // 64-bit stores: the product is an aligned register pair as it leaves the multiplier (128-bit stores cost ~24 register
    // moves per round to line up quads; the 2-way bank conflict of the 8-byte pattern is invisible next to that)
    const uint64_t w = mul_wide(kb_mul_lazy(a[k], a[k]), a[k]);
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(c.a_row + (k >> 1) * 128 + (k & 1) * 8), "l"(w) : "memory");
  }
  p1u_product<4 * N_OUT, 4, P1U_B_MDS, 8>(c);
#pragma unroll
  for (int h = 0; h < N_OUT / 8; h++) {
    uint32_t v[32];
    p1u_ld32(c.tmem + 32 * h, v);
#pragma unroll
    for (int i = 0; i < 8; i++)
      out[8 * h + i] = p1u_combine_redc<0>(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], rc ? rc[8 * h + i] : 0u);
  }
}

#ifdef LM_P1_IMM
// one partial round: z_R = s0^3, s0 <- (y + FR0[R] z_R + sum_{K0 <= k < R} GTRI[R][k] z_k) / 2^32, where y (< 2^48) already
// holds D_R R and the triangle terms of the earlier blocks.  At most 4 products (< 0.2462 * 2^64 each) sit on a folded accumulator.
template <int R, int K0, int... I>
__device__ __forceinline__ void p1u_partial_round(uint32_t& s0, uint64_t y, uint32_t z[20], std::integer_sequence<int, I...>) {
  constexpr P1Tables t = p1_tables_constexpr();
  z[R] = kb_canon(p1_sbox_lazy(s0));
  uint64_t acc = y;
  {
    constexpr uint32_t c = t.FR0[R];
    acc = mad_wide(z[R], c, acc);
  }
  (([&] {
     if ((1 + I) % 4 == 0) acc = kb_fold(acc);
     constexpr uint32_t c = t.GTRI[R][K0 + I];
     acc = mad_wide(z[K0 + I], c, acc);
   }()),
   ...);
  s0 = kb_redc_lazy(kb_fold(acc));
}
template <int K0, int... J>
__device__ __forceinline__ void p1u_partial_block(uint32_t& s0, const uint64_t* y, uint32_t z[20], std::integer_sequence<int, J...>) {
  ((p1u_partial_round<K0 + J, K0>(s0, y[J], z, p1_seq<J>{})), ...);
}
#endif

// The partial section of p1_partial_section: x = state entering (lanes < p + 2^15), a = state leaving.
// D = G x' stays in tensor memory as four columns per round; the rounds run in blocks of 8, 8 and 4: a block reads its columns
// (64-bit values, no reduction), runs its rounds with the triangle terms of its OWN block as scalar multiply-accumulates, stores
// its z and lets one more product add GTRI z to the columns of the later blocks: 62 + 20 scalar multiply-accumulates instead of 210.
template <bool SYNC>
__device__ __forceinline__ void p1u_partial_section(P1uCtx& c, const uint32_t x[16], uint32_t a[16], const P1Tables& t) {
#ifdef LM_P1_IMM  // (the host pass of nvcc parses this body too)
#pragma unroll
  for (int k = 0; k < 4; k++) p1u_store_chunk(c, k, x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
  p1u_product<96, 2, P1U_B_G, 4>(c);
  uint32_t z[20];
  uint32_t s0;
  uint64_t y[8];
  {
    uint32_t v[4];
    p1u_ld4(c.tmem + 80, v);
    s0 = p1u_combine_redc<0>(v[0], v[1], v[2], v[3], 0u);
  }
  {
    uint32_t v[32];
    p1u_ld32(c.tmem, v);
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = p1u_combine64(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], t.G_CONST[1 + i]);
  }
  p1u_partial_block<0>(s0, y, z, p1_seq<8>{});
  p1u_store_chunk(c, 4, z[0], z[1], z[2], z[3]);
  p1u_store_chunk(c, 5, z[4], z[5], z[6], z[7]);
  p1u_product<48, 1, P1U_B_T0, 2, 4, 32, true>(c);
  {
    uint32_t v[32];
    p1u_ld32(c.tmem + 32, v);
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = p1u_combine64(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], t.G_CONST[9 + i]);
  }
  p1u_partial_block<8>(s0, y, z, p1_seq<8>{});
  p1u_store_chunk(c, 6, z[8], z[9], z[10], z[11]);
  p1u_store_chunk(c, 7, z[12], z[13], z[14], z[15]);
  p1u_product<16, 1, P1U_B_T1, 2, 6, 64, true>(c);
  {
    uint32_t v[32];  // columns 64..79 are D_16..19 (80..83: the first lane, already used; the rest is not written)
    p1u_ld32(c.tmem + 64, v);
#pragma unroll
    for (int i = 0; i < 4; i++) y[i] = p1u_combine64(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], t.G_CONST[17 + i]);
  }
  p1u_partial_block<16>(s0, y, z, p1_seq<4>{});
  a[0] = s0;
  p1u_store_chunk(c, 8, z[16], z[17], z[18], z[19]);
  p1u_product<64, 5, P1U_B_MV, 10>(c);
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t v[32];
    p1u_ld32(c.tmem + 32 * h, v);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int o = 8 * h + i;
      if (o < 15) a[o + 1] = p1u_combine_redc<0>(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3], 0u);
    }
  }
#endif
}

// Permutation; N_OUT = 16 or 8 (digest half).  s: canonical in, canonical out.
template <int N_OUT, bool SYNC>
__device__ __forceinline__ void p1u_permute(P1uCtx& c, uint32_t s[16], const P1Tables& T) {
  uint32_t a[16], x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = kb_add(s[i], T.RC0[i]);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    p1u_sbox_mds_redc<16>(c, a, T.RC_INIT[r], x);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = x[i];
  }
  p1u_partial_section<SYNC>(c, x, a, T);
#pragma unroll 1
  for (int r = 0; r < 3; r++) {
    p1u_sbox_mds_redc<16>(c, a, T.RC_TERM[r], x);
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = x[i];
  }
  p1u_sbox_mds_redc<N_OUT>(c, a, nullptr, x);
#pragma unroll
  for (int i = 0; i < N_OUT; i++) s[i] = kb_mul(x[i], T.FIX);
}

template <int N_OUT, bool SYNC>
__device__ __forceinline__ void p1u_compress(P1uCtx& c, uint32_t s[16], const P1Tables& T) {
  uint32_t in[N_OUT];
#pragma unroll
  for (int i = 0; i < N_OUT; i++) in[i] = s[i];
  p1u_permute<N_OUT, SYNC>(c, s, T);
#pragma unroll
  for (int i = 0; i < N_OUT; i++) s[i] = kb_add(s[i], in[i]);
}

#endif  // __CUDACC__

}  // namespace lm

// Poseidon1-KoalaBear width-16 for sm_100a, 32 states per WARP, the circulant MDS of the full rounds on the tensor cores.
//
// Same function as poseidon1.cuh (poseidon1_koalabear_16.rs:873-912 permute_generic / :1020-1030 compress_in_place of the
// reference), bit for bit: the exact integer  y = 4 (C x)_i + rc_i  of p1_mds_redc is produced by integer MMA instead of the
// FP64 pipe, everything before and after it (lazy S-boxes, one Montgomery reduction per lane, pre-scaled constants, the
// partial section) is the code of poseidon1.cuh.
//
// Why.  The one-state-per-thread kernels are bound by the multiplier pipe (IMAD.WIDE 4 cycles, DFMA ~2.2 cycles per warp
// instruction and sub-partition; ncu: 79 % busy) and the 8 MDS layers are ~2.5 k of its ~11 k cycles per warp-permutation.
// A "constant matrix x batch of states" product is what mma.sync.m16n8k32.u8.u8.s32 (SASS IMMA.16832.U8.U8, 8.0 cycles per
// sub-partition on B200, profiles/r02_imma_gonogo.txt) does on a pipe these kernels leave idle.
//
// Fragment layout (no packing, no shuffles).  lane = 4 g + t.  A warp holds 32 states; register f[m][q] of lane (g, t) is
// element 4 t + q of state g + 8 m (m, q = 0..3).  With the states on the M dimension, the four BYTES of an A register of
// m16n8k32 are consecutive k: they are the four u8 limbs of ONE state word, so the S-box outputs are A fragments as they
// stand:  a0 = f[2 mt][2 s], a1 = f[2 mt + 1][2 s], a2 = f[2 mt][2 s + 1], a3 = f[2 mt + 1][2 s + 1]  (m-tile mt, k-step s),
// k = 16 j + 4 t + i  <->  (element 4 t + 2 s + j, limb i).
// B (k x 8 columns) of n-tile (q, h): column 2 t' + b  <->  (output element 4 t' + q, limb 2 h + b), entry
// C[(out - in) mod 16] where the limbs agree, else 0 — the limb sums of an output stay separate columns:
//   S_l = sum_e C[(o - e) mod 16] limb_l(x_e)  <  371 * 255,
// and the C fragment hands lane (g, t) exactly S_{2h}, S_{2h+1} of output element 4 t + q of states g + 8 (2 mt), g + 8 (2 mt + 1):
// outputs arrive in the layout the inputs were in.  Each B register of a lane has ONE non-zero byte, one of 7 values
// C[(4 (g >> 1) - 4 t + d) mod 16], d = -3..3, shifted: 7 registers hold every B fragment of the layer.
// Per warp and layer: 32 IMMA (2 m-tiles x 8 n-tiles x 2 k-steps), then per output  y = 4 (S0 + 2^8 S1 + 2^16 S2 + 2^24 S3) + rc
// with shifts and adds on the ALU pipe and the same kb_redc_lazy as before.
//
// The partial section keeps one state per lane (its 20 S-boxes are a serial chain per state): the warp transposes through
// shared memory (lane (g, t) takes state g + 8 t, i.e. a transposition inside each quad), runs p1_partial_section of
// poseidon1.cuh unchanged, and transposes back.
#pragma once
#include "poseidon1.cuh"

namespace lm {

#ifdef __CUDACC__

// shared memory a CTA of W warps needs: round constants + one transposition buffer per warp
template <int WARPS>
struct P1wSmem {
  uint32_t rc[9][16];         // RC0, RC_INIT[0..3], RC_TERM[0..2], zeros (last round adds nothing)
  uint32_t xp[WARPS][32 * 16];  // 32 states x 16 words, 16-byte blocks XOR-swizzled by (state >> 3)
};

template <int WARPS>
__device__ __forceinline__ void p1w_smem_init(P1wSmem<WARPS>& sm, const P1Tables& T) {
  for (int i = threadIdx.x; i < 9 * 16; i += blockDim.x) {
    const int r = i >> 4, e = i & 15;
    sm.rc[r][e] = r == 0 ? T.RC0[e] : r <= 4 ? T.RC_INIT[r - 1][e] : r <= 7 ? T.RC_TERM[r - 5][e] : 0u;
  }
  __syncthreads();
}

// the 7 distinct B values of this lane, b = g & 1 already applied:  cb[d + 3] = C[(4 (g >> 1) - 4 t + d) & 15] << (8 (g & 1))
struct P1wB {
  uint32_t cb[7];
};
__device__ __forceinline__ P1wB p1w_b_init() {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  P1wB B;
#pragma unroll
  for (int d = -3; d <= 3; d++) B.cb[d + 3] = c_kb.mds[(4 * (g >> 1) - 4 * t + d) & 15] << (8 * (g & 1));
  return B;
}

__device__ __forceinline__ void p1w_imma0(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
__device__ __forceinline__ void p1w_imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// out[m][q] = redc(rc[q] + 4 * sum_e C[(o - e) mod 16] a[m][e]),  o = 4 t + q: p1_mds_redc<16> on the warp's 32 states
__device__ __forceinline__ void p1w_mds_redc(const uint32_t (&a)[4][4], const P1wB& B, const uint4 rc, uint32_t (&out)[4][4]) {
  const uint32_t rcq[4] = {rc.x, rc.y, rc.z, rc.w};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int acc[2][2][4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      // k-step s, half j: d = q - 2 s - j
      const uint32_t b00 = B.cb[q + 3] << (16 * h), b01 = B.cb[q + 2] << (16 * h);
      const uint32_t b10 = B.cb[q + 1] << (16 * h), b11 = B.cb[q] << (16 * h);
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        p1w_imma0(acc[mt][h], a[2 * mt][0], a[2 * mt + 1][0], a[2 * mt][1], a[2 * mt + 1][1], b00, b01);
        p1w_imma(acc[mt][h], a[2 * mt][2], a[2 * mt + 1][2], a[2 * mt][3], a[2 * mt + 1][3], b10, b11);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const int mt = m >> 1, half = m & 1;
      const uint32_t x = (uint32_t)acc[mt][0][2 * half] + ((uint32_t)acc[mt][0][2 * half + 1] << 8);
      const uint32_t y = (uint32_t)acc[mt][1][2 * half] + ((uint32_t)acc[mt][1][2 * half + 1] << 8);
      const uint64_t v = ((uint64_t)x << 2) + ((uint64_t)y << 18) + rcq[q];
      out[m][q] = kb_redc_lazy(v);
    }
  }
}

// fragment layout -> one state per lane (lane (g, t) receives state g + 8 t) and back, through the warp's buffer
__device__ __forceinline__ void p1w_to_lanes(const uint32_t (&f)[4][4], uint32_t x[16], uint32_t* xp) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 4; m++)
    *reinterpret_cast<uint4*>(xp + 16 * (g + 8 * m) + 4 * (t ^ m)) = make_uint4(f[m][0], f[m][1], f[m][2], f[m][3]);
  __syncwarp();
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const uint4 v = *reinterpret_cast<const uint4*>(xp + 16 * (g + 8 * t) + 4 * (b ^ t));
    x[4 * b] = v.x, x[4 * b + 1] = v.y, x[4 * b + 2] = v.z, x[4 * b + 3] = v.w;
  }
}
__device__ __forceinline__ void p1w_from_lanes(const uint32_t x[16], uint32_t (&f)[4][4], uint32_t* xp) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int b = 0; b < 4; b++)
    *reinterpret_cast<uint4*>(xp + 16 * (g + 8 * t) + 4 * (b ^ t)) = make_uint4(x[4 * b], x[4 * b + 1], x[4 * b + 2], x[4 * b + 3]);
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const uint4 v = *reinterpret_cast<const uint4*>(xp + 16 * (g + 8 * m) + 4 * (t ^ m));
    f[m][0] = v.x, f[m][1] = v.y, f[m][2] = v.z, f[m][3] = v.w;
  }
}

// Permutation of the warp's 32 states in fragment layout; canonical in, canonical out.  Every lane of the warp must call it.
template <bool SYNC, int WARPS>
__device__ __forceinline__ void p1w_permute(uint32_t (&f)[4][4], const P1wB& B, P1wSmem<WARPS>& sm, const P1Tables& T) {
  const int lane = threadIdx.x & 31, t = lane & 3;
  uint32_t* xp = sm.xp[threadIdx.x >> 5];
  uint32_t a[4][4];
  {
    const uint4 rc = *reinterpret_cast<const uint4*>(&sm.rc[0][4 * t]);
    const uint32_t rcq[4] = {rc.x, rc.y, rc.z, rc.w};
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) a[m][q] = kb_add(f[m][q], rcq[q]);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) a[m][q] = p1_sbox_lazy(a[m][q]);
    p1w_mds_redc(a, B, *reinterpret_cast<const uint4*>(&sm.rc[1 + r][4 * t]), f);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) a[m][q] = f[m][q];
    LM_P1_BARRIER();
  }
  {
    uint32_t x[16], y[16];
    p1w_to_lanes(a, x, xp);
    p1_partial_section<P1Tables, SYNC>(x, y, T);
    p1w_from_lanes(y, a, xp);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) a[m][q] = p1_sbox_lazy(a[m][q]);
    p1w_mds_redc(a, B, *reinterpret_cast<const uint4*>(&sm.rc[5 + r][4 * t]), f);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) a[m][q] = f[m][q];
    LM_P1_BARRIER();
  }
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int q = 0; q < 4; q++) f[m][q] = kb_mul(a[m][q], T.FIX);
}

// compress_in_place on the warp's 32 states: only elements 0..7 (lanes t = 0, 1) of the result are meaningful
template <bool SYNC, int WARPS>
__device__ __forceinline__ void p1w_compress(uint32_t (&f)[4][4], const P1wB& B, P1wSmem<WARPS>& sm, const P1Tables& T) {
  uint32_t in[4][4];
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int q = 0; q < 4; q++) in[m][q] = f[m][q];
  p1w_permute<SYNC, WARPS>(f, B, sm, T);
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int q = 0; q < 4; q++) f[m][q] = kb_add(f[m][q], in[m][q]);
}

#endif  // __CUDACC__

}  // namespace lm

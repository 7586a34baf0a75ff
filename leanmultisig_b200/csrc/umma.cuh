// tcgen05 (5th-generation tensor core) plumbing shared by the kernels that run "constant matrix x batch" products of KoalaBear
// words as u8-limb integer MMAs (poseidon1_umma.cuh, the eq-weights GEMM of sumcheck.cu): shared-memory descriptors for the
// K-major no-swizzle layout, the MMA / commit / mbarrier / tcgen05.ld wrappers, and the recombination of byte columns.
//
// Conventions established by tools/microbench/umma_i8_probe.cu on B200: an operand is rows of K bytes stored as 8-row x 16-byte
// core matrices; LBO (distance of K-adjacent core matrices) = 128 bytes, SBO = bytes per 8-row group; one MMA covers K = 32
// bytes, the next K-step starts 256 bytes further; kind::i8 with both formats 0 = unsigned 8 bit, s32 accumulate.
#pragma once
#include "kb.cuh"

namespace lm {

// Recombination of four accumulator columns and Montgomery reduction in one: returns a value congruent to
//   ((v0 + 2^8 v1 + 2^16 v2 + 2^24 v3) 2^SHIFT + init) / 2^32   in (., . + p],   provided init + (v0 << SHIFT) + (v1 << (8 + SHIFT)) < 2^32 (MDS: v <= 128 * 255^2 < 2^23, init < p: < 2^32 - 2^24;
// G: v < 2^22, init < p; MI | V: v < 2^23.2, init = 0 — its constant is a column of B).
// The 64-bit sum is never formed by the multiplier: its low word and its high word (a shift and a carry) are built on the ALU
// pipe and handed to the two multiplications of the reduction (m = lo p^-1, hi(m p)) — the multiplier pipe is what bounds the
// kernel, and a mad.wide with a 64-bit addend per shift would put 10 of its cycles on every output.
template <int SHIFT>
LM_HD uint32_t p1u_combine_redc(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t init) {
#if defined(__CUDA_ARCH__) && defined(LM_COMBINE_PRMT)
  // Experiment (off): byte shifts as PRMT and three-input adds, forms ptxas cannot turn into IMAD.  ptxas balances the ALU and FMA
  // pipes by instruction count and so puts ~3 IMAD per output on the pipe that bounds the kernel; forcing them onto the ALU pipe
  // costs two more instructions per output and measured SLOWER (6.35 against 6.05 ms, profiles/r02_p1_umma.txt): issue slots.
  static_assert(SHIFT == 0, "the byte-permute form is for byte-aligned columns");
  const uint32_t a = init + v0 + __byte_perm(v1, 0, 0x2104);                                       // < 2^32: no carry
  const uint64_t s = (uint64_t)a + __byte_perm(v2, 0, 0x1044) + __byte_perm(v3, 0, 0x0444);        // + (v2 << 16) + (v3 << 24), mod 2^32 each
  const uint32_t lo = (uint32_t)s;
  const uint32_t hi = (uint32_t)(s >> 32) + __byte_perm(v2, 0, 0x4432) + __byte_perm(v3, 0, 0x4321);  // + (v2 >> 16) + (v3 >> 8)
#elif defined(__CUDA_ARCH__) && !defined(LM_COMBINE_NO_CARRY_ASM)
  const uint32_t a = init + (v0 << SHIFT) + (v1 << (8 + SHIFT));  // < 2^32: no carry
  const uint32_t b = v2 + (v3 << 8);
  uint32_t lo, hi;  // the carry of the low word travels in the carry flag (IADD3 / IADD3.X), not through a compare + select
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b << (16 + SHIFT)), "r"(b >> (16 - SHIFT)));
#else
  const uint32_t a = init + (v0 << SHIFT) + (v1 << (8 + SHIFT));  // < 2^32: no carry
  const uint32_t b = v2 + (v3 << 8);
  const uint32_t lo = a + (b << (16 + SHIFT));
  const uint32_t hi = (b >> (16 - SHIFT)) + (lo < a ? 1u : 0u);
#endif
  const uint32_t m = lo * 0x81000001u;
  const uint64_t u = mul_wide(m, LM_KB_P_OPAQUE);
  return hi - (uint32_t)(u >> 32) + LM_KB_P_OPAQUE;
}

// the 64-bit value v0 + 2^8 v1 + 2^16 v2 + 2^24 v3 + init (same no-carry condition as p1u_combine_redc), built on the ALU pipe
LM_HD uint64_t p1u_combine64(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t init) {
  const uint32_t a = init + v0 + (v1 << 8);
  const uint32_t b = v2 + (v3 << 8);
#ifdef __CUDA_ARCH__
  uint32_t lo, hi;
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b << 16), "r"(b >> 16));
#else
  const uint32_t lo = a + (b << 16);
  const uint32_t hi = (b >> 16) + (lo < a ? 1u : 0u);
#endif
  return ((uint64_t)hi << 32) | lo;
}


#ifdef __CUDACC__

__device__ __forceinline__ uint32_t p1u_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: start address, LBO = 128 (K-adjacent core matrices are contiguous), SBO = bytes per 8-row group
__device__ __forceinline__ uint64_t p1u_desc(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void p1u_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0)
      : "memory");
}
__device__ __forceinline__ bool p1u_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool p1u_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void p1u_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
      "%26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void p1u_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

LM_HD constexpr uint32_t p1u_idesc(uint32_t n) { return (2u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }


// 16 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void p1u_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void p1u_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#endif  // __CUDACC__

}  // namespace lm

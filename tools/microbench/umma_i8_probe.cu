// Probe: one tcgen05.mma.kind::i8 (u8 x u8 -> s32, M = 128, N = 64, K = 64 as two K = 32 steps) with both operands in
// shared memory in the no-swizzle K-major canonical layout (8 rows x 16 bytes core matrices), accumulator in TMEM, read
// back with tcgen05.ld.32x32b — checked against the CPU.  Establishes the descriptor conventions (LBO / SBO) used by
// csrc/poseidon1_umma.cuh and measures the round trip (smem store -> fence -> barrier -> MMA -> commit -> mbarrier -> ld).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench/umma_i8_probe.cu -o tools/microbench/umma_i8_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(128) probe(const uint8_t* A, const uint8_t* B, int32_t* D, uint32_t lbo, uint32_t sbo, int reps,
                                             long long* cycles) {
  __shared__ __align__(128) uint8_t sA[128 * 64];
  __shared__ __align__(128) uint8_t sB[64 * 64];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 64)
    for (int c = 0; c < 4; c++)
      *reinterpret_cast<uint4*>(sB + (tid / 8) * 512 + c * 128 + (tid % 8) * 16) = *reinterpret_cast<const uint4*>(B + tid * 64 + c * 16);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_base;
  const uint32_t idesc = (2u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t v[64];
  uint32_t parity = 0;
  long long t0 = 0;
  for (int rep = 0; rep < reps; rep++) {
    if (rep == 1) t0 = clock64();
    for (int c = 0; c < 4; c++)
      *reinterpret_cast<uint4*>(sA + (tid / 8) * 512 + c * 128 + (tid % 8) * 16) = *reinterpret_cast<const uint4*>(A + tid * 64 + c * 16);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t da = make_desc(smem_u32(sA), lbo, sbo), db = make_desc(smem_u32(sB), lbo, sbo);
      umma_i8(tm, da, db, idesc, 0);
      umma_i8(tm, da + ((2 * lbo) >> 4), db + ((2 * lbo) >> 4), idesc, 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t spins = 0;
    while (!mbar_try_wait(smem_u32(&mbar), parity)) {
      if (++spins > 20000000u) {
        if ((tid & 31) == 0) printf("mbarrier wait timed out (rep %d)\n", rep);
        __trap();
      }
    }
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, "
        "%50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
          "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
          "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
          "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  for (int j = 0; j < 64; j++) D[tid * 64 + j] = (int32_t)v[j];
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64));
}

int main() {
  std::vector<uint8_t> A(128 * 64), B(64 * 64);
  uint32_t x = 12345;
  for (auto& v : A) x = x * 1664525u + 1013904223u, v = x >> 24;
  for (auto& v : B) x = x * 1664525u + 1013904223u, v = x >> 24;
  std::vector<int32_t> ref(128 * 64), got(128 * 64);
  for (int r = 0; r < 128; r++)
    for (int n = 0; n < 64; n++) {
      int32_t s = 0;
      for (int k = 0; k < 64; k++) s += (int32_t)A[r * 64 + k] * B[n * 64 + k];
      ref[r * 64 + n] = s;
    }
  uint8_t *dA, *dB;
  int32_t* dD;
  long long* dC;
  cudaMalloc(&dA, A.size()), cudaMalloc(&dB, B.size()), cudaMalloc(&dD, 4 * got.size()), cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice), cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  const uint32_t conv[2][2] = {{128, 512}, {512, 128}};
  for (int c = 0; c < 2; c++) {
    cudaMemset(dD, 0xff, 4 * got.size());
    probe<<<1, 128>>>(dA, dB, dD, conv[c][0], conv[c][1], 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("LBO %u SBO %u: CUDA error %s\n", conv[c][0], conv[c][1], cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(got.data(), dD, 4 * got.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t i = 0; i < got.size(); i++) bad += got[i] != ref[i];
    printf("LBO %u SBO %u: %d mismatches of %zu  (D[0][0] = %d, expected %d; D[5][9] = %d, expected %d)\n", conv[c][0], conv[c][1], bad,
           got.size(), got[0], ref[0], got[5 * 64 + 9], ref[5 * 64 + 9]);
    if (bad == 0) {
      for (int blocks : {1, 148, 148 * 4}) {
        probe<<<blocks, 128>>>(dA, dB, dD, conv[c][0], conv[c][1], 1001, dC);
        cudaDeviceSynchronize();
        long long cyc;
        cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        printf("  round trip (4 STS.128 + fences + barrier + 2 MMA + commit + mbarrier wait + tcgen05.ld x64), %d CTAs of 128: %.1f cycles\n",
               blocks, cyc / 1000.0);
      }
    }
  }
  return 0;
}

// How fast do 32-byte row pieces reach a PEER GPU's memory over NVLink, by kind of store?  The fused exchange of the row-sharded
// commit (csrc/ntt.cu, scatter pass) writes 8 columns (32 bytes) of a row at a 256-byte row pitch into the matrix of the rank
// that owns the row.  Variants, all writing the same 128 MiB region of a matrix with 64 u32 columns on device 1 from device 0:
//   sm32    every thread stores its 32 bytes with two st.global.v4 (what the scatter pass does today)
//   sm128   every thread stores 128 contiguous bytes (a 32-column tile: the "wider tile" alternative)
//   tma32   one elected thread per CTA issues cp.async.bulk.tensor.2d stores of boxes of 8 columns x 32 rows from shared memory
//   tma256  the same with boxes of 8 columns x 256 rows
//   local   sm32 into device 0's own memory (reference)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/microbench/peer_store_probe.cu -lcuda -o tools/microbench/peer_store_probe
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int W = 64;            // u32 columns of the matrix
constexpr int ROWS = 1 << 19;    // 128 MiB
constexpr int TILE_ROWS = 2048;  // rows per CTA (one 8-column tile of 64 KiB, as in the NTT pass)

__global__ void __launch_bounds__(256) sm32_kernel(uint32_t* dst) {
  // CTA = (row block, column tile); thread t stores rows t, t + 256, ...
  const int ct = blockIdx.x % (W / 8), rb = blockIdx.x / (W / 8);
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (int r = threadIdx.x; r < TILE_ROWS; r += 256) {
    uint4* p = reinterpret_cast<uint4*>(dst + ((size_t)rb * TILE_ROWS + r) * W + 8 * ct);
    p[0] = v;
    p[1] = v;
  }
}
__global__ void __launch_bounds__(256) sm128_kernel(uint32_t* dst) {
  const int ct = blockIdx.x % (W / 32), rb = blockIdx.x / (W / 32);
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (int r = threadIdx.x; r < TILE_ROWS / 4; r += 256) {
    uint4* p = reinterpret_cast<uint4*>(dst + ((size_t)rb * (TILE_ROWS / 4) + r) * W + 32 * ct);
#pragma unroll
    for (int k = 0; k < 8; k++) p[k] = v;
  }
}
// a CTA writes a tile of 64 KiB = (TILE_ROWS * 8 / BOX_COLS) rows x BOX_COLS columns, in boxes of BOX_ROWS rows
template <int BOX_ROWS, int BOX_COLS = 8>
__global__ void __launch_bounds__(256) tma_kernel(const __grid_constant__ CUtensorMap map) {
  extern __shared__ __align__(128) uint8_t sm[];
  constexpr int T_ROWS = TILE_ROWS * 8 / BOX_COLS;
  const int ct = blockIdx.x % (W / BOX_COLS), rb = blockIdx.x / (W / BOX_COLS);
  for (int i = threadIdx.x; i < TILE_ROWS * 8; i += 256) reinterpret_cast<uint32_t*>(sm)[i] = i + blockIdx.x;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int b = 0; b < T_ROWS / BOX_ROWS; b++) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm + (size_t)b * BOX_ROWS * BOX_COLS * 4);
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map), "r"(src), "r"(BOX_COLS * ct),
                   "r"(rb * T_ROWS + b * BOX_ROWS)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

static int make_map(CUtensorMap* m, void* base, int box_rows, int box_cols = 8) {
  cuuint64_t dims[2] = {W, ROWS}, strides[1] = {W * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, el[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return r != CUDA_SUCCESS;
}

template <class F>
static float time_it(F&& launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  return best;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) return printf("needs 2 GPUs\n"), 0;
  uint32_t *d_local, *d_peer;
  const size_t bytes = (size_t)ROWS * W * 4;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&d_peer, bytes));
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&d_local, bytes));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CUtensorMap m32, m256, l32;
  if (make_map(&m32, d_peer, 32) || make_map(&m256, d_peer, 256) || make_map(&l32, d_local, 32)) return 1;
  const int ctas = ROWS / TILE_ROWS * (W / 8);
  CK(cudaFuncSetAttribute(tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_ROWS * 32));
  CK(cudaFuncSetAttribute(tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_ROWS * 32));
  struct { const char* name; float ms; } res[6];
  res[0] = {"local  sm32  ", time_it([&] { sm32_kernel<<<ctas, 256>>>(d_local); })};
  res[1] = {"peer   sm32  ", time_it([&] { sm32_kernel<<<ctas, 256>>>(d_peer); })};
  res[2] = {"peer   sm128 ", time_it([&] { sm128_kernel<<<ctas, 256>>>(d_peer); })};
  res[3] = {"peer   tma32 ", time_it([&] { tma_kernel<32><<<ctas, 256, TILE_ROWS * 32>>>(m32); })};
  res[4] = {"peer   tma256", time_it([&] { tma_kernel<256><<<ctas, 256, TILE_ROWS * 32>>>(m256); })};
  res[5] = {"local  tma32 ", time_it([&] { tma_kernel<32><<<ctas, 256, TILE_ROWS * 32>>>(l32); })};
  CK(cudaDeviceSynchronize());
  for (auto& r : res) printf("%s  %7.3f ms  %7.1f GB/s\n", r.name, r.ms, bytes / r.ms * 1e-6);
  // wider row pieces through TMA (64 and 128 bytes per row), one direction
  CUtensorMap m16, m32c;
  if (make_map(&m16, d_peer, 128, 16) || make_map(&m32c, d_peer, 64, 32)) return 1;
  CK(cudaFuncSetAttribute(tma_kernel<128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_ROWS * 32));
  CK(cudaFuncSetAttribute(tma_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_ROWS * 32));
  const float t16 = time_it([&] { tma_kernel<128, 16><<<ctas, 256, TILE_ROWS * 32>>>(m16); });
  const float t32 = time_it([&] { tma_kernel<64, 32><<<ctas, 256, TILE_ROWS * 32>>>(m32c); });
  printf("peer   tma 16 cols (64 B per row)   %7.3f ms  %7.1f GB/s\n", t16, bytes / t16 * 1e-6);
  printf("peer   tma 32 cols (128 B per row)  %7.3f ms  %7.1f GB/s\n", t32, bytes / t32 * 1e-6);
  // both directions at once (what the exchange of the sharded commit does): device 1 writes into device 0 at the same time
  CK(cudaSetDevice(1));
  CK(cudaDeviceEnablePeerAccess(0, 0));
  CUtensorMap r8, r32;
  if (make_map(&r8, d_local, 32, 8) || make_map(&r32, d_local, 64, 32)) return 1;
  CK(cudaFuncSetAttribute(tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_ROWS * 32));
  CK(cudaFuncSetAttribute(tma_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_ROWS * 32));
  cudaStream_t s1;
  CK(cudaStreamCreate(&s1));
  CK(cudaSetDevice(0));
  auto both = [&](int which) {
    // device 1 runs 8 launches back to back while device 0's single launch is timed
    CK(cudaSetDevice(1));
    for (int k = 0; k < 8; k++) {
      if (which == 0) sm32_kernel<<<ctas, 256, 0, s1>>>(d_local);
      if (which == 1) tma_kernel<32><<<ctas, 256, TILE_ROWS * 32, s1>>>(r8);
      if (which == 2) tma_kernel<64, 32><<<ctas, 256, TILE_ROWS * 32, s1>>>(r32);
    }
    CK(cudaSetDevice(0));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    if (which == 0) sm32_kernel<<<ctas, 256>>>(d_peer);
    if (which == 1) tma_kernel<32><<<ctas, 256, TILE_ROWS * 32>>>(m32);
    if (which == 2) tma_kernel<64, 32><<<ctas, 256, TILE_ROWS * 32>>>(m32c);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaSetDevice(1));
    CK(cudaStreamSynchronize(s1));
    CK(cudaSetDevice(0));
    const char* names[3] = {"sm32", "tma 8 cols", "tma 32 cols"};
    printf("both directions busy, %-12s %7.3f ms  %7.1f GB/s per direction\n", names[which], ms, bytes / ms * 1e-6);
    return 0;
  };
  for (int w = 0; w < 3; w++)
    if (both(w)) return 1;
  return 0;
}

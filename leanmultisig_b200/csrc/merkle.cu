// Poseidon1 Merkle commitment over a row-major KoalaBear matrix (sm_100a).
//
// Device replacement for
//   crates/whir/src/merkle.rs:59-88,215-288      build_merkle_tree_koalabear, first_digest_layer(_with_initial_state)
//   crates/backend/symetric/src/sponge.rs:28-108 precompute_zero_suffix_state, hash_rtl_iter, absorb_rtl_chunks
//   crates/backend/symetric/src/merkle.rs:21-35,50-90  MerkleTree::from_first_layer, compress_layer
//
// Kernels (the wide ones run Poseidon1 with its constant-matrix products on the tensor cores, poseidon1_umma.cuh; LM_P1_SCALAR=1
// selects the one-state-per-thread form of poseidon1.cuh)
//   leaf_sponge_kernel     one thread per matrix row; the row is hashed right to left, rate 8 / width 16, as if zero-extended to
//                          `full_w`; a run of >= 2 all-zero trailing rate chunks is replaced by the pre-computed sponge state of
//                          those zeros (passed by value).  Multiplier-pipe bound: ~3.1 k instructions per compression against 32 B
//                          of row data (6.7 k one state per thread).
//   leaf_absorb_kernel     the same sponge a run of chunks at a time (state kept in the digest buffer): lets the commit hash
//                          columns while later columns are still crossing PCIe
//   tree_level_kernel      one level, one thread per parent (levels with >= 8192 parents)
//   tree_level_warp_kernel, tree_top_warp_kernel   the narrow levels, one WARP per parent (latency, not throughput)
//   tree_levels_kernel     one CTA folds 2*T consecutive digests of a layer through up to log2(2T) levels (LM_TREE_TAIL_THREADS=1)
//   pow_grind_umma_kernel  Fiat-Shamir proof-of-work search, one candidate per thread
//   verify_openings_kernel  verifier side: one warp per opening hashes the leaf (hash_slice) and walks its sibling path to the root
// Every intermediate layer is written (all layers are retained for openings).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include "launch_count.h"
#include "merkle.h"
#include "poseidon1.cuh"
#include "poseidon1_umma.cuh"
#include "devfs.cuh"
#include <mutex>
#include <vector>

namespace lm {

__constant__ P1Tables c_p1 =
#include "poseidon1_tables.inc"
    ;
static const P1Tables h_p1 =
#include "poseidon1_tables.inc"
    ;

// CTA shape of the permutation-bound kernels (profiles/r01_leaf_barrier_sweep.txt): two 256-thread CTAs per SM,
// warps kept together by barriers inside the permutation.
#ifndef LEAF_MIN_BLOCKS
#define LEAF_MIN_BLOCKS 2
#endif
#ifndef LEAF_THREADS
#define LEAF_THREADS 256
#endif
#ifndef LEAF_SYNC
#define LEAF_SYNC 1
#endif
#ifndef LEAF_DEFAULT_GRID
#define LEAF_DEFAULT_GRID 0
#endif

struct State16 {
  uint32_t v[16];
};

// ---- tensor-core formulation (poseidon1_umma.cuh) ---------------------------------------------------------------------
// The wide kernels (leaf sponge, leaf absorb, wide tree levels, explicit states) run the permutation with its linear maps on
// tcgen05; LM_P1_SCALAR=1 selects the one-state-per-thread form of poseidon1.cuh (same outputs, kept for cross-checks).
// CTA shape of the tensor-core kernels: ONE group of 128 threads per CTA, four CTAs per SM (TMEM: 4 x 128 columns; measured
// against 2 x 256 and 1 x 512 threads in profiles/r02_p1_umma.txt — independent groups cover each other's MMA round trips best)
#ifndef UMMA_THREADS
#define UMMA_THREADS 128
#endif
#ifndef UMMA_MIN_BLOCKS
#define UMMA_MIN_BLOCKS 4
#endif
static_assert(UMMA_THREADS % 128 == 0 && UMMA_THREADS <= 512, "the tensor-core permutation works on groups of 128 threads");
constexpr int LEAF_GROUPS = UMMA_THREADS / 128;
constexpr bool UMMA_SYNC = LEAF_SYNC != 0 && UMMA_THREADS > 128;

static bool use_umma() {
  static const bool on = [] {
    const char* e = getenv("LM_P1_SCALAR");
    return !(e && atoi(e) != 0);
  }();
  return on;
}
// the B matrices in shared-memory layout, uploaded once per device
static cudaError_t umma_b_image(const uint8_t** out) {
  static std::mutex mu;
  static uint8_t* per_device[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lock(mu);
  if (!per_device[dev]) {
    static uint8_t img[P1U_B_BYTES];
    p1u_build_b_image(h_p1, img);
    uint8_t* d = nullptr;
    if ((e = cudaMalloc(&d, P1U_B_BYTES)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d, img, P1U_B_BYTES, cudaMemcpyHostToDevice)) != cudaSuccess) {
      cudaFree(d);
      return e;
    }
    per_device[dev] = d;
  }
  *out = per_device[dev];
  return cudaSuccess;
}
// opt a kernel in to its dynamic shared memory once per device
template <class K>
static cudaError_t umma_attr(K kernel, int groups) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p1u_smem_bytes(groups));
}

// Element `pos` of the virtual row: stored value if pos < lim else 0.
// Fast path: whole 8-chunk below lim and 16-byte aligned -> two 128-bit loads.
__device__ __forceinline__ void load_chunk8(const uint32_t* __restrict__ row, int64_t first, uint32_t lim, bool vec_ok,
                                            uint32_t out[8]) {
  if (vec_ok && first >= 0 && (uint64_t)first + 8 <= lim) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(row + first));
    const uint4 hi = __ldg(reinterpret_cast<const uint4*>(row + first) + 1);
    out[0] = lo.x, out[1] = lo.y, out[2] = lo.z, out[3] = lo.w;
    out[4] = hi.x, out[5] = hi.y, out[6] = hi.z, out[7] = hi.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      int64_t pos = first + k;
      out[k] = (pos >= 0 && (uint64_t)pos < lim) ? __ldg(row + pos) : 0u;
    }
  }
}

// Grid-stride over blocks of LEAF_THREADS rows: the launcher may start one CTA per block or a persistent grid (a multiple
// of the SM count) whose CTAs keep the ~100 KiB permutation hot in the instruction cache.
template <bool U>
__global__ void __launch_bounds__(U ? UMMA_THREADS : LEAF_THREADS, U ? UMMA_MIN_BLOCKS : LEAF_MIN_BLOCKS)
leaf_sponge_kernel(const uint32_t* __restrict__ mat, uint64_t h, uint32_t stored_w, uint32_t lim, uint32_t virt_w,
                   int from_state, State16 init, uint32_t* __restrict__ digests, const uint8_t* __restrict__ b_image) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  P1uCtx uc;
  if constexpr (U) uc = p1u_setup(dsm, b_image, LEAF_GROUPS);
  const bool vec_ok = (stored_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(mat) & 15) == 0);
  const int64_t n_chunks = virt_w / 8;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < h; base += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = base + threadIdx.x;
    const bool live = r < h;
    if (!live) r = h - 1;  // every thread runs the sponge (CTA-wide barriers inside the permutation); no store
    const uint32_t* row = mat + r * stored_w;
    uint32_t s[16];
    // One compression call site (the unrolled permutation is ~100 KiB of code): the sponge absorbs rate chunks
    // n-1, n-2, ..., 0 into lanes 8..15; without a precomputed state the first compression also takes chunk n-2
    // as lanes 0..7, so the chunk sequence is n-1, n-3, n-4, ...
    int64_t chunk = n_chunks - 1, n_comp = n_chunks;
    if (from_state) {
#pragma unroll
      for (int i = 0; i < 16; i++) s[i] = init.v[i];
    } else {
      load_chunk8(row, 8 * (n_chunks - 2), lim, vec_ok, s);
      n_comp = n_chunks - 1;
    }
    for (int64_t it = 0; it < n_comp; it++) {
      load_chunk8(row, 8 * chunk, lim, vec_ok, s + 8);
      if constexpr (U)
        p1u_compress<8, UMMA_SYNC>(uc, s, c_p1);
      else
        p1_compress<8, P1Tables, LEAF_SYNC != 0>(s, c_p1);
      chunk -= (it == 0 && !from_state) ? 2 : 1;
    }
    if (live) {
      uint4* out = reinterpret_cast<uint4*>(digests + 8 * r);
      out[0] = make_uint4(s[0], s[1], s[2], s[3]);
      out[1] = make_uint4(s[4], s[5], s[6], s[7]);
    }
  }
  if constexpr (U) p1u_teardown(uc, LEAF_GROUPS);
}

// `count` sponge steps for every row: state (lanes 0..7, kept in the digest buffer between launches) absorbs rate chunks
// chunk_hi, chunk_hi - 1, ..., chunk_hi - count + 1 of the row.  Lets the commit hash columns as soon as they are
// transformed, right to left, while the host-to-device copy of the columns further left is still in flight.
template <bool U>
__global__ void __launch_bounds__(U ? UMMA_THREADS : LEAF_THREADS, U ? UMMA_MIN_BLOCKS : LEAF_MIN_BLOCKS)
leaf_absorb_kernel(const uint32_t* __restrict__ mat, uint64_t h, uint32_t stored_w, uint32_t chunk_hi, uint32_t count, int first,
                   State16 init, uint32_t* __restrict__ digests, const uint8_t* __restrict__ b_image) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  P1uCtx uc;
  if constexpr (U) uc = p1u_setup(dsm, b_image, LEAF_GROUPS);
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < h; base += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = base + threadIdx.x;
    const bool live = r < h;
    if (!live) r = h - 1;
    uint32_t s[16];
    uint4* dg = reinterpret_cast<uint4*>(digests + 8 * r);
    if (first) {
#pragma unroll
      for (int i = 0; i < 8; i++) s[i] = init.v[i];
    } else {
      const uint4 a = dg[0], b = dg[1];
      s[0] = a.x, s[1] = a.y, s[2] = a.z, s[3] = a.w, s[4] = b.x, s[5] = b.y, s[6] = b.z, s[7] = b.w;
    }
    const uint4* src = reinterpret_cast<const uint4*>(mat + r * stored_w + 8 * chunk_hi);
    for (uint32_t k = 0; k < count; k++, src -= 2) {
      const uint4 lo = __ldg(src), hi = __ldg(src + 1);
      s[8] = lo.x, s[9] = lo.y, s[10] = lo.z, s[11] = lo.w, s[12] = hi.x, s[13] = hi.y, s[14] = hi.z, s[15] = hi.w;
      if constexpr (U)
        p1u_compress<8, UMMA_SYNC>(uc, s, c_p1);
      else
        p1_compress<8, P1Tables, LEAF_SYNC != 0>(s, c_p1);
    }
    if (live) {
      dg[0] = make_uint4(s[0], s[1], s[2], s[3]);
      dg[1] = make_uint4(s[4], s[5], s[6], s[7]);
    }
  }
  if constexpr (U) p1u_teardown(uc, LEAF_GROUPS);
}

// One level: next[i] = C(prev[2i] || prev[2i+1])[0..8), one thread per parent.  Used while a level still fills
// the machine; the short tail of the tree goes through tree_levels_kernel below.
template <bool U>
__global__ void __launch_bounds__(U ? UMMA_THREADS : LEAF_THREADS, U ? UMMA_MIN_BLOCKS : LEAF_MIN_BLOCKS)
tree_level_kernel(const uint32_t* __restrict__ prev, uint64_t n_next, uint32_t* __restrict__ next, const uint8_t* __restrict__ b_image) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  P1uCtx uc;
  if constexpr (U) uc = p1u_setup(dsm, b_image, LEAF_GROUPS);
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n_next; base += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = base + threadIdx.x;
    const bool live = i < n_next;
    if (!live) i = n_next - 1;  // CTA-wide barriers inside the permutation: every thread runs it
    uint32_t s[16];
    const uint4* src = reinterpret_cast<const uint4*>(prev + 16 * i);
    const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
    s[0] = a.x, s[1] = a.y, s[2] = a.z, s[3] = a.w, s[4] = b.x, s[5] = b.y, s[6] = b.z, s[7] = b.w;
    s[8] = c.x, s[9] = c.y, s[10] = c.z, s[11] = c.w, s[12] = d.x, s[13] = d.y, s[14] = d.z, s[15] = d.w;
    if constexpr (U)
      p1u_compress<8, UMMA_SYNC>(uc, s, c_p1);
    else
      p1_compress<8, P1Tables, LEAF_SYNC != 0>(s, c_p1);
    if (live) {
      uint4* dst = reinterpret_cast<uint4*>(next + 8 * i);
      dst[0] = make_uint4(s[0], s[1], s[2], s[3]);
      dst[1] = make_uint4(s[4], s[5], s[6], s[7]);
    }
  }
  if constexpr (U) p1u_teardown(uc, LEAF_GROUPS);
}

// layer0: n0 digests (n0 = 2 * T * gridDim.x at full size). CTA b owns digests [b*2T, (b+1)*2T) and writes
// its part of each of the next `levels` layers.  layer_out[l] points at the start of layer (l+1).
template <int T>
__global__ void __launch_bounds__(T)
tree_levels_kernel(const uint32_t* __restrict__ layer0, uint64_t n0, int levels, uint32_t* __restrict__ next_base) {
  __shared__ uint32_t sm[2][T][8 + 1];  // +1: avoid 8-way bank conflicts on the strided reads
  const int t = threadIdx.x;
  uint64_t n_prev = n0;
  uint32_t* out_layer = next_base;
  uint64_t cta_span = 2 * (uint64_t)T;  // digests of the current layer owned by this CTA
  int buf = 0;
  for (int l = 0; l < levels; l++) {
    const uint64_t span_next = cta_span / 2;
    const uint64_t base_next = (uint64_t)blockIdx.x * span_next;
    if ((uint64_t)t < span_next && base_next + t < n_prev / 2) {
      uint32_t s[16];
      if (l == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(layer0 + 16 * (base_next + t));
        uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
        s[0] = a.x, s[1] = a.y, s[2] = a.z, s[3] = a.w, s[4] = b.x, s[5] = b.y, s[6] = b.z, s[7] = b.w;
        s[8] = c.x, s[9] = c.y, s[10] = c.z, s[11] = c.w, s[12] = d.x, s[13] = d.y, s[14] = d.z, s[15] = d.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          s[k] = sm[buf ^ 1][2 * t][k];
          s[8 + k] = sm[buf ^ 1][2 * t + 1][k];
        }
      }
      p1_compress<8>(s, c_p1);
      uint4* dst = reinterpret_cast<uint4*>(out_layer + 8 * (base_next + t));
      dst[0] = make_uint4(s[0], s[1], s[2], s[3]);
      dst[1] = make_uint4(s[4], s[5], s[6], s[7]);
#pragma unroll
      for (int k = 0; k < 8; k++) sm[buf][t][k] = s[k];
    }
    __syncthreads();
    buf ^= 1;
    out_layer += 8 * (n_prev / 2);
    n_prev /= 2;
    cta_span = span_next;
  }
}

// ---- the narrow levels: one WARP per parent ------------------------------------------------------------------------------
// A level with fewer parents than the machine has lanes costs one compression of LATENCY, whatever kernel runs it: ~15-25 us
// with one state per thread, 13 such levels per tree (a quarter of the tree's time on one GPU, a fifth of a whole sharded commit
// on eight).  The transcript's warp-wide permutation (devfs.cuh: the 16 words of a state across the lanes, an MDS row split
// between the half-warps) has ~3.3 us of latency: levels of at most 4096 parents run one warp per parent, one launch per level
// (2^15: measured slower, the warp-wide form has a quarter of the throughput),
// and the last levels (<= 32 parents) in ONE CTA with a barrier between levels.
constexpr int TREE_WARP_MAX_PARENTS = 4096, TREE_WARP_TOP = 32;
__device__ __forceinline__ void tree_warp_compress(const uint32_t* __restrict__ children, uint32_t* __restrict__ parent,
                                                   const uint32_t* rc_s, const uint32_t (&mds_h)[8]) {
  const int lane = threadIdx.x & 31;
  const uint32_t x = children[lane & 15];
  const uint32_t y = kb_add(fs_warp_permute(x, rc_s, mds_h), x);  // compress_in_place: permute + feed-forward, first 8 words kept
  if (lane < 8) parent[lane] = y;
}
__global__ void __launch_bounds__(256) tree_level_warp_kernel(const uint32_t* __restrict__ prev, uint64_t n_next, uint32_t* __restrict__ next) {
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  for (int t = threadIdx.x; t < DEVFS_RC_WORDS; t += blockDim.x) rc_s[t] = (&d_fs_tables.rc[0][0])[t];
  __syncthreads();
  uint32_t mds_h[8];
  fs_load_mds_half(mds_h);
  const uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n_next) return;  // whole warps leave together
  tree_warp_compress(prev + 16 * i, next + 8 * i, rc_s, mds_h);
}
// layer0: n0 <= 2 * TREE_WARP_TOP digests; all remaining levels, layer (l + 1) behind layer l
__global__ void __launch_bounds__(32 * TREE_WARP_TOP) tree_top_warp_kernel(uint32_t* layer0, uint64_t n0) {
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  for (int t = threadIdx.x; t < DEVFS_RC_WORDS; t += blockDim.x) rc_s[t] = (&d_fs_tables.rc[0][0])[t];
  __syncthreads();
  uint32_t mds_h[8];
  fs_load_mds_half(mds_h);
  const uint64_t w = threadIdx.x >> 5;
  uint32_t* cur = layer0;
  for (uint64_t n = n0; n > 1; n >>= 1) {
    uint32_t* next = cur + 8 * n;
    if (w < n / 2) tree_warp_compress(cur + 16 * w, next + 8 * w, rc_s, mds_h);
    __syncthreads();  // the next level reads what this CTA just wrote to global memory
    cur = next;
  }
}

static State16 zero_suffix_state_host(uint32_t n_zero_chunks) {
  // sponge.rs:28-48 evaluated with the same arithmetic header on the host (a 16-word constant per commit shape)
  State16 st;
  for (int i = 0; i < 16; i++) st.v[i] = 0;
  p1_compress<16>(st.v, h_p1);
  for (uint32_t k = 0; k + 2 < n_zero_chunks; k++) {
    for (int i = 8; i < 16; i++) st.v[i] = 0;
    p1_compress<16>(st.v, h_p1);
  }
  return st;
}

// CTAs of a leaf kernel over h rows: one per LEAF_THREADS rows, capped at LM_LEAF_GRID (environment, 0 = no cap) CTAs
static uint64_t leaf_grid(uint64_t h, int threads = LEAF_THREADS) {
  static const long cap = [] {
    const char* e = getenv("LM_LEAF_GRID");
    return e ? atol(e) : (long)LEAF_DEFAULT_GRID;
  }();
  const uint64_t blocks = (h + threads - 1) / threads;
  return cap > 0 && blocks > (uint64_t)cap ? (uint64_t)cap : blocks;
}

cudaError_t merkle_leaf_digests(cudaStream_t stream, const uint32_t* d_mat, uint64_t h, uint32_t stored_w,
                                uint32_t full_w, uint32_t eff_w, uint32_t* d_digests) {
  if (h == 0) return cudaSuccess;
  if (full_w % 8 != 0 || full_w < 16 || eff_w > full_w || stored_w > full_w) return cudaErrorInvalidValue;
  const uint32_t n_zero_chunks = (full_w - eff_w) / 8;
  State16 init{};
  uint32_t lim, virt_w;
  int from_state = 0;
  if (n_zero_chunks >= 2) {
    init = zero_suffix_state_host(n_zero_chunks);
    from_state = 1;
    lim = eff_w < stored_w ? eff_w : stored_w;
    virt_w = (eff_w + 7) / 8 * 8;
  } else {
    lim = stored_w;
    virt_w = full_w;
  }
  const int T = LEAF_THREADS;
  const uint64_t blocks = leaf_grid(h);
  if (use_umma()) {
    const uint64_t ublocks = leaf_grid(h, UMMA_THREADS);
    const uint8_t* img = nullptr;
    cudaError_t e = umma_b_image(&img);
    if (e != cudaSuccess) return e;
    if ((e = umma_attr(leaf_sponge_kernel<true>, LEAF_GROUPS)) != cudaSuccess) return e;
    leaf_sponge_kernel<true><<<(unsigned)ublocks, UMMA_THREADS, p1u_smem_bytes(LEAF_GROUPS), stream>>>(
        d_mat, h, stored_w, lim, virt_w, from_state, init, d_digests, img);
  } else {
    leaf_sponge_kernel<false><<<(unsigned)blocks, T, 0, stream>>>(d_mat, h, stored_w, lim, virt_w, from_state, init, d_digests, nullptr);
  }
  count_launch();
  return cudaGetLastError();
}

// true when the leaf sponge can be run one rate chunk at a time starting from the zero-suffix state
bool merkle_leaf_chunked_ok(uint32_t stored_w, uint32_t full_w, uint32_t eff_w) {
  return full_w % 8 == 0 && eff_w % 8 == 0 && stored_w % 8 == 0 && eff_w <= stored_w && eff_w > 0 && (full_w - eff_w) / 8 >= 2;
}
// absorb chunks chunk_hi, chunk_hi - 1, .., chunk_hi - count + 1 (chunks must be fed from eff_w / 8 - 1 down to 0); the
// call that takes chunk eff_w / 8 - 1 seeds the state
cudaError_t merkle_leaf_absorb_chunks(cudaStream_t stream, const uint32_t* d_mat, uint64_t h, uint32_t stored_w,
                                      uint32_t full_w, uint32_t eff_w, uint32_t chunk_hi, uint32_t count, uint32_t* d_digests) {
  if (!merkle_leaf_chunked_ok(stored_w, full_w, eff_w) || chunk_hi >= eff_w / 8 || count == 0 || count > chunk_hi + 1)
    return cudaErrorInvalidValue;
  if (h == 0) return cudaSuccess;
  const int first = chunk_hi == eff_w / 8 - 1;
  State16 init{};
  if (first) init = zero_suffix_state_host((full_w - eff_w) / 8);
  if (use_umma()) {
    const uint8_t* img = nullptr;
    cudaError_t e = umma_b_image(&img);
    if (e != cudaSuccess) return e;
    if ((e = umma_attr(leaf_absorb_kernel<true>, LEAF_GROUPS)) != cudaSuccess) return e;
    leaf_absorb_kernel<true><<<(unsigned)leaf_grid(h, UMMA_THREADS), UMMA_THREADS, p1u_smem_bytes(LEAF_GROUPS), stream>>>(
        d_mat, h, stored_w, chunk_hi, count, first, init, d_digests, img);
  } else {
    leaf_absorb_kernel<false><<<(unsigned)leaf_grid(h), LEAF_THREADS, 0, stream>>>(d_mat, h, stored_w, chunk_hi, count, first, init,
                                                                                   d_digests, nullptr);
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t merkle_tree_from_digests(cudaStream_t stream, uint32_t* d_layers, uint64_t h) {
  // layers back to back: h, h/2, ..., 1 digests
  constexpr int T = 128;
  uint32_t* cur = d_layers;
  uint64_t n = h;
  // wide levels: one launch per level, every thread busy
  const uint8_t* img = nullptr;
  static const bool warp_tail = getenv("LM_TREE_TAIL_THREADS") == nullptr;  // =1: the one-state-per-thread tail kernels
  const uint64_t wide_min = 2 * (uint64_t)TREE_WARP_MAX_PARENTS;  // parents of the narrowest "wide" level (8192)
  if (use_umma() && h / 2 >= wide_min) {
    cudaError_t e = umma_b_image(&img);
    if (e != cudaSuccess) return e;
    if ((e = umma_attr(tree_level_kernel<true>, LEAF_GROUPS)) != cudaSuccess) return e;
  }
  while (n / 2 >= wide_min) {
    uint32_t* next = cur + 8 * n;
    if (img)
      tree_level_kernel<true><<<(unsigned)leaf_grid(n / 2, UMMA_THREADS), UMMA_THREADS, p1u_smem_bytes(LEAF_GROUPS), stream>>>(
          cur, n / 2, next, img);
    else
      tree_level_kernel<false><<<(unsigned)((n / 2 + LEAF_THREADS - 1) / LEAF_THREADS), LEAF_THREADS, 0, stream>>>(cur, n / 2, next, nullptr);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cur = next;
    n >>= 1;
  }
  while (warp_tail && n > 1) {
    if (n / 2 > (uint64_t)TREE_WARP_MAX_PARENTS) break;
    if (n <= 2 * (uint64_t)TREE_WARP_TOP) {
      tree_top_warp_kernel<<<1, 32 * TREE_WARP_TOP, 0, stream>>>(cur, n);
      count_launch();
      return cudaGetLastError();
    }
    uint32_t* next = cur + 8 * n;
    tree_level_warp_kernel<<<(unsigned)((n / 2 + 7) / 8), 256, 0, stream>>>(cur, n / 2, next);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cur = next;
    n >>= 1;
  }
  while (n > 1) {
    int levels = 0;
    uint64_t m = n;
    while (m > 1 && levels < 8) m >>= 1, levels++;  // 2T = 256 digests per CTA -> up to 8 levels
    const uint64_t blocks = (n + 2 * T - 1) / (2 * T);
    uint32_t* next = cur + 8 * n;
    tree_levels_kernel<T><<<(unsigned)blocks, T, 0, stream>>>(cur, n, levels, next); count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    for (int l = 0; l < levels; l++) {
      cur += 8 * n;
      n >>= 1;
    }
  }
  return cudaSuccess;
}

// Openings (crates/whir/src/merkle.rs:205-211 + crates/backend/symetric/src/merkle.rs:43-47): one CTA per query
// packs the zero-extended row and the sibling of every level into contiguous buffers (one D2H copy each).
__global__ void open_gather_kernel(const uint32_t* __restrict__ mat, const uint32_t* __restrict__ layers, uint64_t h,
                                   uint32_t log_h, uint32_t stored_w, uint32_t full_w,
                                   const uint64_t* __restrict__ indices, uint32_t* __restrict__ rows,
                                   uint32_t* __restrict__ paths) {
  const uint32_t q = blockIdx.x;
  const uint64_t idx = indices[q];
  for (uint32_t c = threadIdx.x; c < full_w; c += blockDim.x)
    rows[(uint64_t)q * full_w + c] = c < stored_w ? mat[idx * stored_w + c] : 0u;
  for (uint32_t k = threadIdx.x; k < log_h * 8; k += blockDim.x) {
    const uint32_t l = k >> 3;
    // layer l starts at digest offset 2h - 2h/2^l
    const uint64_t off = 2 * h - ((2 * h) >> l);
    paths[(uint64_t)q * log_h * 8 + k] = layers[8 * (off + ((idx >> l) ^ 1)) + (k & 7)];
  }
}

cudaError_t merkle_open_gather(cudaStream_t stream, const uint32_t* d_mat, const uint32_t* d_layers, uint64_t h,
                               uint32_t stored_w, uint32_t full_w, const uint64_t* d_indices, uint32_t n,
                               uint32_t* d_rows, uint32_t* d_paths) {
  if (n == 0) return cudaSuccess;
  uint32_t log_h = 0;
  while (((uint64_t)1 << log_h) < h) log_h++;
  open_gather_kernel<<<n, 128, 0, stream>>>(d_mat, d_layers, h, log_h, stored_w, full_w, d_indices, d_rows, d_paths); count_launch();
  return cudaGetLastError();
}

// Verifier side: the openings of one STIR round checked against the round's root in one launch.
// Replaces the per-query loops of crates/whir/src/verify.rs:313-318,333-338 over
// crates/backend/symetric/src/merkle.rs:92-122 (merkle_verify) and sponge.rs:7-25 (hash_slice).
// One WARP per opening on the transcript's warp-wide permutation (a STIR round has 20..230 openings: latency, not
// throughput): lane j & 15 carries word j of the sponge / compression state.  A word that is not a canonical Montgomery
// residue (>= p) fails the opening instead of entering the arithmetic.
__global__ void __launch_bounds__(256) verify_openings_kernel(const uint32_t* __restrict__ root, uint32_t log_h,
                                                              const uint64_t* __restrict__ indices, uint32_t n,
                                                              const uint32_t* __restrict__ rows, uint32_t width,
                                                              const uint32_t* __restrict__ paths, uint32_t* __restrict__ ok) {
  __shared__ uint32_t rc_s[DEVFS_RC_WORDS];
  for (int t = threadIdx.x; t < DEVFS_RC_WORDS; t += blockDim.x) rc_s[t] = (&d_fs_tables.rc[0][0])[t];
  __syncthreads();
  uint32_t mds_h[8];
  fs_load_mds_half(mds_h);
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n) return;  // whole warps leave together
  const int lane = threadIdx.x & 31, j = lane & 15;
  const uint32_t* row = rows + (uint64_t)q * width;
  bool good = true;
  // hash_slice: the last 16 words, then one rate chunk at a time towards the front
  uint32_t x = row[width - 16 + j];
  good &= x < KB_P;
  uint32_t y = kb_add(fs_warp_permute(x, rc_s, mds_h), x);
  for (int64_t c = (int64_t)(width >> 3) - 3; c >= 0; c--) {
    if (j >= 8) {
      x = row[c * 8 + (j - 8)];
      good &= x < KB_P;
    } else {
      x = y;
    }
    y = kb_add(fs_warp_permute(x, rc_s, mds_h), x);
  }
  // the path, leaf level first: (node, sibling) or (sibling, node) by the index bit
  uint64_t idx = indices[q];
  const uint32_t* sib = paths + (uint64_t)q * log_h * 8;
  for (uint32_t l = 0; l < log_h; l++, sib += 8, idx >>= 1) {
    const uint32_t moved = __shfl_sync(0xffffffffu, y, lane ^ 8);
    const bool node_left = (idx & 1) == 0;
    if ((j < 8) == node_left) {
      x = j < 8 ? y : moved;
    } else {
      x = sib[j & 7];
      good &= x < KB_P;
    }
    y = kb_add(fs_warp_permute(x, rc_s, mds_h), x);
  }
  good &= idx == 0;  // the index addresses a leaf of this tree
  if (j < 8) good &= y == root[j];
  good = __all_sync(0xffffffffu, good);
  if (lane == 0) ok[q] = good ? 1u : 0u;
}

cudaError_t merkle_verify_openings(cudaStream_t stream, const uint32_t* d_root, uint32_t log_h, const uint64_t* d_indices,
                                   uint32_t n, const uint32_t* d_rows, uint32_t width, const uint32_t* d_paths, uint32_t* d_ok) {
  if (n == 0) return cudaSuccess;
  if (width < 16 || (width & 7)) return cudaErrorInvalidValue;
  verify_openings_kernel<<<(n + 7) / 8, 256, 0, stream>>>(d_root, log_h, d_indices, n, d_rows, width, d_paths, d_ok);
  count_launch();
  return cudaGetLastError();
}


// Batched permutation / compression of explicit states (parity tests, PoW grinding building block).
__global__ void __launch_bounds__(128) permute_states_kernel(uint32_t* states, uint64_t n, int compress) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[16];
  uint4* p = reinterpret_cast<uint4*>(states + 16 * i);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint4 v = p[k];
    s[4 * k] = v.x, s[4 * k + 1] = v.y, s[4 * k + 2] = v.z, s[4 * k + 3] = v.w;
  }
  if (compress)
    p1_compress<16>(s, c_p1);
  else
    p1_permute<16>(s, c_p1);
#pragma unroll
  for (int k = 0; k < 4; k++) p[k] = make_uint4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
}

// the same on the tensor cores: one group of 128 states per CTA, every thread runs the permutation (clamped index)
__global__ void __launch_bounds__(128) permute_states_umma_kernel(uint32_t* states, uint64_t n, int compress,
                                                                  const uint8_t* __restrict__ b_image) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  P1uCtx uc = p1u_setup(dsm, b_image, 1);
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;
  uint32_t s[16];
  uint4* p = reinterpret_cast<uint4*>(states + 16 * i);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint4 v = p[k];
    s[4 * k] = v.x, s[4 * k + 1] = v.y, s[4 * k + 2] = v.z, s[4 * k + 3] = v.w;
  }
  if (compress)
    p1u_compress<16, false>(uc, s, c_p1);
  else
    p1u_permute<16, false>(uc, s, c_p1);
  // a dead thread re-computed state n - 1: it must not race with the live thread's store
  if (live) {
#pragma unroll
    for (int k = 0; k < 4; k++) p[k] = make_uint4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
  }
  p1u_teardown(uc, 1);
}

cudaError_t poseidon1_states(cudaStream_t stream, uint32_t* d_states, uint64_t n, int compress) {
  if (n == 0) return cudaSuccess;
  if (use_umma()) {
    const uint8_t* img = nullptr;
    cudaError_t e = umma_b_image(&img);
    if (e != cudaSuccess) return e;
    if ((e = umma_attr(permute_states_umma_kernel, 1)) != cudaSuccess) return e;
    permute_states_umma_kernel<<<(unsigned)((n + 127) / 128), 128, p1u_smem_bytes(1), stream>>>(d_states, n, compress, img);
  } else {
    permute_states_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_states, n, compress);
  }
  count_launch();
  return cudaGetLastError();
}

// Fiat-Shamir proof-of-work search (reference: ProverState::pow_grinding, fiat-shamir/src/prover.rs:135-167):
// witness w (canonical, < p) is accepted when lane 8 of permute(capacity | w | 0^7), read as a canonical integer,
// has its low `bits` bits clear.  One candidate per thread; the smallest hit of the batch wins (atomicMin), so the
// search is deterministic where the reference's rayon find_any is not.
__global__ void __launch_bounds__(128) pow_grind_kernel(State16 base, uint64_t start, uint64_t n, uint32_t mask,
                                                        unsigned long long* best) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t w = start + i;
  uint32_t s[16];
#pragma unroll
  for (int k = 0; k < 8; k++) s[k] = base.v[k];
  s[8] = kb_mul((uint32_t)w, KB_R2);
#pragma unroll
  for (int k = 9; k < 16; k++) s[k] = 0;
  p1_permute<9>(s, c_p1);
  const uint32_t canon = kb_canon(kb_redc_lazy((uint64_t)s[8]));
  if ((canon & mask) == 0) atomicMin(best, (unsigned long long)w);
}

// the same search with the permutation on the tensor cores (every thread of a CTA runs it: out-of-range candidates are clamped
// to the last one and not reported)
__global__ void __launch_bounds__(128, 4) pow_grind_umma_kernel(State16 base, uint64_t start, uint64_t n, uint32_t mask,
                                                                unsigned long long* best, const uint8_t* __restrict__ b_image) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  P1uCtx uc = p1u_setup(dsm, b_image, 1);
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  if (!live) i = n - 1;
  const uint64_t w = start + i;
  uint32_t s[16];
#pragma unroll
  for (int k = 0; k < 8; k++) s[k] = base.v[k];
  s[8] = kb_mul((uint32_t)w, KB_R2);
#pragma unroll
  for (int k = 9; k < 16; k++) s[k] = 0;
  p1u_permute<16, false>(uc, s, c_p1);
  const uint32_t canon = kb_canon(kb_redc_lazy((uint64_t)s[8]));
  if (live && (canon & mask) == 0) atomicMin(best, (unsigned long long)w);
  p1u_teardown(uc, 1);
}

cudaError_t pow_grind(cudaStream_t stream, const uint32_t state[16], uint32_t bits, uint64_t start,
                      unsigned long long* d_best, uint64_t* witness) {
  if (bits >= 31) return cudaErrorInvalidValue;
  State16 base;
  for (int i = 0; i < 16; i++) base.v[i] = state[i];
  const uint32_t mask = (1u << bits) - 1u;
  // expected 2^bits candidates: one launch of 4 * 2^bits (at least 2^16) finds a witness with probability 1 - e^-4
  uint64_t batch = (uint64_t)4 << bits;
  if (batch < (1u << 16)) batch = 1u << 16;
  if (batch > (1u << 24)) batch = 1u << 24;
  const unsigned long long none = ~0ull;
  for (uint64_t at = start; at < KB_P; at += batch) {
    const uint64_t n = at + batch <= KB_P ? batch : KB_P - at;
    cudaError_t e = cudaMemcpyAsync(d_best, &none, sizeof(none), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    if (use_umma()) {
      const uint8_t* img = nullptr;
      if ((e = umma_b_image(&img)) != cudaSuccess) return e;
      if ((e = umma_attr(pow_grind_umma_kernel, 1)) != cudaSuccess) return e;
      pow_grind_umma_kernel<<<(unsigned)((n + 127) / 128), 128, p1u_smem_bytes(1), stream>>>(base, at, n, mask, d_best, img);
    } else {
      pow_grind_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(base, at, n, mask, d_best);
    }
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    unsigned long long got = none;
    if ((e = cudaMemcpyAsync(&got, d_best, sizeof(got), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    if (got != none) {
      *witness = got;
      return cudaSuccess;
    }
  }
  return cudaErrorNotReady;  // no witness below p (probability ~ e^(-p / 2^bits))
}

// The transcript sponge itself (challenger.rs:8-76) is one permutation per 8 absorbed words, strictly sequential:
// it stays on the host, evaluated with the same arithmetic header as the device code.
void poseidon1_permute_host(uint32_t state[16]) { p1_permute<16>(state, h_p1); }

// the B image of the tensor-core formulation as the kernels load it (P1U_B_BYTES bytes, layout in poseidon1_umma.cuh)
size_t poseidon1_umma_image_host(uint8_t* out, size_t capacity) {
  if (out && capacity >= (size_t)P1U_B_BYTES) p1u_build_b_image(h_p1, out);
  return (size_t)P1U_B_BYTES;
}
// CPU model of the tensor-core formulation (poseidon1_umma.cuh): the B image the kernels load, the MMAs as integer dot products
void poseidon1_permute_umma_model_host(uint32_t state[16]) {
  static const std::vector<uint8_t> img = [] {
    std::vector<uint8_t> v(P1U_B_BYTES);
    p1u_build_b_image(h_p1, v.data());
    return v;
  }();
  p1u_model_permute(h_p1, img.data(), state);
}

}  // namespace lm

// Reed-Solomon encode of the stacked witness on sm_100a: block gather + radix-2 "DFT on evaluations".
//
// Device replacement for
//   crates/whir/src/utils.rs:69-150   reorder_and_dft / prepare_evals_for_fft_unpacked
//   crates/whir/src/dft.rs:52-62      roots_of_unity_table
//   crates/whir/src/dft.rs:79-155     dft_batch_by_evals / dft_algebra_batch_by_evals
//   crates/whir/src/dft.rs:546-568    butterflies (a, b) -> (a + t (b - a), a - t (b - a))
//
// Layout.  The matrix is row-major h x w u32 (an EF matrix is the same memory with w = 5 * n_cols, dft.rs:147-155).
// The transform is independent per column; layer l (half-block m = 2^l) pairs rows (i, i + m) with twiddle
// w_h^((i mod m) * h / 2m).  A pass kernel keeps a tile of 2^L rows x 8 columns (32 B per row = one DRAM sector)
// in shared memory and runs L consecutive layers on it, radix-8 in registers between barriers, so a 2^22-row
// transform is two passes over HBM (11 + 11 layers) instead of the CPU's eight.  Rows of a tile are 2^l0 apart
// (l0 = first layer of the pass); the first pass can gather straight from the evaluation vector, where column j
// is the contiguous slice evals[j << (n - k) ..] with every value repeated 2^r times, so the first r layers
// (identities on repeated data) are skipped.
//
// Tile movement.  In-place passes load their tile with TMA: the matrix is described to the hardware as a 3-D tensor
// (column, row mod 2^l0, row div 2^l0), a tile is 2^L / 256 boxes of 8 columns x 1 x 256 rows, fetched by ONE elected
// thread with cp.async.bulk.tensor into shared memory and signalled through an mbarrier; every pass writes its tile back
// with TMA stores (cp.async.bulk.tensor global <- shared).  This replaces 16 cp.async + 16 st.global and their address
// arithmetic per thread (the pass kernels are issue-slot bound, ncu 66-68 %): 2.31 -> 2.00 ms on the 2^22 x 64 transform.
//
// Variants of the pass (round 2).  Domains above 2^22 rows: tiles of 2^12 rows, one 1024-thread CTA per SM, so that 2^23 and 2^24
// rows stay at two passes (BIG_TILE_LOG).  Row-sharded commit: the last pass stores its tile into the matrices of the ranks that
// own the rows after the exchange, by TMA through one tensor map per peer (NttPeerMaps), and runs on tiles of 16 or 32 columns
// (ntt_pass_wide_kernel) because the NVLink fabric carries 128-byte row pieces 2.3x faster than 32-byte ones.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cuda_pipeline.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "launch_count.h"
#include "kb.cuh"
#include "ntt.h"

namespace lm {

// ------------------------------------------------------------------------------------------------ twiddles
// tw[e] = g^e for e < 2^(log_n - 1), g = primitive 2^log_n-th root (koala_bear.rs:46-54), Montgomery form.
static const uint32_t TWO_ADIC_GEN_CANON[25] = {
    0x1,        0x7f000000, 0x7e010002, 0x6832fe4a, 0x8dbd69c,  0xa28f031,  0x5c4a5b99, 0x29b75a80, 0x17668b8a,
    0x27ad539b, 0x334d48c7, 0x7744959c, 0x768fc6fa, 0x303964b2, 0x3e687d4d, 0x45a60e61, 0x6e2f4d7a, 0x163bd499,
    0x6c4a8a45, 0x143ef899, 0x514ddcad, 0x484ef19b, 0x205d63c3, 0x68e7dd49, 0x6ac49f88,
};

uint32_t two_adic_generator_monty(unsigned bits) { return kb_mul(TWO_ADIC_GEN_CANON[bits], KB_R2); }

__global__ void twiddle_kernel(uint32_t* tw, uint64_t n_half, uint32_t g) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_half) return;
  uint32_t r = KB_R1, b = g;
  for (uint64_t k = e; k; k >>= 1) {
    if (k & 1) r = kb_mul(r, b);
    b = kb_mul(b, b);
  }
  tw[e] = r;
}

cudaError_t ntt_fill_twiddles(cudaStream_t stream, uint32_t* d_tw, unsigned log_n) {
  if (log_n == 0) return cudaSuccess;
  const uint64_t n_half = (uint64_t)1 << (log_n - 1);
  const uint32_t g = two_adic_generator_monty(log_n);
  twiddle_kernel<<<(unsigned)((n_half + 255) / 256), 256, 0, stream>>>(d_tw, n_half, g); count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ gather
// out[i][j*dim + d] = evals[(((j << log_block) + i) >> r) * dim + d]      (utils.rs:128-150)
__global__ void gather_kernel(const uint32_t* __restrict__ evals, uint32_t* __restrict__ out, uint64_t h, uint32_t n_cols,
                              uint32_t dim, uint32_t log_block, uint32_t r) {
  const uint64_t w = (uint64_t)n_cols * dim;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= h * w) return;
  const uint64_t i = idx / w, c = idx % w;
  const uint64_t j = c / dim, d = c % dim;
  out[idx] = __ldg(evals + ((((j << log_block) + i) >> r) * dim + d));
}

// ------------------------------------------------------------------------------------------------ butterflies
__device__ __forceinline__ void bfly(uint32_t& a, uint32_t& b, uint32_t t) {
  const uint32_t x = kb_mul(kb_sub(b, a), t);
  b = kb_sub(a, x);
  a = kb_add(a, x);
}
__device__ __forceinline__ void bfly4(uint4& a, uint4& b, uint32_t t) {
  bfly(a.x, b.x, t);
  bfly(a.y, b.y, t);
  bfly(a.z, b.z, t);
  bfly(a.w, b.w, t);
}

#ifndef NTT_THREADS
#define NTT_THREADS 256
#endif
#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 3
#endif
constexpr int NTT_MAX_PEERS = 16;
constexpr int TILE_COLS = 8;       // u32 columns per tile = 32 B per row
constexpr int MAX_TILE_LOG = 11;   // 2048 rows x 32 B = 64 KiB of shared memory: three 256-thread CTAs per SM
// Domains above 2^22 rows: tiles of 2^12 rows (128 KiB, ONE 1024-thread CTA per SM - the same 32 warps) keep 2^23 and 2^24 rows
// at two passes over HBM; with 2^11-row tiles 2^24 rows need three 8-layer passes of small, sector-granular tiles (16.5 ms
// against 1.85 ms for 2^22 rows, profiles/r02_p1_umma.txt section 8).
constexpr int BIG_TILE_LOG = 12;

// 16-byte slot of (row j, half) inside the shared tile: dense row-major, 32 B per row - the layout a TMA box has.
// Layout study on the 2^22 x 64 transform (tools/sweep_ntt_tma.sh, profiles/r02_ntt_tma_sweep.txt): dense + TMA 2.00 ms, dense
// or XOR-swizzled with cp.async loads / st.global stores 2.31-2.32 ms, i.e. once the tile moves by TMA the 4-way bank
// conflicts of the first register group (rows 8 apart, 32-byte pitch) no longer decide the time.  TMA's own swizzle modes
// cannot remove them at this row pitch: the 128-byte pattern needs 128-byte rows (a 32-byte inner box faults), the 32-byte
// pattern does not touch the bits that conflict.
#if defined(NTT_SWIZZLE_XOR)
__device__ __forceinline__ int tile_slot(int j, int half) { return ((j << 1) | half) ^ (((j >> 3) & 3) << 1); }
#else
__device__ __forceinline__ int tile_slot(int j, int half) { return (j << 1) | half; }
#endif

// ---- TMA / mbarrier primitives (PTX ISA: cp.async.bulk.tensor, mbarrier) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  // bounded: a transaction count that never completes (a malformed tensor map) must fault, not hang the device
  for (uint32_t spins = 0;; spins++) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    if (done) return;
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
constexpr int NTT_TMA_LOAD = 1, NTT_TMA_STORE = 2;
constexpr int NTT_BOX_ROWS = 256;  // TMA box dimensions are at most 256

// One group of G layers (local layers lp .. lp+G-1) on the shared tile, radix 2^G in registers.
// Work item = (u, half): rows j0 + q * 2^lp (q < 2^G) of 16-byte half `half`.
// tw_s: per-CTA twiddles, tw_s[(1 << ll) + i] = twiddle of local layer ll for (row index mod 2^ll) == i.
template <int G>
__device__ __forceinline__ void tile_group(uint4* tile, const uint32_t* tw_s, int L, int lp) {
  constexpr int Q = 1 << G;
  const int n_items = 1 << (L - G + 1);
  for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
    const int half = item & 1;
    const int u = item >> 1;
    const int j_lo = u & ((1 << lp) - 1);
    const int j0 = ((u >> lp) << (lp + G)) | j_lo;
    uint4 v[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = tile[tile_slot(j0 + (q << lp), half)];
#pragma unroll
    for (int s = 0; s < G; s++) {
      const uint32_t* tw_l = tw_s + (1 << (lp + s));
#pragma unroll
      for (int q = 0; q < Q; q++) {
        if (q & (1 << s)) continue;
        const uint32_t t = tw_l[j_lo + ((q & ((1 << s) - 1)) << lp)];
        bfly4(v[q], v[q | (1 << s)], t);
      }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) tile[tile_slot(j0 + (q << lp), half)] = v[q];
  }
}

// Compact twiddle table of one pass: tw_pass[(1 << ll) + i] = w_h^((i << l0) << (log_h - (l0 + ll) - 1)), i < 2^ll, ll < L:
// the twiddle of local layer ll for a tile whose rows start at a multiple of 2^(l0 + L).  Read contiguously by every
// CTA of the pass (8 KiB, L1/L2 resident) instead of 2^L scattered words of the big table per tile.
__global__ void ntt_pass_twiddles_kernel(uint32_t* __restrict__ tw_pass, int log_h, int l0, int L,
                                         const uint32_t* __restrict__ tw, int tw_shift) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 1 || k >= (1 << L)) return;
  const int ll = 31 - __clz(k);
  const uint64_t i = (uint64_t)k - ((uint64_t)1 << ll);
  const uint64_t e = (i << l0) << (log_h - (l0 + ll) - 1);
  tw_pass[k] = __ldg(tw + (e << tw_shift));
}

// Fused all-to-all of the row-sharded commit: when enabled, the LAST pass of a rank's local transform
// stores row i of its result straight into the matrix of the rank that owns it after the exchange — rank i >> log_run,
// local row (this rank << log_run) | (i mod 2^log_run) — through peer pointers (NVLink stores, 32 contiguous bytes per
// row and tile), so the exchange needs no separate collective and overlaps the butterflies of the other tiles.
struct NttScatter {
  uint32_t* dst[NTT_MAX_PEERS];
  int log_run, rank, enabled;
  int tma_box_rows;  // > 0: the tile leaves through TMA stores into the peers' tensor maps, boxes of this many rows
};
// Tensor maps of the peers' matrices for the scattering pass, (column, row mod 2^l0, row div 2^l0) like the local one.
// tools/microbench/peer_store_probe.cu: SM stores reach a peer's memory at 176 GB/s whatever their width (32 or 128 bytes per
// thread), TMA stores of the same 32-byte row pieces at 426 GB/s — and they do not hold the storing warps.
struct NttPeerMaps {
  CUtensorMap m[NTT_MAX_PEERS];
};

// Pass over layers [l0 + skip, l0 + L) of an h x w matrix; one CTA per (column tile, row group).
// src == nullptr: in place on `mat`; the tile arrives by 16-byte cp.async (all loads of a CTA in flight at once, no
// register staging).  src != nullptr (only with l0 == 0): gather from the evaluation vector (dim = 1):
// element (row, col) = src[(col << log_block) + (row >> r)].
template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
ntt_pass_kernel(uint32_t* __restrict__ mat, const uint32_t* __restrict__ src, uint64_t w, int log_h, int l0, int L,
                int skip, uint32_t log_block, uint32_t r, const uint32_t* __restrict__ tw, int tw_shift,
                uint32_t tile0, uint32_t n_col_tiles, const uint32_t* __restrict__ tw_pass, const NttScatter sc,
                const __grid_constant__ CUtensorMap tmap, int tma, const __grid_constant__ NttPeerMaps peer_maps) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // [2^L][2] slots (TMA destination / source: 128-byte aligned), 2^L twiddle words, the mbarrier of the tile load
  uint4* tile = reinterpret_cast<uint4*>(smem_raw);
  uint32_t* tw_s = reinterpret_cast<uint32_t*>(tile + ((size_t)2 << L));
  uint64_t& tma_bar = *reinterpret_cast<uint64_t*>(tw_s + ((size_t)1 << L));
  const uint64_t col0 = (uint64_t)(tile0 + blockIdx.x % n_col_tiles) * TILE_COLS;  // column tile varies fastest: CTAs that
  const uint64_t grp = blockIdx.x / n_col_tiles;                           // run together cover whole rows
  // rows of this tile: row(j) = ((grp >> l0) << (l0 + L)) + (grp & (2^l0 - 1)) + j * 2^l0
  const uint64_t row_lo = grp & (((uint64_t)1 << l0) - 1);
  const uint64_t row_base = ((grp >> l0) << (l0 + L)) + row_lo;
  const int n_rows = 1 << L;
  const bool two_halves = col0 + 8 <= w;  // w % 4 == 0 guaranteed by the launcher

  const int box_rows = n_rows < NTT_BOX_ROWS ? n_rows : NTT_BOX_ROWS;
  const int row_hi0 = (int)((grp >> l0) << L);  // coordinate of the tile's first row in the (row div 2^l0) dimension
  if (src == nullptr && (tma & NTT_TMA_LOAD)) {
    if (threadIdx.x == 0) mbar_init(&tma_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&tma_bar, (uint32_t)n_rows * 32u);
      for (int t = 0; t < n_rows; t += box_rows) tma_load_3d(tile + 2 * t, &tmap, (int)col0, (int)row_lo, row_hi0 + t, &tma_bar);
    }
  } else if (src == nullptr) {
    for (int item = threadIdx.x; item < 2 * n_rows; item += blockDim.x) {
      const int j = item >> 1, half = item & 1;
      if (half == 0 || two_halves) {
        const uint64_t row = row_base + ((uint64_t)j << l0);
        __pipeline_memcpy_async(&tile[tile_slot(j, half)], mat + row * w + col0 + 4 * half, 16);
      } else {
        tile[tile_slot(j, half)] = make_uint4(0, 0, 0, 0);
      }
    }
    __pipeline_commit();
  } else {
    uint32_t* t32 = reinterpret_cast<uint32_t*>(tile);
    const int n_cols_here = two_halves ? 8 : 4;
    const int rpl = 4 << r;  // rows covered by one 16-byte load of a column (every value is repeated 2^r times)
    if (log_block >= r + 2 && n_rows >= rpl && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // lane -> (column c = lane & 7, row group): 4 lanes read 64 contiguous bytes of one column, and the 8 columns x
      // 4 row groups of a warp store to 32 distinct banks (tile_slot swizzles rows 8 / 16 apart)
      const int n_items = (n_rows / rpl) * 8;
      for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
        const int c = item & 7, j_first = (item >> 3) * rpl;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < n_cols_here)
          v = __ldg(reinterpret_cast<const uint4*>(src + (((col0 + c) << (log_block - r)) + ((row_base + j_first) >> r))));
        const uint32_t vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; e++)
          for (int rep = 0; rep < (1 << r); rep++) {
            const int j = j_first + (e << r) + rep;
            t32[tile_slot(j, c >> 2) * 4 + (c & 3)] = vals[e];
          }
      }
    } else {
      // column c of the tile is contiguous in src: consecutive threads read consecutive rows of one column
      for (int item = threadIdx.x; item < 8 * n_rows; item += blockDim.x) {
        const int c = item >> L, j = item & (n_rows - 1);
        const uint64_t row = row_base + j;
        const uint32_t v = c < n_cols_here ? __ldg(src + ((((col0 + c) << log_block) + row) >> r)) : 0u;
        t32[tile_slot(j, c >> 2) * 4 + (c & 3)] = v;
      }
    }
  }

  // twiddles of this CTA: layer l = l0 + ll pairs rows whose index mod 2^l is row_lo + (i << l0), i < 2^ll, with
  // w_h^((row mod 2^l) << (log_h - l - 1)) = tw_pass[2^ll + i] * w_h^(row_lo << (log_h - l - 1))
  for (int k = threadIdx.x + 1; k < n_rows; k += blockDim.x) {
    uint32_t t = __ldg(tw_pass + k);
    if (row_lo) {
      const int ll = 31 - __clz(k);
      t = kb_mul(t, __ldg(tw + ((row_lo << (log_h - (l0 + ll) - 1)) << tw_shift)));
    }
    tw_s[k] = t;
  }
  if (src == nullptr && (tma & NTT_TMA_LOAD))
    mbar_wait(&tma_bar, 0);
  else if (src == nullptr)
    __pipeline_wait_prior(0);
  __syncthreads();

  int lp = skip;
  while (lp < L) {
    const int rem = L - lp;
    const int g = rem <= 4 ? (rem == 4 ? 2 : rem) : 3;  // 3,3,...,then 3 / 2+2 / 2 / 1
    if (g == 3)
      tile_group<3>(tile, tw_s, L, lp);
    else if (g == 2)
      tile_group<2>(tile, tw_s, L, lp);
    else
      tile_group<1>(tile, tw_s, L, lp);
    lp += g;
    __syncthreads();
  }

  if ((tma & NTT_TMA_STORE) && !sc.enabled) {
    // generic-proxy writes of the butterflies -> visible to the async proxy, then one thread hands the tile to TMA
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int t = 0; t < n_rows; t += box_rows) tma_store_3d(&tmap, (int)col0, (int)row_lo, row_hi0 + t, tile + 2 * t);
      tma_store_commit_wait();
    }
    return;
  }
  if (sc.enabled && sc.tma_box_rows > 0) {
    // fused exchange through TMA: rows j of the tile with the same (row div 2^l0) >> (log_run - l0) go to one peer, to
    // consecutive positions of its (row div 2^l0) dimension: local row (rank << log_run) | (row mod 2^log_run)
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int sh = sc.log_run - l0;
      for (int t = 0; t < n_rows; t += sc.tma_box_rows) {
        const int row_hi = row_hi0 + t;
        tma_store_3d(&peer_maps.m[row_hi >> sh], (int)col0, (int)row_lo, (sc.rank << sh) | (row_hi & ((1 << sh) - 1)), tile + 2 * t);
      }
      tma_store_commit_wait();
    }
    return;
  }
  for (int item = threadIdx.x; item < 2 * n_rows; item += blockDim.x) {
    const int j = item >> 1, half = item & 1;
    if (half == 1 && !two_halves) continue;
    const uint64_t row = row_base + ((uint64_t)j << l0);
    if (sc.enabled) {
      const uint64_t lrow = ((uint64_t)sc.rank << sc.log_run) | (row & (((uint64_t)1 << sc.log_run) - 1));
      *reinterpret_cast<uint4*>(sc.dst[row >> sc.log_run] + lrow * w + col0 + 4 * half) = tile[tile_slot(j, half)];
    } else {
      *reinterpret_cast<uint4*>(mat + row * w + col0 + 4 * half) = tile[tile_slot(j, half)];
    }
  }
}

// ---- the scattering pass on WIDE tiles ----------------------------------------------------------------------------------
// tools/microbench/peer_store_probe.cu (profiles/r02_peer_store_probe.txt): with both directions of the fabric busy, TMA stores of
// 32-byte row pieces (8 columns) reach a peer at 290 GB/s, of 128-byte pieces (32 columns) at 663 GB/s.  The last pass of the
// sharded commit's local transform therefore runs on tiles of COLS = 16 or 32 columns x 2^L rows (64 KiB: L <= 10 / 9), loaded
// and stored by TMA only; everything else (twiddles, radix-8 groups in registers) is the 8-column pass.
template <int G, int H>
__device__ __forceinline__ void tile_group_wide(uint4* tile, const uint32_t* tw_s, int L, int lp) {
  constexpr int Q = 1 << G;
  const int n_items = (1 << (L - G)) * H;
  for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
    const int h = item % H, u = item / H;
    const int j_lo = u & ((1 << lp) - 1);
    const int j0 = ((u >> lp) << (lp + G)) | j_lo;
    uint4 v[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = tile[(j0 + (q << lp)) * H + h];
#pragma unroll
    for (int s = 0; s < G; s++) {
      const uint32_t* tw_l = tw_s + (1 << (lp + s));
#pragma unroll
      for (int q = 0; q < Q; q++) {
        if (q & (1 << s)) continue;
        const uint32_t t = tw_l[j_lo + ((q & ((1 << s) - 1)) << lp)];
        bfly4(v[q], v[q | (1 << s)], t);
      }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) tile[(j0 + (q << lp)) * H + h] = v[q];
  }
}
template <int COLS>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS)
ntt_pass_wide_kernel(int log_h, int l0, int L, const uint32_t* __restrict__ tw, int tw_shift, uint32_t tile0, uint32_t n_col_tiles,
                     const uint32_t* __restrict__ tw_pass, const NttScatter sc, const __grid_constant__ CUtensorMap tmap,
                     const __grid_constant__ NttPeerMaps peer_maps) {
  constexpr int H = COLS / 4;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* tile = reinterpret_cast<uint4*>(smem_raw);  // [2^L][H]
  uint32_t* tw_s = reinterpret_cast<uint32_t*>(tile + ((size_t)H << L));
  uint64_t& tma_bar = *reinterpret_cast<uint64_t*>(tw_s + ((size_t)1 << L));
  const uint64_t col0 = (uint64_t)(tile0 + blockIdx.x % n_col_tiles) * COLS;
  const uint64_t grp = blockIdx.x / n_col_tiles;
  const uint64_t row_lo = grp & (((uint64_t)1 << l0) - 1);
  const int n_rows = 1 << L;
  const int box_rows = n_rows < NTT_BOX_ROWS ? n_rows : NTT_BOX_ROWS;
  const int row_hi0 = (int)((grp >> l0) << L);
  if (threadIdx.x == 0) mbar_init(&tma_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&tma_bar, (uint32_t)n_rows * COLS * 4u);
    for (int t = 0; t < n_rows; t += box_rows) tma_load_3d(tile + (size_t)t * H, &tmap, (int)col0, (int)row_lo, row_hi0 + t, &tma_bar);
  }
  for (int k = threadIdx.x + 1; k < n_rows; k += blockDim.x) {
    uint32_t t = __ldg(tw_pass + k);
    if (row_lo) {
      const int ll = 31 - __clz(k);
      t = kb_mul(t, __ldg(tw + ((row_lo << (log_h - (l0 + ll) - 1)) << tw_shift)));
    }
    tw_s[k] = t;
  }
  mbar_wait(&tma_bar, 0);
  __syncthreads();
  int lp = 0;
  while (lp < L) {
    const int rem = L - lp;
    const int g = rem <= 4 ? (rem == 4 ? 2 : rem) : 3;
    if (g == 3)
      tile_group_wide<3, H>(tile, tw_s, L, lp);
    else if (g == 2)
      tile_group_wide<2, H>(tile, tw_s, L, lp);
    else
      tile_group_wide<1, H>(tile, tw_s, L, lp);
    lp += g;
    __syncthreads();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sc.enabled) {
      const int sh = sc.log_run - l0;
      for (int t = 0; t < n_rows; t += sc.tma_box_rows) {
        const int row_hi = row_hi0 + t;
        tma_store_3d(&peer_maps.m[row_hi >> sh], (int)col0, (int)row_lo, (sc.rank << sh) | (row_hi & ((1 << sh) - 1)), tile + (size_t)t * H);
      }
    } else {
      for (int t = 0; t < n_rows; t += box_rows) tma_store_3d(&tmap, (int)col0, (int)row_lo, row_hi0 + t, tile + (size_t)t * H);
    }
    tma_store_commit_wait();
  }
}

// Generic single-layer kernel for widths that are not a multiple of 4 (not on the benchmark path).
__global__ void ntt_layer_kernel(uint32_t* __restrict__ mat, uint64_t h, uint64_t w, int l, int log_h,
                                 const uint32_t* __restrict__ tw, int tw_shift) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (h / 2) * w) return;
  const uint64_t pair = idx / w, c = idx % w;
  const uint64_t m = (uint64_t)1 << l;
  const uint64_t blk = pair >> l, i = pair & (m - 1);
  uint32_t* lo = mat + (blk * 2 * m + i) * w + c;
  uint32_t* hi = lo + m * w;
  const uint32_t t = __ldg(tw + ((i << (log_h - l - 1)) << tw_shift));
  uint32_t a = *lo, b = *hi;
  bfly(a, b, t);
  *lo = a;
  *hi = b;
}

// Layers [l_first, log_h) on a locally held subset of rows (row-sharded multi-GPU commit):
// local row (m, j') <-> global row m * block + offset + j'; the partner of a row differs only in m.
__global__ void ntt_layer_mapped_kernel(uint32_t* __restrict__ mat, uint64_t w4, int log_h, int l, uint64_t n_blocks,
                                        uint64_t run, uint64_t block, uint64_t offset, const uint32_t* __restrict__ tw,
                                        int tw_shift) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = (n_blocks / 2) * run * w4;
  if (idx >= total) return;
  const uint64_t c4 = idx % w4, pr = idx / w4;
  const uint64_t pair = pr / run, jp = pr % run;
  const uint64_t m_stride = ((uint64_t)1 << l) / block;
  const uint64_t m_lo = (pair / m_stride) * 2 * m_stride + pair % m_stride;
  const uint64_t m_hi = m_lo + m_stride;
  const uint64_t grow = m_lo * block + offset + jp;
  const uint64_t e = (grow & (((uint64_t)1 << l) - 1)) << (log_h - l - 1);
  const uint32_t t = __ldg(tw + (e << tw_shift));
  uint4* lo = reinterpret_cast<uint4*>(mat + (m_lo * run + jp) * (w4 * 4)) + c4;
  uint4* hi = reinterpret_cast<uint4*>(mat + (m_hi * run + jp) * (w4 * 4)) + c4;
  uint4 a = *lo, b = *hi;
  bfly4(a, b, t);
  *lo = a;
  *hi = b;
}

// All log2(NB) mapped layers in one pass: a thread holds the NB rows (one per block m) of its (j', 16-byte column group)
// in registers and runs the radix-NB butterfly network on them — one read and one write of the local matrix instead of
// one per layer.  Requires 2^l_first == block (the row-sharded commit's case): layer l_first + s pairs m and m + 2^s.
template <int NB_LOG>
__global__ void __launch_bounds__(256)
ntt_layers_mapped_fused_kernel(const uint32_t* mat, uint32_t* out, uint64_t w4, int log_h, int l_first, uint64_t run, uint64_t block,
                               uint64_t offset, const uint32_t* __restrict__ tw, int tw_shift, uint64_t c4_begin, uint64_t c4_count) {
  constexpr int NB = 1 << NB_LOG;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= run * c4_count) return;
  const uint64_t c4 = c4_begin + idx % c4_count, jp = idx / c4_count;
  uint4 v[NB];
#pragma unroll
  for (int m = 0; m < NB; m++) v[m] = *(reinterpret_cast<const uint4*>(mat + ((uint64_t)m * run + jp) * (w4 * 4)) + c4);
#pragma unroll
  for (int s = 0; s < NB_LOG; s++) {
    const int l = l_first + s;
#pragma unroll
    for (int m = 0; m < NB; m++) {
      if (m & (1 << s)) continue;
      const uint64_t grow = (uint64_t)m * block + offset + jp;  // global row of the low element
      const uint64_t e = (grow & (((uint64_t)1 << l) - 1)) << (log_h - l - 1);
      bfly4(v[m], v[m | (1 << s)], __ldg(tw + (e << tw_shift)));
    }
  }
#pragma unroll
  for (int m = 0; m < NB; m++) *(reinterpret_cast<uint4*>(out + ((uint64_t)m * run + jp) * (w4 * 4)) + c4) = v[m];
}

cudaError_t ntt_layers_mapped(cudaStream_t stream, uint32_t* d_mat, uint64_t w, unsigned log_h, unsigned l_first,
                              uint64_t n_blocks, uint64_t run, uint64_t block, uint64_t offset, const uint32_t* d_tw,
                              unsigned tw_log_n, uint64_t col_begin, uint64_t col_count, uint32_t* d_out) {
  if (!d_out) d_out = d_mat;  // in place
  if (w % 4 != 0 || log_h > tw_log_n || l_first > log_h || n_blocks < 2 || (((uint64_t)1 << l_first) < block))
    return cudaErrorInvalidValue;
  if (col_count == 0) col_begin = 0, col_count = w;
  if (col_begin % 4 || col_count % 4 || col_begin + col_count > w) return cudaErrorInvalidValue;
  const int tw_shift = (int)tw_log_n - (int)log_h;
  if ((((uint64_t)1 << l_first) == block) && n_blocks == ((uint64_t)1 << (log_h - l_first)) && n_blocks <= 8) {
    const uint64_t c4b = col_begin / 4, c4n = col_count / 4;
    const uint64_t items = run * c4n;
    const unsigned grid = (unsigned)((items + 255) / 256);
    if (n_blocks == 2)
      ntt_layers_mapped_fused_kernel<1><<<grid, 256, 0, stream>>>(d_mat, d_out, w / 4, (int)log_h, (int)l_first, run, block, offset, d_tw, tw_shift, c4b, c4n);
    else if (n_blocks == 4)
      ntt_layers_mapped_fused_kernel<2><<<grid, 256, 0, stream>>>(d_mat, d_out, w / 4, (int)log_h, (int)l_first, run, block, offset, d_tw, tw_shift, c4b, c4n);
    else
      ntt_layers_mapped_fused_kernel<3><<<grid, 256, 0, stream>>>(d_mat, d_out, w / 4, (int)log_h, (int)l_first, run, block, offset, d_tw, tw_shift, c4b, c4n);
    count_launch();
    return cudaGetLastError();
  }
  if (col_begin != 0 || col_count != w) return cudaErrorInvalidValue;  // the per-layer fallback works on whole rows
  if (d_out != d_mat) {  // per-layer fallback: copy once, then in place on the copy
    cudaError_t e = cudaMemcpyAsync(d_out, d_mat, n_blocks * run * w * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    d_mat = d_out;
  }
  const uint64_t total = (n_blocks / 2) * run * (w / 4);
  for (unsigned l = l_first; l < log_h; l++) {
    ntt_layer_mapped_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_mat, w / 4, (int)log_h, (int)l, n_blocks, run,
                                                                                block, offset, d_tw, tw_shift);
    count_launch();
  }
  return cudaGetLastError();
}

// col_tile0 / n_tiles: restrict the pass kernels to a range of 8-column tiles (n_tiles == 0: all of them)
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links only the static runtime)
typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmaEncodeFn tma_encode_fn() {
  static TmaEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TmaEncodeFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}
// the h x w matrix as the tensor (column, row mod 2^l0, row div 2^l0); box = 8 columns x 1 x min(2^L, 256) rows, dense in shared memory
static bool tma_make_map(CUtensorMap* map, uint32_t* d_mat, uint64_t h, uint64_t w, int l0, int L, int box_rows = 0,
                         int box_cols = TILE_COLS) {
  const cuuint64_t dims[3] = {w, (cuuint64_t)1 << l0, h >> l0};
  const cuuint64_t strides[2] = {w * 4, (w * 4) << l0};  // bytes, dimensions 1 and 2
  if (box_rows <= 0) box_rows = (1 << L) < NTT_BOX_ROWS ? (1 << L) : NTT_BOX_ROWS;
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, 1, (cuuint32_t)box_rows};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_NONE;
  return tma_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d_mat, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static cudaError_t run_layers(cudaStream_t stream, uint32_t* d_mat, const uint32_t* d_src, uint32_t log_block,
                              uint32_t r, uint64_t h, uint64_t w, int skip, const uint32_t* d_tw, unsigned tw_log_n,
                              uint32_t col_tile0 = 0, uint32_t n_tiles = 0, const NttScatter* scatter = nullptr) {
  int log_h = 0;
  while (((uint64_t)1 << log_h) < h) log_h++;
  if (((uint64_t)1 << log_h) != h || (unsigned)log_h > tw_log_n) return cudaErrorInvalidValue;
  const int tw_shift = (int)tw_log_n - log_h;  // w_h^e = w_N^(e << tw_shift)
  if (log_h == 0) return cudaSuccess;

  if (w % 4 != 0) {
    if (d_src != nullptr || scatter) return cudaErrorInvalidValue;
    for (int l = skip; l < log_h; l++) {
      const uint64_t n = (h / 2) * w;
      ntt_layer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_mat, h, w, l, log_h, d_tw, tw_shift); count_launch();
    }
    return cudaGetLastError();
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ntt_pass_kernel<NTT_THREADS, NTT_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << MAX_TILE_LOG) * 36 + 16);
    cudaFuncSetAttribute(ntt_pass_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << BIG_TILE_LOG) * 36 + 16);
    attr_set = true;
  }
#if defined(NTT_SWIZZLE_XOR)
  const bool tma_allowed = false;  // the XOR layout is not a TMA box
#else
  const bool tma_allowed = getenv("LM_NTT_NO_TMA") == nullptr;
#endif
  const bool tma_ok = tma_allowed && tma_encode_fn() != nullptr && (reinterpret_cast<uintptr_t>(d_mat) & 15) == 0 &&
                      h < ((uint64_t)1 << 31) && w < ((uint64_t)1 << 31);
  // split log_h layers into ceil(log_h / 11) passes of nearly equal depth
  static const bool big_tiles = getenv("LM_NTT_NO_BIG_TILES") == nullptr;
  const int max_tile_log = (big_tiles && log_h > 2 * MAX_TILE_LOG && log_h <= 2 * BIG_TILE_LOG) ? BIG_TILE_LOG : MAX_TILE_LOG;
  const int n_pass = (log_h + max_tile_log - 1) / max_tile_log;
  // the scattering pass of a two-pass transform prefers 9 layers: its tile can then be 32 columns wide (128-byte row pieces over NVLink)
  static const bool wide_allowed = getenv("LM_NTT_NO_WIDE_SCATTER") == nullptr;
  const uint32_t tiles8 = n_tiles ? n_tiles : (uint32_t)((w + TILE_COLS - 1) / TILE_COLS);
  const bool wide_possible = scatter && wide_allowed && tma_ok && n_pass == 2 && getenv("LM_NTT_NO_TMA_SCATTER") == nullptr &&
                             w % 16 == 0 && (col_tile0 * TILE_COLS) % 16 == 0 && (tiles8 * TILE_COLS) % 16 == 0;
  const int last_L_pref = (wide_possible && log_h - 9 <= MAX_TILE_LOG && w % 32 == 0 && (col_tile0 * TILE_COLS) % 32 == 0 &&
                           (tiles8 * TILE_COLS) % 32 == 0 && log_h >= 10) ? 9 : 0;
  int l0 = 0, skip_left = skip;
  for (int p = 0; p < n_pass; p++) {
    int L = (log_h - l0 + (n_pass - p) - 1) / (n_pass - p);
    if (last_L_pref && n_pass == 2) L = p == 0 ? log_h - last_L_pref : last_L_pref;
    const int sk = skip_left < L ? skip_left : L;
    skip_left -= sk;
    const uint32_t tiles = n_tiles ? n_tiles : (uint32_t)((w + TILE_COLS - 1) / TILE_COLS);
    const uint64_t n_cta = (uint64_t)tiles * (h >> L);
    const size_t smem = ((size_t)1 << L) * 36 + 16;  // tile + per-CTA twiddles + mbarrier
    // compact twiddles of this pass, in the scratch words behind the big table (ntt.h: NTT_TW_SCRATCH_WORDS)
    uint32_t* tw_pass = const_cast<uint32_t*>(d_tw) + ((size_t)1 << (tw_log_n - 1)) + (size_t)(p % 4) * ((size_t)1 << BIG_TILE_LOG);
    ntt_pass_twiddles_kernel<<<((1 << L) + 255) / 256, 256, 0, stream>>>(tw_pass, log_h, l0, L, d_tw, tw_shift); count_launch();
    NttScatter sc{};
    NttPeerMaps peer_maps;  // only read by the scattering pass (2 KiB of kernel parameters; per call: rank threads share the process)
    memset(&peer_maps.m[0], 0, sizeof(CUtensorMap));
    if (scatter && p == n_pass - 1) {
      sc = *scatter;
      // TMA stores into the peers' matrices when a run holds at least 8 rows of the (row div 2^l0) dimension
      static const bool tma_scatter_allowed = getenv("LM_NTT_NO_TMA_SCATTER") == nullptr;
      const int sh = sc.log_run - l0;
      if (tma_ok && tma_scatter_allowed && sh >= 3) {
        int box = 1 << (sh < L ? sh : L);
        if (box > NTT_BOX_ROWS) box = NTT_BOX_ROWS;
        bool ok = true;
        for (int q = 0; q < NTT_MAX_PEERS && sc.dst[q] && ok; q++)
          ok = (reinterpret_cast<uintptr_t>(sc.dst[q]) & 15) == 0 && tma_make_map(&peer_maps.m[q], sc.dst[q], h, w, l0, L, box);
        // wide tile for this pass: 32 columns when it has <= 9 layers, 16 when it has 10
        int wide_cols = 0;
        if (wide_possible && p > 0 && sk == 0) wide_cols = (L <= 9 && last_L_pref) ? 32 : (L <= 10 ? 16 : 0);
        if (ok && wide_cols) {
          for (int q = 0; q < NTT_MAX_PEERS && sc.dst[q] && ok; q++) ok = tma_make_map(&peer_maps.m[q], sc.dst[q], h, w, l0, L, box, wide_cols);
          CUtensorMap wmap;
          if (ok && tma_make_map(&wmap, d_mat, h, w, l0, L, 0, wide_cols)) {
            sc.tma_box_rows = box;
            const uint32_t wtiles = tiles8 * TILE_COLS / wide_cols, wtile0 = col_tile0 * TILE_COLS / wide_cols;
            const size_t wsmem = ((size_t)1 << L) * (wide_cols * 4 + 4) + 16;
            static bool wattr = false;
            if (!wattr) {
              cudaFuncSetAttribute(ntt_pass_wide_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 10) * 68 + 16);
              cudaFuncSetAttribute(ntt_pass_wide_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << 9) * 132 + 16);
              wattr = true;
            }
            const unsigned wg = (unsigned)((uint64_t)wtiles * (h >> L));
            if (wide_cols == 32)
              ntt_pass_wide_kernel<32><<<wg, NTT_THREADS, wsmem, stream>>>(log_h, l0, L, d_tw, tw_shift, wtile0, wtiles, tw_pass, sc, wmap, peer_maps);
            else
              ntt_pass_wide_kernel<16><<<wg, NTT_THREADS, wsmem, stream>>>(log_h, l0, L, d_tw, tw_shift, wtile0, wtiles, tw_pass, sc, wmap, peer_maps);
            count_launch();
            static const bool dbgw = getenv("LM_NTT_DEBUG") != nullptr;
            if (dbgw) fprintf(stderr, "[lm ntt] scatter pass l0=%d L=%d on %d-column tiles (box %d rows)\n", l0, L, wide_cols, box);
            cudaError_t ew = cudaGetLastError();
            if (ew != cudaSuccess) return ew;
            l0 += L;
            continue;
          }
          ok = false;  // fall back to the 8-column pass with SM stores (the peer maps now describe wide boxes)
        }
        if (ok) sc.tma_box_rows = box;
        static const bool dbg = getenv("LM_NTT_DEBUG") != nullptr;
        if (dbg) fprintf(stderr, "[lm ntt] scatter pass l0=%d L=%d log_run=%d: %s (box %d rows)\n", l0, L, sc.log_run, ok ? "TMA stores into the peer matrices" : "tensor map of a peer matrix refused, SM stores", box);
      }
    }
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int tma = 0;
    if (tma_ok && tma_make_map(&tmap, d_mat, h, w, l0, L)) tma = NTT_TMA_LOAD | NTT_TMA_STORE;
    if (L > MAX_TILE_LOG)
      ntt_pass_kernel<1024, 1><<<(unsigned)n_cta, 1024, smem, stream>>>(d_mat, p == 0 ? d_src : nullptr, w, log_h, l0, L, sk, log_block, r,
                                                                         d_tw, tw_shift, col_tile0, tiles, tw_pass, sc, tmap, tma, peer_maps);
    else
      ntt_pass_kernel<NTT_THREADS, NTT_MIN_BLOCKS><<<(unsigned)n_cta, NTT_THREADS, smem, stream>>>(
          d_mat, p == 0 ? d_src : nullptr, w, log_h, l0, L, sk, log_block, r, d_tw, tw_shift, col_tile0, tiles, tw_pass, sc, tmap, tma, peer_maps);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    l0 += L;
  }
  return cudaSuccess;
}

// gather + DFT of columns [col_begin, col_begin + col_count) only (both multiples of 8; dim = 1, width % 4 == 0):
// used to overlap the host-to-device copy of later columns with the transform of earlier ones
cudaError_t ntt_reorder_and_dft_cols(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding_factor,
                                     uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t col_begin, uint32_t col_count,
                                     uint32_t* d_out, const uint32_t* d_tw, unsigned tw_log_n) {
  if (col_begin % 8 || col_count % 8 || col_count == 0 || col_begin + col_count > dft_n_cols || dft_n_cols % 4)
    return cudaErrorInvalidValue;
  const uint32_t log_block = n_vars + log_inv_rate - folding_factor;
  const uint64_t h = (uint64_t)1 << log_block;
  const int skip = (int)(log_inv_rate < log_block ? log_inv_rate : log_block);
  return run_layers(stream, d_out, d_evals, log_block, log_inv_rate, h, dft_n_cols, skip, d_tw, tw_log_n, col_begin / 8,
                    col_count / 8);
}

// gather + DFT of a rank's shard with the exchange of the row-sharded commit fused into the last pass: d_work is the
// rank's own block x w scratch, peers[q] the matrix of rank q (q == rank: its own), each block x w words
cudaError_t ntt_reorder_and_dft_scatter(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding_factor,
                                        uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_work, uint32_t* const* peers,
                                        uint32_t world, uint32_t rank, const uint32_t* d_tw, unsigned tw_log_n,
                                        uint32_t col_begin, uint32_t col_count) {
  if (folding_factor > n_vars + log_inv_rate || dft_n_cols % 4 || dft_n_cols == 0 || world < 2 || world > NTT_MAX_PEERS ||
      (world & (world - 1)) || rank >= world)
    return cudaErrorInvalidValue;
  if (col_count == 0) col_begin = 0, col_count = dft_n_cols;
  if (col_begin % 8 || (col_count % 8 && col_begin + col_count != dft_n_cols) || col_begin + col_count > dft_n_cols)
    return cudaErrorInvalidValue;
  const uint32_t log_block = n_vars + log_inv_rate - folding_factor;
  uint32_t g = 0;
  while ((1u << g) < world) g++;
  if (log_block < 2 * g) return cudaErrorInvalidValue;
  const uint64_t h = (uint64_t)1 << log_block;
  const int skip = (int)(log_inv_rate < log_block ? log_inv_rate : log_block);
  NttScatter sc{};
  for (uint32_t q = 0; q < world; q++) sc.dst[q] = peers[q];
  sc.log_run = (int)(log_block - g);
  sc.rank = (int)rank;
  sc.enabled = 1;
  return run_layers(stream, d_work, d_evals, log_block, log_inv_rate, h, dft_n_cols, skip, d_tw, tw_log_n, col_begin / 8,
                    (col_count + 7) / 8, &sc);
}

cudaError_t ntt_dft_batch_by_evals(cudaStream_t stream, uint32_t* d_mat, uint64_t h, uint64_t w, int skip_layers,
                                   const uint32_t* d_tw, unsigned tw_log_n) {
  return run_layers(stream, d_mat, nullptr, 0, 0, h, w, skip_layers, d_tw, tw_log_n);
}

cudaError_t ntt_reorder_and_dft(cudaStream_t stream, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim,
                                uint32_t folding_factor, uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_out,
                                const uint32_t* d_tw, unsigned tw_log_n) {
  if (folding_factor > n_vars + log_inv_rate) return cudaErrorInvalidValue;
  const uint32_t log_block = n_vars + log_inv_rate - folding_factor;
  const uint64_t h = (uint64_t)1 << log_block;
  const uint64_t w = (uint64_t)dft_n_cols * dim;
  if (dft_n_cols == 0) return cudaSuccess;
  const int skip = (int)(log_inv_rate < log_block ? log_inv_rate : log_block);
  if (dim == 1 && w % 4 == 0) return run_layers(stream, d_out, d_evals, log_block, log_inv_rate, h, w, skip, d_tw, tw_log_n);
  const uint64_t n = h * w;
  gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_evals, d_out, h, dft_n_cols, dim, log_block, log_inv_rate); count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return run_layers(stream, d_out, nullptr, 0, 0, h, w, skip, d_tw, tw_log_n);
}

}  // namespace lm

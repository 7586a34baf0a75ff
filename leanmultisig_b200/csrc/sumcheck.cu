// WHIR-open kernels on sm_100a: statement weights, product-sumcheck rounds, STIR equality updates.
//
// Device replacement for
//   crates/whir/src/open.rs:518-584                        combine_statement (eq / next weights with selectors)
//   crates/backend/poly/src/next_mle.rs:35-58              matrix_next_mle_folded
//   crates/whir/src/open.rs:337-382                        add_new_equality / add_new_base_equality
//   crates/backend/poly/src/eq_mle.rs:372-430              compute_eval_eq_base_packed_batched
//   crates/backend/sumcheck/src/product_computation.rs:127-170   round polynomial (c0, c2)
//   crates/backend/sumcheck/src/product_computation.rs:242-304   fold with the previous challenge + next round
// Tables are AoS EF (5 words) exactly like the reference's Vec<EF>; the polynomial may still be base field in
// the first round.  MSB-first folding: t'[i] = t[i] + r (t[i + n/2] - t[i]).
//
// All kernels stream their tables once (HBM bound: 24 B per index in round 0, 40 B afterwards) with delayed
// modular reduction in 64-bit accumulators; sums are reduced warp -> CTA -> one partial per CTA -> final kernel.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "launch_count.h"
#include "kb.cuh"
#include "reduce.cuh"
#include "umma.cuh"
#include "poly.h"
#include "sumcheck.h"

namespace lm {

constexpr int SPLIT_LO = 10;  // eq(point, x) = hi[x >> 10] * lo[x & 1023]
constexpr int WG_HI_VARS = 10;  // tensor-core path: eq(point, x) = hi[x >> (m - 10)] * lo[x & (2^(m-10) - 1)], 1024 constants

// ---------------------------------------------------------------------------------------------- weights
// w[base + x] += hi[x >> lo_vars] * lo[x & mask]
__global__ void weights_add_split_kernel(uint32_t* __restrict__ w, uint64_t base, uint64_t n, int lo_vars,
                                         const uint32_t* __restrict__ hi, const uint32_t* __restrict__ lo) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const Ef h = ld_ef(hi + 5 * (x >> lo_vars));
  const Ef l = ld_ef(lo + 5 * (x & (((uint64_t)1 << lo_vars) - 1)));
  uint32_t* dst = w + 5 * (base + x);
  st_ef(dst, ef_add(ld_ef_rw(dst), ef_mul(h, l)));
}

cudaError_t weights_add_eq(cudaStream_t stream, uint32_t* d_w, uint64_t selector, const uint32_t* d_point, uint32_t m,
                           const uint32_t scalar[5], uint32_t* d_scratch) {
  const int lo_vars = m < (uint32_t)SPLIT_LO ? (int)m : SPLIT_LO;
  const int hi_vars = (int)m - lo_vars;
  uint32_t* d_hi = d_scratch;
  uint32_t* d_lo = d_hi + 5 * ((uint64_t)1 << hi_vars);
  const uint32_t one[5] = {KB_R1, 0, 0, 0, 0};
  cudaError_t e;
  if ((e = eq_table(stream, d_point, hi_vars, scalar, d_hi)) != cudaSuccess) return e;
  if ((e = eq_table(stream, d_point + 5 * hi_vars, lo_vars, one, d_lo)) != cudaSuccess) return e;
  const uint64_t n = (uint64_t)1 << m;
  weights_add_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_w, selector << m, n, lo_vars, d_hi, d_lo);
  count_launch();
  return cudaGetLastError();
}
// K statements with the same selector and point length in ONE pass over the weights (combine_statement, open.rs:518-584,
// adds them one after the other; at 2^27 entries each separate pass is a 5.4 GB read-modify-write):
// w[base + x] += sum_k hi_k[x >> lo_vars] * lo_k[x & mask], one reduction per coefficient for all K products.
// hi / lo: K tables one after the other.
template <int KMAX>
__global__ void __launch_bounds__(256)
weights_add_split_batch_kernel(uint32_t* __restrict__ w, uint64_t base, uint64_t n, int lo_vars, const uint32_t* __restrict__ hi,
                               const uint32_t* __restrict__ lo, int K) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const uint64_t n_hi = n >> lo_vars, n_lo = (uint64_t)1 << lo_vars;
  const uint64_t xh = x >> lo_vars, xl = x & (n_lo - 1);
  uint64_t acc[5] = {0, 0, 0, 0, 0};
  int pending = 0;
  for (int k = 0; k < K; k++) {
    const Ef a = ld_ef(hi + 5 * (k * n_hi + xh));
    const EfRows b = ef_rows(ld_ef(lo + 5 * (k * n_lo + xl)));
    const uint32_t rows[5][5] = {{b.b0, b.b4, b.b3, b.b2, b.b1m4},
                                 {b.b1, b.b0, b.b4, b.b3, b.b2},
                                 {b.b2, b.b1m4, b.b0m3, b.b4m2, b.b3m14},
                                 {b.b3, b.b2, b.b1m4, b.b0m3, b.b4m2},
                                 {b.b4, b.b3, b.b2, b.b1m4, b.b0m3}};
#pragma unroll
    for (int t = 0; t < 5; t++) {
      if (pending == 4) {
#pragma unroll
        for (int i = 0; i < 5; i++) acc[i] = kb_fold(acc[i]);
        pending = 0;
      }
      pending++;
#pragma unroll
      for (int i = 0; i < 5; i++) acc[i] = mad_wide(a.c[t], rows[i][t], acc[i]);
    }
  }
  uint32_t* dst = w + 5 * (base + x);
  Ef cur = ld_ef_rw(dst);
#pragma unroll
  for (int i = 0; i < 5; i++) cur.c[i] = kb_add(cur.c[i], kb_canon(kb_redc_lazy(kb_fold(acc[i]))));
  st_ef(dst, cur);
}


// ---- the same batch as ONE integer GEMM on the tensor cores --------------------------------------------------------------
// Split the index as x = (x_hi, x_lo) with only 10 HIGH variables: for a fixed x_hi the K values hi_k[x_hi] are constants shared
// by every x_lo, so  out[x_lo][x_hi] = sum_k lo_k[x_lo] * hi_k[x_hi]  is the product of the (n_lo x 5K) matrix of lo coefficients
// with a (5K x 5 * 1024) matrix built from the hi tables (entry ((k, t), (x_hi, i)) = row matrix of ef_mul: coefficient i of
// the product takes a_t * rows[i][t]).  Rows = x_lo = consecutive table entries, so a warp's accesses to w are contiguous.  Evaluated like the
// Poseidon1 products (umma.cuh, poseidon1_umma.cuh): the words of A are their own u8 limbs, limb l of an input meets the
// constant pre-shifted mod p (c 2^(8 l) mod p) split into four byte columns, so one output coefficient is four s32 accumulator
// columns, recombined on the ALU pipe and reduced ONCE — 25 K multiply-accumulates with folds per table entry become 5
// reductions.  M = 128 values of x_lo per CTA (thread r = row r = TMEM lane r), N = 80 columns = 4 values of x_hi per tile,
// 256 tiles split over WG_SPLIT CTAs; the B image (80 columns x 20K bytes per tile, shared-memory layout) is built once per
// call by weights_gemm_image_kernel and streamed from L2 with cp.async, double buffered.
constexpr int WG_TILE_XLO = 4;                      // values of the 10-variable index per N tile
constexpr int WG_TILE_N = WG_TILE_XLO * 5 * 4;      // 80 accumulator columns
constexpr int WG_TMEM_COLS = 128;
// 12 statements = 240 input bytes per row: every accumulator column stays below 240 * 255^2 < 2^24 and the carry-free low word
// T0 + 2^8 T1 below 2^32 for ANY inputs (13 would not); larger batches are split by the caller (weights_add_eq_batch_max)
constexpr int WG_KMAX = 12;

__host__ __device__ inline uint32_t wg_kbytes(int K) { return (uint32_t)((20 * K + 31) / 32 * 32); }

// image bytes of one (x, k, t): the five constants rows[i][t] of ef_mul for b = tab_k[x], pre-shifted for the four limbs of a_t
// (shared by the device kernel and the CPU model below)
LM_HD void wg_image_fill(const Ef& bval, uint32_t x, uint32_t k, uint32_t t, int K, uint8_t* img) {
  const EfRows b = ef_rows(bval);
  const uint32_t rows[5][5] = {{b.b0, b.b4, b.b3, b.b2, b.b1m4},
                               {b.b1, b.b0, b.b4, b.b3, b.b2},
                               {b.b2, b.b1m4, b.b0m3, b.b4m2, b.b3m14},
                               {b.b3, b.b2, b.b1m4, b.b0m3, b.b4m2},
                               {b.b4, b.b3, b.b2, b.b1m4, b.b0m3}};
  const uint32_t kb = wg_kbytes(K), kchunks = kb / 16;
  uint8_t* tile = img + (size_t)(x / WG_TILE_XLO) * WG_TILE_N * kb;
  const uint32_t q = 5 * k + t;  // word index of a_t of statement k in the A row
  for (int i = 0; i < 5; i++) {
    uint64_t m = rows[i][t];
    for (int l = 0; l < 4; l++) {
      const uint32_t kbyte = 4 * q + l;
      for (int j = 0; j < 4; j++) {
        const uint32_t n = ((x % WG_TILE_XLO) * 5 + i) * 4 + j;
        tile[(n / 8) * (kchunks * 128) + (kbyte / 16) * 128 + (n % 8) * 16 + (kbyte % 16)] = (uint8_t)(m >> (8 * j));
      }
      m = (m << 8) % KB_P;
    }
  }
}
// one thread per (x, k, t) of the 2^vars-entry tables tab_k (K tables one after the other)
__global__ void weights_gemm_image_kernel(const uint32_t* __restrict__ tab, int K, int vars, uint8_t* __restrict__ img) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = 1u << vars;
  if (idx >= n * (uint32_t)K * 5) return;
  const uint32_t t = idx % 5, k = (idx / 5) % K, x = idx / (5 * K);
  wg_image_fill(ld_ef(tab + 5 * ((size_t)k * n + x)), x, k, t, K, img);
}

// CPU model of weights_gemm_kernel (test hook of the CPU tier, lm_host_eq_gemm_model): the same image builder, the same row
// layout and recombination, the MMAs as the integer dot products they stand for.
//   w[(x_hi << lo_vars) + x_lo] += sum_k hi_k[x_hi] * lo_k[x_lo];   hi: K x 2^hi_vars EF, lo: K x 2^lo_vars EF, hi_vars >= 2
void weights_gemm_model_host(uint32_t* w, const uint32_t* hi, const uint32_t* lo, int K, int hi_vars, int lo_vars) {
  const uint32_t n_hi = 1u << hi_vars, n_lo = 1u << lo_vars, kb = wg_kbytes(K), kchunks = kb / 16;
  std::vector<uint8_t> img((size_t)n_hi / WG_TILE_XLO * WG_TILE_N * kb, 0), row(kb, 0);
  for (uint32_t x = 0; x < n_hi; x++)
    for (int k = 0; k < K; k++)
      for (uint32_t t = 0; t < 5; t++) {
        Ef b;
        for (int c = 0; c < 5; c++) b.c[c] = hi[5 * ((size_t)k * n_hi + x) + c];
        wg_image_fill(b, x, (uint32_t)k, t, K, img.data());
      }
  for (uint32_t xr = 0; xr < n_lo; xr++) {
    for (int k = 0; k < K; k++)
      for (int t = 0; t < 5; t++) {
        const uint32_t word = lo[5 * ((size_t)k * n_lo + xr) + t];
        for (int l = 0; l < 4; l++) row[4 * (5 * k + t) + l] = (uint8_t)(word >> (8 * l));
      }
    for (uint32_t xh = 0; xh < n_hi; xh++) {
      const uint8_t* tile = img.data() + (size_t)(xh / WG_TILE_XLO) * WG_TILE_N * kb;
      uint32_t* dst = w + 5 * (((size_t)xh << lo_vars) + xr);
      for (int i = 0; i < 5; i++) {
        uint32_t T[4];
        for (int j = 0; j < 4; j++) {
          const uint32_t n = ((xh % WG_TILE_XLO) * 5 + i) * 4 + j;
          uint32_t sum = 0;
          for (uint32_t kbyte = 0; kbyte < kb; kbyte++)
            sum += (uint32_t)row[kbyte] * tile[(n / 8) * (kchunks * 128) + (kbyte / 16) * 128 + (n % 8) * 16 + (kbyte % 16)];
          T[j] = sum;
        }
        dst[i] = kb_add(dst[i], kb_canon(p1u_combine_redc<0>(T[0], T[1], T[2], T[3], 0u)));
      }
    }
  }
}

// grid = (n_hi / 128) x WG_SPLIT: CTA (rb, sp) takes x_hi in [128 rb, 128 rb + 128) and the tiles t = sp (mod WG_SPLIT)
constexpr int WG_SPLIT = 4;
__global__ void __launch_bounds__(128, 4)
weights_gemm_kernel(uint32_t* __restrict__ w, uint64_t base, uint64_t n_rows, int row_vars, const uint32_t* __restrict__ a_tab, int K,
                    const uint8_t* __restrict__ img) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  const uint32_t kb = wg_kbytes(K), kchunks = kb / 16, k_steps = kb / 32;
  const uint32_t tile_bytes = WG_TILE_N * kb;
  uint8_t* sa = dsm;                                  // 128 rows x kb
  uint8_t* sb = dsm + 128 * kb;                       // two B tiles: 80 x kb each
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 2 * tile_bytes);
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int r = threadIdx.x, warp = r >> 5;
  const uint64_t xr = (uint64_t)blockIdx.x * 128 + r;  // this thread's row: the low index, entry base + (x_hi << row_vars) + xr
  const uint32_t n_tiles = (1u << WG_HI_VARS) / WG_TILE_XLO;
  const uint32_t sb_addr = p1u_smem_u32(sb);
  // B tile `tile` -> buffer `buf`, asynchronously (the image is already in the descriptor's layout)
  auto fetch_b = [&](uint32_t tile, uint32_t buf) {
    const uint8_t* src = img + (size_t)tile * tile_bytes;
    for (uint32_t i = r; i < tile_bytes / 16; i += 128)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb_addr + buf * tile_bytes + 16 * i), "l"(src + 16 * (size_t)i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  uint32_t tile = blockIdx.y;
  if (tile < n_tiles) fetch_b(tile, 0);
  // A row of this thread: the 5K words of lo_k[xr], zero padding up to kb
  const uint32_t a_row = p1u_smem_u32(sa) + (r >> 3) * (kchunks * 128) + (r & 7) * 16;
  for (uint32_t c = 0; c < kchunks; c++) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a_row + c * 128), "r"(0) : "memory");
  for (int k = 0; k < K; k++) {
    const Ef a = ld_ef(a_tab + 5 * ((size_t)k * n_rows + xr));
#pragma unroll
    for (int t = 0; t < 5; t++) {
      const uint32_t q = 5 * k + t;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_row + (q >> 2) * 128 + (q & 3) * 4), "r"(a.c[t]) : "memory");
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p1u_smem_u32(tm_slot)), "r"(WG_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (r == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p1u_smem_u32(bar)), "r"(1) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tm_slot, tmem = tmem_d + ((uint32_t)(warp * 32) << 16);
  const bool issuer_warp = __shfl_sync(0xffffffffu, warp, 0) == 0;
  uint32_t parity = 0, buf = 0;
  uint32_t* wcol = w + 5 * (base + xr);
  const uint64_t hi_stride = (uint64_t)5 << row_vars;  // words between x_hi and x_hi + 1
  // this thread's 4 table entries of a tile (one per x_hi of the tile; the 32 entries of a warp are contiguous in w)
  uint32_t cur[20], nxt[20];
  if (tile < n_tiles) {
#pragma unroll
    for (int e = 0; e < WG_TILE_XLO; e++)
#pragma unroll
      for (int i = 0; i < 5; i++) cur[5 * e + i] = wcol[(uint64_t)(WG_TILE_XLO * tile + e) * hi_stride + i];
  }
  for (; tile < n_tiles; tile += WG_SPLIT, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (issuer_warp) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (p1u_elect_one()) {
        const uint64_t da = p1u_desc(p1u_smem_u32(sa), kchunks * 128), db = p1u_desc(sb_addr + buf * tile_bytes, kchunks * 128);
        for (uint32_t k = 0; k < k_steps; k++)
          p1u_mma(tmem_d, da + (uint64_t)(k * 256 >> 4), db + (uint64_t)(k * 256 >> 4), p1u_idesc(WG_TILE_N), k > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(p1u_smem_u32(bar)) : "memory");
      }
      __syncwarp();
    }
    // while the product runs: the next tile's B (the other buffer was last read by the product we waited for one iteration
    // ago) and the next tile's table entries
    const uint32_t next = tile + WG_SPLIT;
    if (next < n_tiles) {
      fetch_b(next, buf ^ 1);
#pragma unroll
      for (int e = 0; e < WG_TILE_XLO; e++)
#pragma unroll
        for (int i = 0; i < 5; i++) nxt[5 * e + i] = wcol[(uint64_t)(WG_TILE_XLO * next + e) * hi_stride + i];
    }
    for (uint32_t spins = 0; !p1u_try_wait(p1u_smem_u32(bar), parity);)
      if (++spins > (1u << 26)) __trap();
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t* cw = cur;
#pragma unroll
    for (int h = 0; h < 5; h++) {  // 16 columns = 4 output coefficients at a time
      uint32_t v[16];
      p1u_ld16(tmem + 16 * h, v);
      p1u_wait_ld();
#pragma unroll
      for (int o = 0; o < 4; o++) {
        const uint32_t y = kb_canon(p1u_combine_redc<0>(v[4 * o], v[4 * o + 1], v[4 * o + 2], v[4 * o + 3], 0u));
        cw[4 * h + o] = kb_add(cw[4 * h + o], y);
      }
    }
#pragma unroll
    for (int e = 0; e < WG_TILE_XLO; e++)
#pragma unroll
      for (int i = 0; i < 5; i++) wcol[(uint64_t)(WG_TILE_XLO * tile + e) * hi_stride + i] = cur[5 * e + i];
#pragma unroll
    for (int i = 0; i < 20; i++) cur[i] = nxt[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(WG_TMEM_COLS) : "memory");
}

static bool weights_gemm_enabled() {
  static const bool on = [] {
    const char* e = getenv("LM_EQ_GEMM");
    return !(e && atoi(e) == 0);
  }();
  return on;
}
static size_t weights_gemm_image_bytes(uint32_t K) { return ((size_t)1 << WG_HI_VARS) / WG_TILE_XLO * WG_TILE_N * wg_kbytes((int)K); }
static bool weights_gemm_ok(uint32_t m, uint32_t K) {
  return weights_gemm_enabled() && m >= (uint32_t)WG_HI_VARS + 7 && K >= 1 && K <= (uint32_t)WG_KMAX;
}

// d_points: K points of m coordinates (K x m x 5 words), scalars: host, K x 5 words
cudaError_t weights_add_eq_batch(cudaStream_t stream, uint32_t* d_w, uint64_t selector, const uint32_t* d_points, uint32_t m,
                                 const uint32_t* scalars, uint32_t K, uint32_t* d_scratch) {
  const bool gemm = weights_gemm_ok(m, K);
  const int lo_vars = gemm ? (int)m - WG_HI_VARS : (m < (uint32_t)SPLIT_LO ? (int)m : SPLIT_LO);
  const int hi_vars = (int)m - lo_vars;
  const size_t n_hi = (size_t)1 << hi_vars, n_lo = (size_t)1 << lo_vars;
  uint32_t* d_hi = d_scratch;
  uint32_t* d_lo = d_hi + 5 * n_hi * K;
  const uint32_t one[5] = {KB_R1, 0, 0, 0, 0};
  cudaError_t e;
  for (uint32_t k = 0; k < K; k++) {
    const uint32_t* pt = d_points + (size_t)k * m * 5;
    if ((e = eq_table(stream, pt, hi_vars, scalars + 5 * k, d_hi + 5 * n_hi * k)) != cudaSuccess) return e;
    if ((e = eq_table(stream, pt + 5 * hi_vars, lo_vars, one, d_lo + 5 * n_lo * k)) != cudaSuccess) return e;
  }
  const uint64_t n = (uint64_t)1 << m;
  if (gemm) {
    // tensor-core path: B image (from the 1024-entry tables of the high variables) behind the tables, 16-byte aligned
    uint8_t* d_img = reinterpret_cast<uint8_t*>(d_lo + 5 * n_lo * K + ((4 - (5 * (n_hi + n_lo) * K) % 4) % 4));
    const size_t img_bytes = weights_gemm_image_bytes(K);
    if ((e = cudaMemsetAsync(d_img, 0, img_bytes, stream)) != cudaSuccess) return e;
    const unsigned items = (unsigned)(n_hi * K * 5);
    weights_gemm_image_kernel<<<(items + 127) / 128, 128, 0, stream>>>(d_hi, (int)K, hi_vars, d_img);
    count_launch();
    const uint32_t kb = wg_kbytes((int)K);
    const int dyn = (int)((128 + 2 * WG_TILE_N) * kb + 64);
    if ((e = cudaFuncSetAttribute(weights_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn)) != cudaSuccess) return e;
    weights_gemm_kernel<<<dim3((unsigned)(n_lo / 128), WG_SPLIT), 128, dyn, stream>>>(d_w, selector << m, n_lo, lo_vars, d_lo, (int)K, d_img);
    count_launch();
    return cudaGetLastError();
  }
  weights_add_split_batch_kernel<16><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_w, selector << m, n, lo_vars, d_hi, d_lo, (int)K);
  count_launch();
  return cudaGetLastError();
}
// statements one call of weights_add_eq_batch takes for points of m coordinates
uint32_t weights_add_eq_batch_max(uint32_t m) { return weights_gemm_ok(m, 1) ? (uint32_t)WG_KMAX : 16u; }
size_t weights_add_eq_batch_scratch_words(uint32_t m, uint32_t K) {
  const int lo_vars = m < (uint32_t)SPLIT_LO ? (int)m : SPLIT_LO;  // (the GEMM path splits the other way round: same table sizes)
  size_t words = 5 * (size_t)K * (((size_t)1 << (m - lo_vars)) + ((size_t)1 << lo_vars)) + 8;
  if (weights_gemm_ok(m, K)) words += weights_gemm_image_bytes(K) / 4 + 8;
  return words;
}

size_t weights_add_eq_scratch_words(uint32_t m) {
  const int lo_vars = m < (uint32_t)SPLIT_LO ? (int)m : SPLIT_LO;
  return 5 * (((size_t)1 << (m - lo_vars)) + ((size_t)1 << lo_vars)) + 8;
}

// next_mle.rs:35-58, term k: w[base + (b << (k+1)) + (1 << k)] += scalar * (1 - oc[n-k-1]) * prod_{j >= n-k} oc[j]
//                                                               * eq(oc[0 .. n-k-1), b)
__global__ void weights_add_next_kernel(uint32_t* __restrict__ w, uint64_t base, const uint32_t* __restrict__ oc, int n,
                                        int k, Ef scalar) {
  const int pre = n - k - 1;
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ((uint64_t)1 << pre)) return;
  Ef acc = scalar;
  {
    Ef z = ld_ef(oc + 5 * pre);
#pragma unroll
    for (int c = 0; c < 5; c++) z.c[c] = kb_neg(z.c[c]);
    z.c[0] = kb_add(z.c[0], KB_R1);
    acc = ef_mul(acc, z);
  }
  for (int j = n - k; j < n; j++) acc = ef_mul(acc, ld_ef(oc + 5 * j));
  for (int i = 0; i < pre; i++) {
    Ef z = ld_ef(oc + 5 * i);
    if (!((b >> (pre - 1 - i)) & 1)) {
#pragma unroll
      for (int c = 0; c < 5; c++) z.c[c] = kb_neg(z.c[c]);
      z.c[0] = kb_add(z.c[0], KB_R1);
    }
    acc = ef_mul(acc, z);
  }
  uint32_t* dst = w + 5 * (base + (b << (k + 1)) + ((uint64_t)1 << k));
  st_ef(dst, ef_add(ld_ef_rw(dst), acc));
}
__global__ void weights_add_next_last_kernel(uint32_t* __restrict__ w, uint64_t idx, const uint32_t* __restrict__ oc, int n,
                                             Ef scalar) {
  Ef acc = scalar;
  for (int j = 0; j < n; j++) acc = ef_mul(acc, ld_ef(oc + 5 * j));
  uint32_t* dst = w + 5 * idx;
  st_ef(dst, ef_add(ld_ef_rw(dst), acc));
}

cudaError_t weights_add_next(cudaStream_t stream, uint32_t* d_w, uint64_t selector, const uint32_t* d_point, uint32_t m,
                             const uint32_t scalar[5]) {
  Ef s;
  for (int c = 0; c < 5; c++) s.c[c] = scalar[c];
  const uint64_t base = selector << m;
  for (int k = 0; k < (int)m; k++) {
    const uint64_t n = (uint64_t)1 << (m - k - 1);
    weights_add_next_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_w, base, d_point, (int)m, k, s);
    count_launch();
  }
  weights_add_next_last_kernel<<<1, 1, 0, stream>>>(d_w, base + (((uint64_t)1 << m) - 1), d_point, (int)m, s);
  count_launch();
  return cudaGetLastError();
}

// w[base + (b << shift) + offset] += scalar * eq(pt[0 .. pre), b), b < 2^pre: the index set of one term of the next-row
// polynomial (shift = k + 1, offset = 2^k) and of every restriction of such a term to a row-range shard, where some index
// bits are fixed and drop out of the local index (leanmultisig_b200/sharded.py ShardedProductSumcheck.add_next)
__global__ void weights_add_strided_eq_kernel(uint32_t* __restrict__ w, uint64_t base, int shift, uint64_t offset,
                                              const uint32_t* __restrict__ pt, int pre, Ef scalar) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ((uint64_t)1 << pre)) return;
  Ef acc = scalar;
  for (int i = 0; i < pre; i++) {
    Ef z = ld_ef(pt + 5 * i);
    if (!((b >> (pre - 1 - i)) & 1)) {
#pragma unroll
      for (int c = 0; c < 5; c++) z.c[c] = kb_neg(z.c[c]);
      z.c[0] = kb_add(z.c[0], KB_R1);
    }
    acc = ef_mul(acc, z);
  }
  uint32_t* dst = w + 5 * (base + (b << shift) + offset);
  st_ef(dst, ef_add(ld_ef_rw(dst), acc));
}

cudaError_t weights_add_strided_eq(cudaStream_t stream, uint32_t* d_w, uint64_t base, uint32_t shift, uint64_t offset,
                                   const uint32_t* d_point, uint32_t pre, const uint32_t scalar[5]) {
  Ef s;
  for (int c = 0; c < 5; c++) s.c[c] = scalar[c];
  const uint64_t n = (uint64_t)1 << pre;
  weights_add_strided_eq_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(d_w, base, (int)shift, offset, d_point, (int)pre, s);
  count_launch();
  return cudaGetLastError();
}

// ---- batched base-field equality (STIR queries): w[x] += sum_q s_q eq(pt_q, x), pt_q in F^m
// tables: a[q][xh] = s_q * eq(pt_q[0..hi), xh)  (EF),  l[q][xl] = eq(pt_q[hi..m), xl)  (F)
__global__ void base_eq_tables_kernel(const uint32_t* __restrict__ pts, const uint32_t* __restrict__ scalars, int m,
                                      int hi_vars, int n_q, uint32_t* __restrict__ a, uint32_t* __restrict__ l) {
  const int lo_vars = m - hi_vars;
  const uint64_t n_hi = (uint64_t)1 << hi_vars, n_lo = (uint64_t)1 << lo_vars;
  const uint64_t per_q = n_hi + n_lo;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_q * n_q) return;
  const int q = (int)(idx / per_q);
  const uint64_t e = idx % per_q;
  const uint32_t* pt = pts + (uint64_t)q * m;
  if (e < n_hi) {
    uint32_t acc = KB_R1;
    for (int i = 0; i < hi_vars; i++) {
      const uint32_t z = __ldg(pt + i);
      acc = kb_mul(acc, ((e >> (hi_vars - 1 - i)) & 1) ? z : kb_sub(KB_R1, z));
    }
    const Ef s = ld_ef(scalars + 5 * q);
    st_ef(a + 5 * ((uint64_t)q * n_hi + e), ef_mul_base(s, acc));
  } else {
    const uint64_t x = e - n_hi;
    uint32_t acc = KB_R1;
    for (int i = 0; i < lo_vars; i++) {
      const uint32_t z = __ldg(pt + hi_vars + i);
      acc = kb_mul(acc, ((x >> (lo_vars - 1 - i)) & 1) ? z : kb_sub(KB_R1, z));
    }
    l[(uint64_t)q * n_lo + x] = acc;
  }
}
__global__ void weights_add_base_eq_kernel(uint32_t* __restrict__ w, uint64_t n, int lo_vars, int n_q,
                                           const uint32_t* __restrict__ a, const uint32_t* __restrict__ l) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const uint64_t n_lo = (uint64_t)1 << lo_vars, n_hi = n >> lo_vars;
  const uint64_t xh = x >> lo_vars, xl = x & (n_lo - 1);
  uint64_t acc[5] = {0, 0, 0, 0, 0};
  int terms = 0;
  for (int q = 0; q < n_q; q++) {
    const uint32_t f = __ldg(l + (uint64_t)q * n_lo + xl);
    const Ef e = ld_ef(a + 5 * ((uint64_t)q * n_hi + xh));
    if (terms == 3) {
#pragma unroll
      for (int c = 0; c < 5; c++) acc[c] = kb_fold(acc[c]);
      terms = 0;
    }
#pragma unroll
    for (int c = 0; c < 5; c++) acc[c] = mad_wide(f, e.c[c], acc[c]);
    terms++;
  }
  uint32_t* dst = w + 5 * x;
  Ef cur = ld_ef_rw(dst);
#pragma unroll
  for (int c = 0; c < 5; c++) cur.c[c] = kb_add(cur.c[c], kb_canon(kb_redc_lazy(kb_fold(acc[c]))));
  st_ef(dst, cur);
}

size_t weights_add_base_eq_scratch_words(uint32_t m, uint32_t n_q) {
  const int lo_vars = m < (uint32_t)SPLIT_LO ? (int)m : SPLIT_LO;
  const int hi_vars = (int)m - lo_vars;
  return (size_t)n_q * (5 * ((size_t)1 << hi_vars) + ((size_t)1 << lo_vars) + m + 5) + 16;
}

cudaError_t weights_add_base_eq(cudaStream_t stream, uint32_t* d_w, uint32_t m, const uint32_t* d_points, uint32_t n_q,
                                const uint32_t* d_scalars, uint32_t* d_scratch) {
  if (n_q == 0) return cudaSuccess;
  const int lo_vars = m < (uint32_t)SPLIT_LO ? (int)m : SPLIT_LO;
  const int hi_vars = (int)m - lo_vars;
  uint32_t* d_a = d_scratch;
  uint32_t* d_l = d_a + 5 * (size_t)n_q * ((size_t)1 << hi_vars);
  const uint64_t n_tab = (uint64_t)n_q * (((uint64_t)1 << hi_vars) + ((uint64_t)1 << lo_vars));
  base_eq_tables_kernel<<<(unsigned)((n_tab + 127) / 128), 128, 0, stream>>>(d_points, d_scalars, (int)m, hi_vars, (int)n_q,
                                                                              d_a, d_l);
  count_launch();
  const uint64_t n = (uint64_t)1 << m;
  weights_add_base_eq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_w, n, lo_vars, (int)n_q, d_a, d_l);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- sumcheck rounds
struct Acc10 {
  uint64_t a[10];  // c0[0..5), c2[0..5)
  int terms;
};

__device__ __forceinline__ void acc_init(Acc10& s) {
#pragma unroll
  for (int c = 0; c < 10; c++) s.a[c] = 0;
  s.terms = 0;
}
__device__ __forceinline__ void acc_maybe_fold(Acc10& s) {
  if (s.terms == 3) {
#pragma unroll
    for (int c = 0; c < 10; c++) s.a[c] = kb_fold(s.a[c]);
    s.terms = 0;
  }
}
// c0 = sum_{i < half} w[i] p[i];  c2 = sum (w[i+half] - w[i]) (p[i+half] - p[i]);  p entries >= live are zero.
template <int DIM>
__global__ void __launch_bounds__(256)
prod_round_kernel(const uint32_t* __restrict__ p, uint64_t live, const uint32_t* __restrict__ w, uint64_t half,
                  uint32_t* __restrict__ partial) {
  Ef c0 = ef_zero(), c2 = ef_zero();
  Acc10 s;
  acc_init(s);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (uint64_t)gridDim.x * blockDim.x) {
    const Ef w0 = ld_ef(w + 5 * i), w1 = ld_ef(w + 5 * (i + half));
    const Ef dw = ef_sub(w1, w0);
    if (DIM == 1) {
      const uint32_t p0 = i < live ? __ldg(p + i) : 0u;
      const uint32_t p1 = i + half < live ? __ldg(p + i + half) : 0u;
      const uint32_t dp = kb_sub(p1, p0);
      acc_maybe_fold(s);
#pragma unroll
      for (int c = 0; c < 5; c++) {
        s.a[c] = mad_wide(p0, w0.c[c], s.a[c]);
        s.a[5 + c] = mad_wide(dp, dw.c[c], s.a[5 + c]);
      }
      s.terms++;
    } else {
      const Ef p0 = i < live ? ld_ef(p + 5 * i) : ef_zero();
      const Ef p1 = i + half < live ? ld_ef(p + 5 * (i + half)) : ef_zero();
      c0 = ef_add(c0, ef_mul(w0, p0));
      c2 = ef_add(c2, ef_mul(dw, ef_sub(p1, p0)));
    }
  }
  if (DIM == 1) {
#pragma unroll
    for (int c = 0; c < 5; c++) {
      c0.c[c] = kb_canon(kb_redc_lazy(kb_fold(s.a[c])));
      c2.c[c] = kb_canon(kb_redc_lazy(kb_fold(s.a[5 + c])));
    }
  }
  block_reduce_pair(c0, c2, partial);
}

// Fold both tables with r (old length n = 4 * quarter) into p_out / w_out (EF, length 2 * quarter) and compute the
// next round's (c0, c2) on the folded tables in the same pass (product_computation.rs:242-304).
template <int DIM>
__global__ void __launch_bounds__(256)
prod_fold_round_kernel(const uint32_t* p, uint64_t live, const uint32_t* w, uint64_t quarter, Ef r, uint32_t* p_out,
                       uint32_t* w_out, uint32_t* __restrict__ partial) {
  Ef c0 = ef_zero(), c2 = ef_zero();
  const uint64_t half = 2 * quarter;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quarter; i += (uint64_t)gridDim.x * blockDim.x) {
    Ef x0, x1;
    if (DIM == 1) {
      const uint32_t a0 = i < live ? __ldg(p + i) : 0u, a1 = i + quarter < live ? __ldg(p + i + quarter) : 0u;
      const uint32_t b0 = i + half < live ? __ldg(p + i + half) : 0u;
      const uint32_t b1 = i + half + quarter < live ? __ldg(p + i + half + quarter) : 0u;
      x0 = ef_add_base(ef_mul_base(r, kb_sub(b0, a0)), a0);
      x1 = ef_add_base(ef_mul_base(r, kb_sub(b1, a1)), a1);
    } else {
      const Ef a0 = i < live ? ld_ef_rw(p + 5 * i) : ef_zero();
      const Ef a1 = i + quarter < live ? ld_ef_rw(p + 5 * (i + quarter)) : ef_zero();
      const Ef b0 = i + half < live ? ld_ef_rw(p + 5 * (i + half)) : ef_zero();
      const Ef b1 = i + half + quarter < live ? ld_ef_rw(p + 5 * (i + half + quarter)) : ef_zero();
      x0 = ef_add(a0, ef_mul(r, ef_sub(b0, a0)));
      x1 = ef_add(a1, ef_mul(r, ef_sub(b1, a1)));
    }
    const Ef u0 = ld_ef_rw(w + 5 * i), u1 = ld_ef_rw(w + 5 * (i + quarter));
    const Ef v0 = ld_ef_rw(w + 5 * (i + half)), v1 = ld_ef_rw(w + 5 * (i + half + quarter));
    const Ef y0 = ef_add(u0, ef_mul(r, ef_sub(v0, u0)));
    const Ef y1 = ef_add(u1, ef_mul(r, ef_sub(v1, u1)));
    st_ef(p_out + 5 * i, x0);
    st_ef(p_out + 5 * (i + quarter), x1);
    st_ef(w_out + 5 * i, y0);
    st_ef(w_out + 5 * (i + quarter), y1);
    c0 = ef_add(c0, ef_mul(y0, x0));
    c2 = ef_add(c2, ef_mul(ef_sub(y1, y0), ef_sub(x1, x0)));
  }
  block_reduce_pair(c0, c2, partial);
}

static unsigned round_grid(uint64_t work) {
  uint64_t blocks = (work + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;  // persistent-style grid: 8 CTAs of 256 threads per SM
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

size_t prod_round_scratch_words() { return 10 * (148 * 8) + 16; }

cudaError_t prod_round(cudaStream_t stream, const uint32_t* d_p, uint32_t dim, uint64_t live, const uint32_t* d_w, uint64_t n,
                       uint32_t* d_scratch, uint32_t* d_out10) {
  if (n < 2 || (dim != 1 && dim != 5)) return cudaErrorInvalidValue;
  const uint64_t half = n / 2;
  const unsigned grid = round_grid(half);
  if (dim == 1)
    prod_round_kernel<1><<<grid, 256, 0, stream>>>(d_p, live, d_w, half, d_scratch);
  else
    prod_round_kernel<5><<<grid, 256, 0, stream>>>(d_p, live, d_w, half, d_scratch);
  count_launch();
  sum_pair_partials_kernel<<<1, 256, 0, stream>>>(d_scratch, (int)grid, d_out10);
  count_launch();
  return cudaGetLastError();
}

cudaError_t prod_fold_round(cudaStream_t stream, const uint32_t* d_p, uint32_t dim, uint64_t live, const uint32_t* d_w,
                            uint64_t n, const uint32_t r[5], uint32_t* d_p_out, uint32_t* d_w_out, uint32_t* d_scratch,
                            uint32_t* d_out10) {
  if (n < 4 || (dim != 1 && dim != 5)) return cudaErrorInvalidValue;
  Ef rr;
  for (int c = 0; c < 5; c++) rr.c[c] = r[c];
  const uint64_t quarter = n / 4;
  const unsigned grid = round_grid(quarter);
  if (dim == 1)
    prod_fold_round_kernel<1><<<grid, 256, 0, stream>>>(d_p, live, d_w, quarter, rr, d_p_out, d_w_out, d_scratch);
  else
    prod_fold_round_kernel<5><<<grid, 256, 0, stream>>>(d_p, live, d_w, quarter, rr, d_p_out, d_w_out, d_scratch);
  count_launch();
  sum_pair_partials_kernel<<<1, 256, 0, stream>>>(d_scratch, (int)grid, d_out10);
  count_launch();
  return cudaGetLastError();
}

}  // namespace lm

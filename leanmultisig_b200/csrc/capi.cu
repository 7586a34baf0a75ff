// extern "C" boundary of libleanmultisig_b200.so — see include/leanmultisig_b200.h for the contract and the
// reference call site each entry point replaces.  Host-side logic only: shape derivation that mirrors
// WhirConfig::commit (crates/whir/src/commit.rs:64-85), device memory ownership, stream plumbing.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <new>
#include <utility>
#include <vector>

#include "../../include/leanmultisig_b200.h"
#include "merkle.h"
#include "ntt.h"
#include "poly.h"
#include "sumcheck.h"
#include "air.h"
#include "eqtab.cuh"
#include "gkr.h"
#include "launch_count.h"

namespace lm {
std::atomic<uint64_t> g_kernel_launches{0};
}

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaErrorMemoryAllocation) return fail(LM_ERR_OOM, "%s: %s", what, cudaGetErrorString(e));
  return fail(LM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e__ = (call);                            \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

}  // namespace

// error reporting for the other translation units of the library (spine.cu)
int lm_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }

// Cache of device allocations, the GPU analogue of the reference's per-proof bump arena
// (crates/backend/zk-alloc/src/lib.rs:102-115): a proof allocates the same few dozen buffer sizes every time, and
// cudaMalloc / cudaFree of GiB-sized buffers cost milliseconds each and synchronise the device, which is where the 2-10x
// run-to-run spread of whole-proof times came from.  Freed buffers are kept by exact size and handed out again; the cache
// is bounded by a byte budget and dropped on allocation failure.
struct DevicePool {
  std::multimap<size_t, void*> cache;   // size -> free buffer
  std::map<void*, size_t> live;         // buffers handed out by alloc()
  size_t cached_bytes = 0;
  size_t budget_bytes = (size_t)64 << 30;
  cudaError_t get(size_t bytes, void** out) {
    if (bytes == 0) bytes = 1;
    auto it = cache.find(bytes);
    if (it != cache.end()) {
      *out = it->second;
      cached_bytes -= bytes;
      cache.erase(it);
      return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {  // drop the cache and retry once
      cudaGetLastError();
      release_all();
      e = cudaMalloc(out, bytes);
    }
    return e;
  }
  void put(size_t bytes, void* p) {
    if (!p) return;
    if (bytes == 0) bytes = 1;
    while (!cache.empty() && cached_bytes + bytes > budget_bytes) {  // evict the largest entries first
      auto last = std::prev(cache.end());
      cudaFree(last->second);
      cached_bytes -= last->first;
      cache.erase(last);
    }
    cache.emplace(bytes, p);
    cached_bytes += bytes;
  }
  // address-keyed form: the size is remembered, free() needs only the pointer
  template <class T>
  cudaError_t alloc(T** out, size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = get(bytes, &p);
    if (e == cudaSuccess) live[p] = bytes ? bytes : 1;
    *out = static_cast<T*>(p);
    return e;
  }
  void free(void* p) {
    if (!p) return;
    auto it = live.find(p);
    if (it == live.end()) {
      cudaFree(p);
      return;
    }
    const size_t bytes = it->second;
    live.erase(it);
    put(bytes, p);
  }
  void release_all() {
    for (auto& e : cache) cudaFree(e.second);
    cache.clear();
    cached_bytes = 0;
  }
};

struct lm_ctx {
  DevicePool pool;
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host-to-device copies that overlap with compute on `stream`
  cudaEvent_t ev_copy[8] = {};         // copy of column group g finished
  cudaEvent_t ev_start = nullptr;
  uint32_t* d_tw = nullptr;
  unsigned tw_log_n = 0;
  uint32_t* d_scratch = nullptr;  // grows on demand (MLE evaluation tables)
  size_t scratch_words = 0;
  uint32_t* d_small = nullptr;  // 64 words: points / results of host-facing calls
  uint32_t* d_point = nullptr;  // up to 64 x 5 words

  int ensure_scratch(size_t words) {
    if (words <= scratch_words) return LM_OK;
    if (d_scratch) cudaFree(d_scratch);
    d_scratch = nullptr;
    scratch_words = 0;
    CU(cudaMalloc(&d_scratch, words * sizeof(uint32_t)));
    scratch_words = words;
    return LM_OK;
  }
};

struct lm_tree {
  lm_ctx* ctx = nullptr;
  uint32_t n_vars = 0, elem_dim = 1;
  uint64_t actual_len = 0;  // in elements
  uint64_t height = 0;
  uint32_t full_width = 0, stored_width = 0, effective_width = 0;  // in words
  uint32_t* d_evals = nullptr;     // live prefix of the polynomial (actual_len * elem_dim words) or null
  bool owns_evals = false;
  uint32_t* d_codeword = nullptr;  // height x stored_width
  uint32_t* d_layers = nullptr;    // (2 height - 1) x 8
  size_t evals_bytes = 0, codeword_bytes = 0, layers_bytes = 0;
};

// Product-sumcheck session of WHIR open (reference: SumcheckSingle, crates/whir/src/open.rs:323-446).
struct lm_sumcheck {
  lm_ctx* ctx = nullptr;
  uint32_t n_vars = 0;       // current number of variables of both tables
  uint32_t p_dim = 1;        // 1 while the polynomial is still the (borrowed) base-field witness
  uint64_t p_live = 0;       // entries of p that may be non-zero
  const uint32_t* d_p = nullptr;  // current polynomial table
  uint32_t* d_p_owned = nullptr;  // EF table owned by the session (after the first fold) or an uploaded copy
  uint32_t* d_w = nullptr;        // weights, 2^n_vars EF, folded in place
  uint32_t* d_out10 = nullptr;    // (c0, c2) of the last round
  uint32_t* d_scratch = nullptr;
  size_t scratch_words = 0;
  int ensure_scratch(size_t words) {
    if (words <= scratch_words) return LM_OK;
    if (d_scratch) ctx->pool.free(d_scratch);
    d_scratch = nullptr;
    scratch_words = 0;
    CU(ctx->pool.alloc(&d_scratch, words * sizeof(uint32_t)));
    scratch_words = words;
    return LM_OK;
  }
};

// AIR sumcheck session (reference: AirSumcheckSession, crates/sub_protocols/src/air_sumcheck.rs:45-292)
struct lm_air {
  lm_ctx* ctx = nullptr;
  uint32_t table_id = 0, n_cols = 0, n_shift = 0, degree = 0;
  uint32_t n_vars0 = 0;     // variables at creation (the eq tables are indexed by it)
  uint32_t log_n = 0;       // current number of variables = rounds left
  bool round_done = false;  // lm_air_round has run for the current variable and lm_air_fold has not yet
  // extension_op / poseidon16 (air_generic.cu): the fold is a pass of its own
  uint32_t dim = 1;             // 1 until the first fold
  uint32_t* d_cols = nullptr;   // current columns (base u32[c][n] or planes u32[c][5][n])
  uint32_t* d_spare = nullptr;  // ping-pong target of the next fold
  size_t cols_words = 0, spare_words = 0;
  // execution table (air.cu): folds are fused into the rounds, see air.h AirExecMode
  bool exec = false;
  bool borrowed = false;  // d_cols belongs to the caller (lm_air_new_dev)
  bool src_is_base = true;
  uint32_t src_log_rows = 0;  // rows of the table the next round reads
  uint32_t pending = 0;       // challenges not yet applied to it (base: <= 2, extension: <= 1)
  uint32_t halo[2] = {0, 0};
  uint32_t r_host[5] = {0, 0, 0, 0, 0};  // the latest challenge (round 1 of the execution table runs on its powers)
  uint32_t* d_ef[2] = {nullptr, nullptr};
  size_t ef_words[2] = {0, 0};
  int src_buf = 0;
  lm::AirDev* d_dev = nullptr;
  uint32_t* d_eq_tab = nullptr;
  uint32_t* d_scratch = nullptr;
  uint32_t* d_out = nullptr;
  std::vector<uint32_t> alpha, la;
  uint32_t beta[5] = {0, 0, 0, 0, 0};
};

// Quotient-GKR session (reference: prove_gkr_quotient, crates/sub_protocols/src/quotient_gkr/mod.rs:31-141)
struct lm_gkr {
  lm_ctx* ctx = nullptr;
  uint32_t n_vars = 0;  // of layer 0
  // layer l has n_vars - l variables; layer 0 numerators are base field; every EF array is five coefficient planes (gkr.h)
  std::vector<uint32_t*> nums, dens;
  uint32_t* d_w[2] = {nullptr, nullptr};  // ping-pong working tables of the current layer sumcheck
  uint32_t* d_eq_tab = nullptr;           // prefix eq tables of the current layer
  uint32_t* d_partial = nullptr;          // per-CTA partial sums
  lm::GkrDev* d_g = nullptr;              // device-side state of the layer sumcheck (claim point, challenges, running sum)
  lm::DevFs* d_fs = nullptr;              // device challenger (lm_gkr_prove)
  uint32_t* d_tr = nullptr;               // transcript words appended by the device
  uint32_t tr_cap = 0;
  // current layer sumcheck when the transcript is driven by the caller (lm_gkr_layer_begin / round / fold / layer_end)
  int cur_layer = -1;
  uint32_t cur_k = 0;     // claim variables = rounds of the current layer
  uint32_t cur_rnd = 0;   // next round
  bool round_done = false;  // lm_gkr_round has run for cur_rnd and lm_gkr_fold has not yet
  void* arena = nullptr;                             // one allocation for the upper layers, working tables and small buffers
  uint32_t top_vars = 5;                             // the up pass stops at 2^top_vars fractions
  lm::GkrLayerArgs layer_args(int layer, uint32_t k, bool device_fs) const {
    lm::GkrLayerArgs a{};
    a.nums = nums[layer], a.dens = dens[layer];
    a.num_dim = layer == 0 ? 1 : 5;
    a.k = k;
    a.w[0] = d_w[0], a.w[1] = d_w[1];
    a.eq_tab = d_eq_tab, a.partial = d_partial, a.g = d_g;
    a.fs = device_fs ? d_fs : nullptr;
    a.tr = device_fs ? d_tr : nullptr;
    return a;
  }
};
static const uint32_t LM_GKR_TOP_VARS = 5;  // N_VARS_TO_SEND_GKR_COEFFS (crates/sub_protocols/src/lib.rs:14)

extern "C" {

const char* lm_last_error(void) { return g_err; }

uint64_t lm_kernel_launches(void) { return lm::g_kernel_launches.load(); }

int lm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int lm_init(int device, uint32_t max_log_domain, lm_ctx** out_ctx) {
  if (!out_ctx) return fail(LM_ERR_INVALID, "lm_init: out_ctx is null");
  *out_ctx = nullptr;
  if (max_log_domain > 24) return fail(LM_ERR_INVALID, "lm_init: KoalaBear two-adicity is 24, got %u", max_log_domain);
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return fail(LM_ERR_NO_DEVICE, "lm_init: no CUDA device visible; this library has no CPU path");
  if (device < 0 || device >= n) return fail(LM_ERR_INVALID, "lm_init: device %d out of range (%d visible)", device, n);
  CU(cudaSetDevice(device));
  lm_ctx* c = new (std::nothrow) lm_ctx();
  if (!c) return fail(LM_ERR_OOM, "lm_init: host allocation failed");
  c->device = device;
  CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev_copy) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
  c->tw_log_n = max_log_domain;
  if (max_log_domain > 0) {
    CU(cudaMalloc(&c->d_tw, (((size_t)1 << (max_log_domain - 1)) + lm::NTT_TW_SCRATCH_WORDS) * sizeof(uint32_t)));
    CU(lm::ntt_fill_twiddles(c->stream, c->d_tw, max_log_domain));
  }
  CU(cudaMalloc(&c->d_small, 64 * sizeof(uint32_t)));
  CU(cudaMalloc(&c->d_point, 64 * 5 * sizeof(uint32_t)));
  CU(cudaStreamSynchronize(c->stream));
  *out_ctx = c;
  return LM_OK;
}

int lm_destroy(lm_ctx* c) {
  if (!c) return LM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->pool.release_all();
  if (c->d_tw) cudaFree(c->d_tw);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->d_small) cudaFree(c->d_small);
  if (c->d_point) cudaFree(c->d_point);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto ev : c->ev_copy)
    if (ev) cudaEventDestroy(ev);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  delete c;
  return LM_OK;
}

int lm_set_stream(lm_ctx* c, void* s) {
  if (!c) return fail(LM_ERR_INVALID, "lm_set_stream: ctx is null");
  c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->own_stream;
  return LM_OK;
}

int lm_sync(lm_ctx* c) {
  if (!c) return fail(LM_ERR_INVALID, "lm_sync: ctx is null");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_host_register(void* ptr, size_t bytes) {
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return LM_OK;
}
int lm_host_unregister(void* ptr) {
  CU(cudaHostUnregister(ptr));
  return LM_OK;
}

// ------------------------------------------------------------------------------------------ device layer
int lm_dev_alloc(lm_ctx* c, size_t bytes, void** out) {
  if (!c || !out) return fail(LM_ERR_INVALID, "lm_dev_alloc: null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(out, bytes ? bytes : 1));
  return LM_OK;
}
int lm_dev_free(lm_ctx* c, void* p) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_free: ctx is null");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaFree(p));
  return LM_OK;
}
int lm_dev_upload(lm_ctx* c, void* d, const void* h, size_t bytes) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_upload: ctx is null");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}
int lm_dev_download(lm_ctx* c, void* h, const void* d, size_t bytes) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_download: ctx is null");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_dev_poseidon1(lm_ctx* c, uint32_t* d_states, uint64_t n, int compress) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_poseidon1: ctx is null");
  CU(cudaSetDevice(c->device));
  CU(lm::poseidon1_states(c->stream, d_states, n, compress));
  return LM_OK;
}

int lm_pow_grind(lm_ctx* c, const uint32_t* state, uint32_t bits, uint64_t* witness) {
  if (!c || !state || !witness) return fail(LM_ERR_INVALID, "lm_pow_grind: null argument");
  if (bits > 30) return fail(LM_ERR_INVALID, "lm_pow_grind: bits %u > 30", bits);
  CU(cudaSetDevice(c->device));
  CU(lm::pow_grind(c->stream, state, bits, 0, reinterpret_cast<unsigned long long*>(c->d_small), witness));
  return LM_OK;
}

int lm_host_poseidon1_permute(uint32_t* state) {
  if (!state) return fail(LM_ERR_INVALID, "lm_host_poseidon1_permute: null state");
  lm::poseidon1_permute_host(state);
  return LM_OK;
}
int lm_host_eq_gemm_model(uint32_t* w, const uint32_t* hi, const uint32_t* lo, uint32_t n_statements, uint32_t hi_vars, uint32_t lo_vars) {
  if (!w || !hi || !lo) return fail(LM_ERR_INVALID, "lm_host_eq_gemm_model: null argument");
  if (n_statements < 1 || n_statements > 12 || hi_vars < 2 || hi_vars > 16 || lo_vars > 20)
    return fail(LM_ERR_INVALID, "lm_host_eq_gemm_model: 1..12 statements, 2..16 high variables, <= 20 low variables");
  lm::weights_gemm_model_host(w, hi, lo, (int)n_statements, (int)hi_vars, (int)lo_vars);
  return LM_OK;
}
uint64_t lm_host_poseidon1_umma_image(uint8_t* out, uint64_t capacity) {
  return (uint64_t)lm::poseidon1_umma_image_host(out, (size_t)capacity);
}
int lm_host_poseidon1_umma_model(uint32_t* state) {
  if (!state) return fail(LM_ERR_INVALID, "lm_host_poseidon1_umma_model: null state");
  lm::poseidon1_permute_umma_model_host(state);
  return LM_OK;
}

int lm_dev_reorder_and_dft(lm_ctx* c, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim, uint32_t folding,
                           uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_out) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_reorder_and_dft: ctx is null");
  if (dim != 1 && dim != 5) return fail(LM_ERR_INVALID, "elem_dim must be 1 or 5, got %u", dim);
  if (folding > n_vars) return fail(LM_ERR_INVALID, "folding_factor %u > n_vars %u", folding, n_vars);
  if (dft_n_cols > (1u << folding)) return fail(LM_ERR_INVALID, "dft_n_cols %u > 2^folding", dft_n_cols);
  if (n_vars + log_inv_rate - folding > c->tw_log_n)
    return fail(LM_ERR_INVALID, "domain 2^%u exceeds the twiddle table 2^%u given to lm_init",
                n_vars + log_inv_rate - folding, c->tw_log_n);
  CU(cudaSetDevice(c->device));
  CU(lm::ntt_reorder_and_dft(c->stream, d_evals, n_vars, dim, folding, log_inv_rate, dft_n_cols, d_out, c->d_tw,
                             c->tw_log_n));
  return LM_OK;
}

int lm_dev_reorder_and_dft_scatter(lm_ctx* c, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding, uint32_t log_inv_rate,
                                   uint32_t dft_n_cols, uint32_t* d_work, const uint64_t* peer_mats, uint32_t world,
                                   uint32_t rank) {
  return lm_dev_reorder_and_dft_scatter_cols(c, d_evals, n_vars, folding, log_inv_rate, dft_n_cols, d_work, peer_mats, world, rank,
                                             0, dft_n_cols);
}

int lm_dev_reorder_and_dft_scatter_cols(lm_ctx* c, const uint32_t* d_evals, uint32_t n_vars, uint32_t folding,
                                        uint32_t log_inv_rate, uint32_t dft_n_cols, uint32_t* d_work, const uint64_t* peer_mats,
                                        uint32_t world, uint32_t rank, uint32_t col_begin, uint32_t col_count) {
  if (!c || !d_evals || !d_work || !peer_mats) return fail(LM_ERR_INVALID, "lm_dev_reorder_and_dft_scatter: null argument");
  if (col_count == 0 || col_begin % 8 || col_begin + col_count > dft_n_cols || (col_count % 8 && col_begin + col_count != dft_n_cols))
    return fail(LM_ERR_INVALID, "lm_dev_reorder_and_dft_scatter_cols: column range [%u, +%u) of %u", col_begin, col_count, dft_n_cols);
  if (world < 2 || world > 16 || (world & (world - 1)) || rank >= world)
    return fail(LM_ERR_INVALID, "lm_dev_reorder_and_dft_scatter: world %u / rank %u", world, rank);
  if (folding > n_vars) return fail(LM_ERR_INVALID, "folding_factor %u > n_vars %u", folding, n_vars);
  if (dft_n_cols > (1u << folding) || dft_n_cols % 4 || dft_n_cols == 0)
    return fail(LM_ERR_INVALID, "lm_dev_reorder_and_dft_scatter: dft_n_cols %u must be a multiple of 4 and <= 2^folding", dft_n_cols);
  if (n_vars + log_inv_rate - folding > c->tw_log_n)
    return fail(LM_ERR_INVALID, "domain 2^%u exceeds the twiddle table 2^%u given to lm_init",
                n_vars + log_inv_rate - folding, c->tw_log_n);
  CU(cudaSetDevice(c->device));
  uint32_t* peers[16];
  for (uint32_t q = 0; q < world; q++) {
    if (!peer_mats[q]) return fail(LM_ERR_INVALID, "lm_dev_reorder_and_dft_scatter: peer %u is null", q);
    peers[q] = reinterpret_cast<uint32_t*>(peer_mats[q]);
  }
  CU(lm::ntt_reorder_and_dft_scatter(c->stream, d_evals, n_vars, folding, log_inv_rate, dft_n_cols, d_work, peers, world, rank,
                                     c->d_tw, c->tw_log_n, col_begin, col_count));
  return LM_OK;
}

int lm_dev_ipc_export(lm_ctx* c, const void* d_ptr, uint8_t handle[64]) {
  if (!c || !d_ptr || !handle) return fail(LM_ERR_INVALID, "lm_dev_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
  memcpy(handle, &h, 64);
  return LM_OK;
}

int lm_dev_ipc_open(lm_ctx* c, const uint8_t handle[64], void** d_peer) {
  if (!c || !handle || !d_peer) return fail(LM_ERR_INVALID, "lm_dev_ipc_open: null argument");
  CU(cudaSetDevice(c->device));  // opened from the device whose kernels will store through the mapping
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU(cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess));
  return LM_OK;
}

int lm_dev_ipc_close(lm_ctx* c, void* d_peer) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_ipc_close: null argument");
  if (!d_peer) return LM_OK;
  CU(cudaSetDevice(c->device));
  CU(cudaIpcCloseMemHandle(d_peer));
  return LM_OK;
}

int lm_dev_dft(lm_ctx* c, uint32_t* d_mat, uint64_t h, uint64_t w) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_dft: ctx is null");
  if (h == 0 || (h & (h - 1))) return fail(LM_ERR_INVALID, "lm_dev_dft: height %llu is not a power of two", (unsigned long long)h);
  if (h > ((uint64_t)1 << c->tw_log_n)) return fail(LM_ERR_INVALID, "lm_dev_dft: height exceeds the twiddle table");
  CU(cudaSetDevice(c->device));
  CU(lm::ntt_dft_batch_by_evals(c->stream, d_mat, h, w, 0, c->d_tw, c->tw_log_n));
  return LM_OK;
}

int lm_dev_dft_layers_mapped(lm_ctx* c, uint32_t* d_mat, uint64_t w, uint32_t log_h, uint32_t l_first, uint64_t n_blocks,
                             uint64_t run, uint64_t block, uint64_t offset) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped: ctx is null");
  if (w % 4) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped: width %llu is not a multiple of 4", (unsigned long long)w);
  if (log_h > c->tw_log_n) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped: domain exceeds the twiddle table");
  CU(cudaSetDevice(c->device));
  CU(lm::ntt_layers_mapped(c->stream, d_mat, w, log_h, l_first, n_blocks, run, block, offset, c->d_tw, c->tw_log_n));
  return LM_OK;
}

int lm_dev_dft_layers_mapped_cols(lm_ctx* c, uint32_t* d_mat, uint64_t w, uint32_t log_h, uint32_t l_first, uint64_t n_blocks,
                                  uint64_t run, uint64_t block, uint64_t offset, uint64_t col_begin, uint64_t col_count) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped_cols: ctx is null");
  if (w % 4 || col_begin % 4 || col_count % 4 || col_count == 0 || col_begin + col_count > w)
    return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped_cols: column range [%llu, +%llu) of %llu", (unsigned long long)col_begin,
                (unsigned long long)col_count, (unsigned long long)w);
  if (log_h > c->tw_log_n) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped_cols: domain exceeds the twiddle table");
  CU(cudaSetDevice(c->device));
  CU(lm::ntt_layers_mapped(c->stream, d_mat, w, log_h, l_first, n_blocks, run, block, offset, c->d_tw, c->tw_log_n, col_begin,
                           col_count));
  return LM_OK;
}

int lm_dev_dft_layers_mapped_out(lm_ctx* c, const uint32_t* d_mat, uint32_t* d_out, uint64_t w, uint32_t log_h, uint32_t l_first,
                                 uint64_t n_blocks, uint64_t run, uint64_t block, uint64_t offset, uint64_t col_begin,
                                 uint64_t col_count) {
  if (!c || !d_mat || !d_out) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped_out: null argument");
  if (w % 4 || col_begin % 4 || col_count % 4 || col_begin + col_count > w)
    return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped_out: column range [%llu, +%llu) of %llu", (unsigned long long)col_begin,
                (unsigned long long)col_count, (unsigned long long)w);
  if (log_h > c->tw_log_n) return fail(LM_ERR_INVALID, "lm_dev_dft_layers_mapped_out: domain exceeds the twiddle table");
  CU(cudaSetDevice(c->device));
  CU(lm::ntt_layers_mapped(c->stream, const_cast<uint32_t*>(d_mat), w, log_h, l_first, n_blocks, run, block, offset, c->d_tw,
                           c->tw_log_n, col_begin, col_count, d_out));
  return LM_OK;
}

int lm_dev_merkle_absorb(lm_ctx* c, const uint32_t* d_mat, uint64_t h, uint32_t stored_w, uint32_t full_w, uint32_t eff_w,
                         uint32_t chunk_hi, uint32_t count, uint32_t* d_digests) {
  if (!c || !d_mat || !d_digests) return fail(LM_ERR_INVALID, "lm_dev_merkle_absorb: null argument");
  if (!lm::merkle_leaf_chunked_ok(stored_w, full_w, eff_w))
    return fail(LM_ERR_INVALID, "lm_dev_merkle_absorb: widths full=%u stored=%u effective=%u cannot be hashed chunk by chunk", full_w,
                stored_w, eff_w);
  if (chunk_hi >= eff_w / 8 || count == 0 || count > chunk_hi + 1)
    return fail(LM_ERR_INVALID, "lm_dev_merkle_absorb: chunks %u down %u of %u", chunk_hi, count, eff_w / 8);
  CU(cudaSetDevice(c->device));
  CU(lm::merkle_leaf_absorb_chunks(c->stream, d_mat, h, stored_w, full_w, eff_w, chunk_hi, count, d_digests));
  return LM_OK;
}

int lm_dev_merkle_tree(lm_ctx* c, const uint32_t* d_mat, uint64_t h, uint32_t stored_w, uint32_t full_w,
                       uint32_t eff_w, uint32_t* d_layers) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_merkle_tree: ctx is null");
  if (h == 0 || (h & (h - 1))) return fail(LM_ERR_INVALID, "merkle: height %llu is not a power of two", (unsigned long long)h);
  if (full_w % 8 || full_w < 16 || eff_w > full_w || stored_w > full_w)
    return fail(LM_ERR_INVALID, "merkle: widths full=%u stored=%u effective=%u", full_w, stored_w, eff_w);
  CU(cudaSetDevice(c->device));
  CU(lm::merkle_leaf_digests(c->stream, d_mat, h, stored_w, full_w, eff_w, d_layers));
  CU(lm::merkle_tree_from_digests(c->stream, d_layers, h));
  return LM_OK;
}

int lm_dev_merkle_leaves(lm_ctx* c, const uint32_t* d_mat, uint64_t h, uint32_t stored_w, uint32_t full_w,
                         uint32_t eff_w, uint32_t* d_layers) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_merkle_leaves: ctx is null");
  if (full_w % 8 || full_w < 16 || eff_w > full_w || stored_w > full_w)
    return fail(LM_ERR_INVALID, "merkle: widths full=%u stored=%u effective=%u", full_w, stored_w, eff_w);
  CU(cudaSetDevice(c->device));
  CU(lm::merkle_leaf_digests(c->stream, d_mat, h, stored_w, full_w, eff_w, d_layers));
  return LM_OK;
}

int lm_dev_merkle_levels(lm_ctx* c, uint32_t* d_layers, uint64_t h) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_merkle_levels: ctx is null");
  if (h == 0 || (h & (h - 1))) return fail(LM_ERR_INVALID, "merkle: height %llu is not a power of two", (unsigned long long)h);
  CU(cudaSetDevice(c->device));
  CU(lm::merkle_tree_from_digests(c->stream, d_layers, h));
  return LM_OK;
}

int lm_dev_mle_eval(lm_ctx* c, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim, uint64_t live_len,
                    const uint32_t* d_point, uint32_t* d_out) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_mle_eval: ctx is null");
  if (dim != 1 && dim != 5) return fail(LM_ERR_INVALID, "elem_dim must be 1 or 5, got %u", dim);
  if (n_vars > 34) return fail(LM_ERR_INVALID, "n_vars %u too large", n_vars);
  CU(cudaSetDevice(c->device));
  int rc = c->ensure_scratch(lm::mle_eval_scratch_words(n_vars));
  if (rc != LM_OK) return rc;
  CU(lm::mle_eval(c->stream, d_evals, n_vars, dim, live_len, d_point, c->d_scratch, d_out));
  return LM_OK;
}

int lm_dev_fold_msb(lm_ctx* c, const uint32_t* d_in, uint64_t n_in, uint32_t dim, const uint32_t r[5], uint32_t* d_out) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_fold_msb: ctx is null");
  if (n_in < 2 || (n_in & (n_in - 1))) return fail(LM_ERR_INVALID, "fold: length must be a power of two >= 2");
  CU(cudaSetDevice(c->device));
  CU(lm::fold_msb(c->stream, d_in, n_in, dim, n_in, r, d_out));
  return LM_OK;
}

int lm_dev_eq_table(lm_ctx* c, const uint32_t* point, uint32_t k, const uint32_t scalar[5], uint32_t* d_out) {
  if (!c) return fail(LM_ERR_INVALID, "lm_dev_eq_table: ctx is null");
  if (k > 64) return fail(LM_ERR_INVALID, "eq table: k = %u too large", k);
  CU(cudaSetDevice(c->device));
  if (k) CU(cudaMemcpyAsync(c->d_point, point, (size_t)k * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(lm::eq_table(c->stream, c->d_point, (int)k, scalar, d_out));
  CU(cudaStreamSynchronize(c->stream));  // c->d_point is reused by the next call
  return LM_OK;
}

// ------------------------------------------------------------------------------------------ commit / open
static int commit_impl(lm_ctx* c, const uint32_t* evals, bool evals_on_device, uint32_t n_vars, uint32_t dim,
                       uint64_t actual_len, uint32_t folding, uint32_t log_inv_rate, bool retain_evals,
                       lm_tree** out_tree, uint32_t out_root[8]) {
  if (!c || !out_tree || !out_root) return fail(LM_ERR_INVALID, "lm_commit: null argument");
  *out_tree = nullptr;
  if (dim != 1 && dim != 5) return fail(LM_ERR_INVALID, "lm_commit: elem_dim must be 1 or 5, got %u", dim);
  if (n_vars > 32) return fail(LM_ERR_INVALID, "lm_commit: n_vars %u too large", n_vars);
  if (folding > n_vars) return fail(LM_ERR_INVALID, "lm_commit: folding_factor %u > n_vars %u", folding, n_vars);
  if (folding < 1) return fail(LM_ERR_INVALID, "lm_commit: folding_factor must be >= 1 (leaf width >= 16 words)");
  const uint64_t evals_len = (uint64_t)1 << n_vars;
  if (actual_len > evals_len) return fail(LM_ERR_INVALID, "lm_commit: actual_len exceeds 2^n_vars");
  const uint32_t log_h = n_vars + log_inv_rate - folding;
  if (log_h > c->tw_log_n)
    return fail(LM_ERR_INVALID, "lm_commit: domain 2^%u exceeds the twiddle table 2^%u given to lm_init", log_h, c->tw_log_n);
  // commit.rs:68-74: number of column blocks that hold data
  const uint64_t n_blocks = (uint64_t)1 << folding;
  const uint64_t block_len = evals_len / n_blocks;
  uint64_t eff_cols = (actual_len + block_len - 1) / block_len;
  // stored width: whole columns, rounded up so that rows stay 16-byte aligned (any width >= eff_cols gives the
  // same root and openings, merkle.rs:205-211,251-288)
  uint64_t dft_cols = eff_cols;
  while ((dft_cols * dim) % 4 != 0 && dft_cols < n_blocks) dft_cols++;
  if (dft_cols == 0) dft_cols = (dim == 1) ? 4 : 4;
  if (dft_cols > n_blocks) dft_cols = n_blocks;
  const uint32_t full_w = (uint32_t)(n_blocks * dim);
  if (full_w % 8 != 0 || full_w < 16)
    return fail(LM_ERR_INVALID, "lm_commit: full leaf width %u must be a multiple of 8 and >= 16", full_w);

  CU(cudaSetDevice(c->device));
  lm_tree* t = new (std::nothrow) lm_tree();
  if (!t) return fail(LM_ERR_OOM, "lm_commit: host allocation failed");
  t->ctx = c;
  t->n_vars = n_vars;
  t->elem_dim = dim;
  t->actual_len = actual_len;
  t->height = (uint64_t)1 << log_h;
  t->full_width = full_w;
  t->stored_width = (uint32_t)(dft_cols * dim);
  t->effective_width = (uint32_t)(eff_cols * dim);
  auto cleanup = [&](int rc) {
    lm_tree_free(t);
    return rc;
  };
#define CUT(call)                                                      \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return cleanup(cuda_fail(e__, #call));     \
  } while (0)

  // the gather reads whole columns: dft_cols * block_len elements (zero beyond actual_len)
  // ... and lm_tree_eval reads the live prefix in rows of min(2^n_vars, 1024) elements
  uint64_t need_len = dft_cols * block_len;
  const uint64_t eval_row = evals_len < 1024 ? evals_len : 1024;
  const uint64_t eval_len = (actual_len + eval_row - 1) / eval_row * eval_row;
  if (eval_len > need_len) need_len = eval_len;
  const uint64_t need_words = need_len * dim, live_words = actual_len * dim;
  const uint32_t* d_src = nullptr;
  if (evals_on_device && !retain_evals) {
    d_src = evals;  // caller guarantees 2^n_vars elements are addressable
  } else {
    t->evals_bytes = (need_words ? need_words : 1) * sizeof(uint32_t);
    CUT(c->pool.get(t->evals_bytes, reinterpret_cast<void**>(&t->d_evals)));
    t->owns_evals = true;
    d_src = t->d_evals;
  }
  t->codeword_bytes = t->height * t->stored_width * sizeof(uint32_t);
  t->layers_bytes = (2 * t->height - 1) * 8 * sizeof(uint32_t);
  CUT(c->pool.get(t->codeword_bytes, reinterpret_cast<void**>(&t->d_codeword)));
  CUT(c->pool.get(t->layers_bytes, reinterpret_cast<void**>(&t->d_layers)));
  // Host input with whole live columns: pipeline  copy(group g) -> transform(group g) -> sponge step(s) of group g
  // over column groups taken right to left (the sponge absorbs the row right to left), so that PCIe, the NTT and
  // the hash overlap.  Everything else takes the plain path.
  const uint32_t stored_w = t->stored_width, eff_w = t->effective_width;
  const bool pipelined = !evals_on_device && dim == 1 && stored_w == eff_w && eff_w % 8 == 0 && eff_w >= 16 &&
                         actual_len == eff_cols * block_len && lm::merkle_leaf_chunked_ok(stored_w, full_w, eff_w);
  if (!pipelined) {
    if (!evals_on_device || retain_evals) {
      if (live_words)
        CUT(cudaMemcpyAsync(t->d_evals, evals, live_words * sizeof(uint32_t),
                            evals_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
      if (need_words > live_words)
        CUT(cudaMemsetAsync(t->d_evals + live_words, 0, (need_words - live_words) * sizeof(uint32_t), c->stream));
    }
    CUT(lm::ntt_reorder_and_dft(c->stream, d_src, n_vars, dim, folding, log_inv_rate, (uint32_t)dft_cols, t->d_codeword,
                                c->d_tw, c->tw_log_n));
    CUT(lm::merkle_leaf_digests(c->stream, t->d_codeword, t->height, t->stored_width, t->full_width, t->effective_width,
                                t->d_layers));
  } else {
    // Up to 8 equal column groups of whole rate chunks (8 columns each).  With the tensor-core sponge the transform + hash of
    // a group takes about as long as its copy over PCIe (1.2 ms each per 64 MiB at config 2), so equal groups keep both busy
    // and leave one group's compute + the tree after the last byte: 13.7 ms end to end against 15.7 ms with the round-1
    // schedule of 1, 1, 2, 4, .. chunks (LM_COMMIT_DOUBLING_GROUPS=1), whose last group is half of the matrix.
    static const bool even_groups = getenv("LM_COMMIT_DOUBLING_GROUPS") == nullptr;
    if (need_words > live_words)
      CUT(cudaMemsetAsync(t->d_evals + live_words, 0, (need_words - live_words) * sizeof(uint32_t), c->stream));
    CUT(cudaEventRecord(c->ev_start, c->stream));
    CUT(cudaStreamWaitEvent(c->copy_stream, c->ev_start, 0));
    const uint32_t n_chunks = eff_w / 8;
    uint32_t g = 0, next = 1;
    for (uint32_t chunk_end = n_chunks; chunk_end > 0; g++) {
      uint32_t take = even_groups ? (n_chunks + 7) / 8 : next;
      if (take > chunk_end) take = chunk_end;
      if (g >= 1 && !even_groups) next *= 2;
      const uint32_t col_begin = (chunk_end - take) * 8, count = take * 8;
      CUT(cudaMemcpyAsync(t->d_evals + (size_t)col_begin * block_len, evals + (size_t)col_begin * block_len,
                          (size_t)count * block_len * sizeof(uint32_t), cudaMemcpyHostToDevice, c->copy_stream));
      CUT(cudaEventRecord(c->ev_copy[g % 8], c->copy_stream));
      CUT(cudaStreamWaitEvent(c->stream, c->ev_copy[g % 8], 0));
      CUT(lm::ntt_reorder_and_dft_cols(c->stream, t->d_evals, n_vars, folding, log_inv_rate, (uint32_t)dft_cols, col_begin,
                                       count, t->d_codeword, c->d_tw, c->tw_log_n));
      CUT(lm::merkle_leaf_absorb_chunks(c->stream, t->d_codeword, t->height, stored_w, full_w, eff_w, chunk_end - 1, take,
                                        t->d_layers));
      chunk_end -= take;
    }
  }
  CUT(lm::merkle_tree_from_digests(c->stream, t->d_layers, t->height));
  CUT(cudaMemcpyAsync(out_root, t->d_layers + (2 * t->height - 2) * 8, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                      c->stream));
  CUT(cudaStreamSynchronize(c->stream));
#undef CUT
  *out_tree = t;
  return LM_OK;
}

int lm_commit(lm_ctx* c, const uint32_t* evals, uint32_t n_vars, uint32_t dim, uint64_t actual_len, uint32_t folding,
              uint32_t log_inv_rate, lm_tree** out_tree, uint32_t out_root[8]) {
  if (!evals && actual_len) return fail(LM_ERR_INVALID, "lm_commit: evals is null");
  return commit_impl(c, evals, false, n_vars, dim, actual_len, folding, log_inv_rate, true, out_tree, out_root);
}

int lm_commit_dev(lm_ctx* c, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim, uint64_t actual_len,
                  uint32_t folding, uint32_t log_inv_rate, int retain_evals, lm_tree** out_tree, uint32_t out_root[8]) {
  if (!d_evals) return fail(LM_ERR_INVALID, "lm_commit_dev: d_evals is null");
  return commit_impl(c, d_evals, true, n_vars, dim, actual_len, folding, log_inv_rate, retain_evals != 0, out_tree,
                     out_root);
}

int lm_commit_stacked(lm_ctx* c, const lm_segment* segments, uint32_t n_segments, uint32_t n_vars, uint32_t folding,
                      uint32_t log_inv_rate, lm_tree** out_tree, uint32_t out_root[8]) {
  if (!c || (!segments && n_segments) || !out_tree || !out_root) return fail(LM_ERR_INVALID, "lm_commit_stacked: null argument");
  if (n_vars > 32) return fail(LM_ERR_INVALID, "lm_commit_stacked: n_vars %u too large", n_vars);
  const uint64_t full = (uint64_t)1 << n_vars;
  uint64_t actual = 0;
  for (uint32_t i = 0; i < n_segments; i++) {
    if (segments[i].len && !segments[i].data) return fail(LM_ERR_INVALID, "lm_commit_stacked: segment %u is null", i);
    if (segments[i].offset + segments[i].len > full)
      return fail(LM_ERR_INVALID, "lm_commit_stacked: segment %u ends beyond 2^%u", i, n_vars);
    if (segments[i].offset + segments[i].len > actual) actual = segments[i].offset + segments[i].len;
  }
  CU(cudaSetDevice(c->device));
  // the live prefix is assembled on the device: gaps between segments are zero (stacked_pcs.rs:118-135 starts from zero_vec)
  uint32_t* d_tmp = nullptr;
  const size_t bytes = (actual ? actual : 1) * sizeof(uint32_t);
  CU(c->pool.get(bytes, reinterpret_cast<void**>(&d_tmp)));
  cudaError_t e = cudaMemsetAsync(d_tmp, 0, bytes, c->stream);
  for (uint32_t i = 0; e == cudaSuccess && i < n_segments; i++)
    if (segments[i].len)
      e = cudaMemcpyAsync(d_tmp + segments[i].offset, segments[i].data, segments[i].len * sizeof(uint32_t),
                          cudaMemcpyHostToDevice, c->stream);
  if (e != cudaSuccess) {
    cudaStreamSynchronize(c->stream);
    c->pool.put(bytes, d_tmp);
    return cuda_fail(e, "lm_commit_stacked");
  }
  const int rc = commit_impl(c, d_tmp, true, n_vars, 1, actual, folding, log_inv_rate, true, out_tree, out_root);
  c->pool.put(bytes, d_tmp);  // commit_impl synchronised the stream
  return rc;
}

int lm_access_counts(lm_ctx* c, const uint32_t* const* index_cols, const uint64_t* n_rows, const uint32_t* n_values,
                     uint32_t n_cols, uint64_t table_len, uint32_t* out_acc) {
  if (!c || (n_cols && (!index_cols || !n_rows || !n_values)) || !out_acc)
    return fail(LM_ERR_INVALID, "lm_access_counts: null argument");
  if (table_len == 0) return LM_OK;
  CU(cudaSetDevice(c->device));
  uint64_t max_rows = 1;
  for (uint32_t k = 0; k < n_cols; k++) {
    if (n_rows[k] && !index_cols[k]) return fail(LM_ERR_INVALID, "lm_access_counts: column %u is null", k);
    if (n_rows[k] > max_rows) max_rows = n_rows[k];
  }
  uint32_t *d_counts = nullptr, *d_col = nullptr;
  CU(c->pool.alloc(&d_counts, (table_len + 1) * sizeof(uint32_t)));  // last word: out-of-range flag
  cudaError_t e = c->pool.alloc(&d_col, max_rows * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_counts, 0, (table_len + 1) * sizeof(uint32_t), c->stream);
  for (uint32_t k = 0; e == cudaSuccess && k < n_cols; k++) {
    if (!n_rows[k]) continue;
    e = cudaMemcpyAsync(d_col, index_cols[k], n_rows[k] * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = lm::access_count(c->stream, d_col, n_rows[k], n_values[k], table_len, d_counts, d_counts + table_len);
  }
  uint32_t bad = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_counts + table_len, sizeof(bad), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = lm::counts_to_monty(c->stream, d_counts, table_len);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_acc, d_counts, table_len * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream);
  c->pool.free(d_counts);
  if (d_col) c->pool.free(d_col);
  if (e != cudaSuccess) return cuda_fail(e, "lm_access_counts");
  if (bad) return fail(LM_ERR_INVALID, "lm_access_counts: an address (+ its value count) lies outside the table of %llu entries",
                       (unsigned long long)table_len);
  return LM_OK;
}

int lm_tree_shape(const lm_tree* t, uint64_t* height, uint32_t* full_w, uint32_t* stored_w, uint32_t* dim) {
  if (!t) return fail(LM_ERR_INVALID, "lm_tree_shape: tree is null");
  if (height) *height = t->height;
  if (full_w) *full_w = t->full_width;
  if (stored_w) *stored_w = t->stored_width;
  if (dim) *dim = t->elem_dim;
  return LM_OK;
}

static int open_impl(lm_tree* t, const uint64_t* indices, uint32_t n, uint32_t* out_rows, uint32_t* out_paths,
                     const uint32_t* fold_point, uint32_t fold_vars, uint32_t* out_evals) {
  if (!t || (n && (!indices || !out_rows || !out_paths))) return fail(LM_ERR_INVALID, "lm_open: null argument");
  lm_ctx* c = t->ctx;
  CU(cudaSetDevice(c->device));
  uint32_t log_h = 0;
  while (((uint64_t)1 << log_h) < t->height) log_h++;
  for (uint32_t q = 0; q < n; q++)
    if (indices[q] >= t->height) return fail(LM_ERR_INVALID, "lm_open: index %llu >= height %llu",
                                             (unsigned long long)indices[q], (unsigned long long)t->height);
  if (fold_point && (((uint32_t)1 << fold_vars) * t->elem_dim != t->full_width || fold_vars > 12))
    return fail(LM_ERR_INVALID, "lm_open_fold: a leaf has %u elements, not 2^%u", t->full_width / t->elem_dim, fold_vars);
  const size_t rows_words = (size_t)n * t->full_width, paths_words = (size_t)n * log_h * 8;
  const size_t evals_words = fold_point ? (size_t)n * 5 + (size_t)fold_vars * 5 : 0;
  if (n == 0) return LM_OK;
  uint64_t* d_idx = nullptr;
  uint32_t* d_buf = nullptr;
  CU(t->ctx->pool.alloc(&d_idx, n * sizeof(uint64_t)));
  cudaError_t e = t->ctx->pool.alloc(&d_buf, (rows_words + paths_words + evals_words + 1) * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, indices, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess)
    e = lm::merkle_open_gather(c->stream, t->d_codeword, t->d_layers, t->height, t->stored_width, t->full_width, d_idx,
                               n, d_buf, d_buf + rows_words);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out_rows, d_buf, rows_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && paths_words)
    e = cudaMemcpyAsync(out_paths, d_buf + rows_words, paths_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  if (fold_point) {
    uint32_t* d_ev = d_buf + rows_words + paths_words;
    uint32_t* d_pt = d_ev + (size_t)n * 5;
    if (e == cudaSuccess && fold_vars)
      e = cudaMemcpyAsync(d_pt, fold_point, (size_t)fold_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = lm::rows_mle_eval(c->stream, d_buf, n, t->elem_dim, fold_vars, d_pt, d_ev);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_evals, d_ev, (size_t)n * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream);
  t->ctx->pool.free(d_idx);
  if (d_buf) t->ctx->pool.free(d_buf);
  if (e != cudaSuccess) return cuda_fail(e, "lm_open");
  return LM_OK;
}

int lm_open(lm_tree* t, const uint64_t* indices, uint32_t n, uint32_t* out_rows, uint32_t* out_paths) {
  return open_impl(t, indices, n, out_rows, out_paths, nullptr, 0, nullptr);
}

int lm_open_fold(lm_tree* t, const uint64_t* indices, uint32_t n, const uint32_t* fold_point, uint32_t fold_vars, uint32_t* out_rows,
                 uint32_t* out_paths, uint32_t* out_evals) {
  if (!fold_point || !out_evals) return fail(LM_ERR_INVALID, "lm_open_fold: null argument");
  return open_impl(t, indices, n, out_rows, out_paths, fold_point, fold_vars, out_evals);
}

int lm_verify_openings(lm_ctx* c, const uint32_t root[8], uint32_t log_height, const uint64_t* indices, uint32_t n,
                       const uint32_t* rows, uint32_t width, uint32_t elem_dim, const uint32_t* paths, const uint32_t* fold_point,
                       uint32_t fold_vars, uint8_t* out_ok, uint32_t* out_evals) {
  if (!c || !root || (n && (!indices || !rows || !out_ok || (log_height && !paths))))
    return fail(LM_ERR_INVALID, "lm_verify_openings: null argument");
  if (width < 16 || (width & 7)) return fail(LM_ERR_INVALID, "lm_verify_openings: row width %u is not a multiple of 8 >= 16", width);
  if (log_height > 40) return fail(LM_ERR_INVALID, "lm_verify_openings: log_height %u", log_height);
  const bool fold = fold_point || out_evals;
  if (fold && (!out_evals || (fold_vars && !fold_point) || (elem_dim != 1 && elem_dim != 5) || fold_vars > 12 ||
               ((uint32_t)1 << fold_vars) * elem_dim != width))
    return fail(LM_ERR_INVALID, "lm_verify_openings: a leaf of %u words is not 2^%u elements of dimension %u", width, fold_vars, elem_dim);
  if (n == 0) return LM_OK;
  CU(cudaSetDevice(c->device));
  const size_t rows_words = (size_t)n * width, paths_words = (size_t)n * log_height * 8;
  const size_t evals_words = fold ? (size_t)n * 5 + (size_t)fold_vars * 5 : 0;
  uint64_t* d_idx = nullptr;
  uint32_t* d_buf = nullptr;  // rows | paths | root | ok | evals | point
  CU(c->pool.alloc(&d_idx, n * sizeof(uint64_t)));
  cudaError_t e = c->pool.alloc(&d_buf, (rows_words + paths_words + 8 + n + evals_words + 1) * sizeof(uint32_t));
  uint32_t* d_paths = d_buf + rows_words;
  uint32_t* d_root = d_paths + paths_words;
  uint32_t* d_ok = d_root + 8;
  uint32_t* d_ev = d_ok + n;
  uint32_t* d_pt = d_ev + (size_t)n * 5;
  std::vector<uint32_t> ok(n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, indices, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_buf, rows, rows_words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && paths_words)
    e = cudaMemcpyAsync(d_paths, paths, paths_words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_root, root, 8 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = lm::merkle_verify_openings(c->stream, d_root, log_height, d_idx, n, d_buf, width, d_paths, d_ok);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ok.data(), d_ok, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  if (fold) {
    if (e == cudaSuccess && fold_vars)
      e = cudaMemcpyAsync(d_pt, fold_point, (size_t)fold_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = lm::rows_mle_eval(c->stream, d_buf, n, elem_dim, fold_vars, d_pt, d_ev);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_evals, d_ev, (size_t)n * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream);
  c->pool.free(d_idx);
  if (d_buf) c->pool.free(d_buf);
  if (e != cudaSuccess) return cuda_fail(e, "lm_verify_openings");
  for (uint32_t q = 0; q < n; q++) out_ok[q] = ok[q] ? 1 : 0;
  return LM_OK;
}

int lm_tree_eval(lm_tree* t, const uint32_t* point, uint32_t out[5]) {
  if (!t || !point || !out) return fail(LM_ERR_INVALID, "lm_tree_eval: null argument");
  if (!t->d_evals) return fail(LM_ERR_INVALID, "lm_tree_eval: the polynomial was not retained by this commit");
  lm_ctx* c = t->ctx;
  CU(cudaSetDevice(c->device));
  if (t->n_vars > 64) return fail(LM_ERR_INVALID, "lm_tree_eval: n_vars too large");
  if (t->n_vars)
    CU(cudaMemcpyAsync(c->d_point, point, (size_t)t->n_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  int rc = lm_dev_mle_eval(c, t->d_evals, t->n_vars, t->elem_dim, t->actual_len, c->d_point, c->d_small);
  if (rc != LM_OK) return rc;
  CU(cudaMemcpyAsync(out, c->d_small, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_tree_read_codeword(lm_tree* t, uint32_t* out) {
  if (!t || !out) return fail(LM_ERR_INVALID, "lm_tree_read_codeword: null argument");
  return lm_dev_download(t->ctx, out, t->d_codeword, t->height * t->stored_width * sizeof(uint32_t));
}
int lm_tree_read_layers(lm_tree* t, uint32_t* out) {
  if (!t || !out) return fail(LM_ERR_INVALID, "lm_tree_read_layers: null argument");
  return lm_dev_download(t->ctx, out, t->d_layers, (2 * t->height - 1) * 8 * sizeof(uint32_t));
}

int lm_tree_free(lm_tree* t) {
  if (!t) return LM_OK;
  if (t->ctx) {
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
  }
  if (t->ctx) {
    if (t->d_evals && t->owns_evals) t->ctx->pool.put(t->evals_bytes, t->d_evals);
    t->ctx->pool.put(t->codeword_bytes, t->d_codeword);
    t->ctx->pool.put(t->layers_bytes, t->d_layers);
  } else {
    if (t->d_evals && t->owns_evals) cudaFree(t->d_evals);
    if (t->d_codeword) cudaFree(t->d_codeword);
    if (t->d_layers) cudaFree(t->d_layers);
  }
  delete t;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------ WHIR open session
int lm_sc_free(lm_sumcheck* s) {
  if (!s) return LM_OK;
  if (s->ctx) {
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
  }
  if (s->d_p_owned) s->ctx->pool.free(s->d_p_owned);
  if (s->d_w) s->ctx->pool.free(s->d_w);
  if (s->d_out10) s->ctx->pool.free(s->d_out10);
  if (s->d_scratch) s->ctx->pool.free(s->d_scratch);
  delete s;
  return LM_OK;
}

static int sc_new_common(lm_ctx* c, uint32_t n_vars, lm_sumcheck** out) {
  if (!c || !out) return fail(LM_ERR_INVALID, "lm_sc_new: null argument");
  *out = nullptr;
  if (n_vars < 1 || n_vars > 32) return fail(LM_ERR_INVALID, "lm_sc_new: n_vars %u out of range", n_vars);
  CU(cudaSetDevice(c->device));
  lm_sumcheck* s = new (std::nothrow) lm_sumcheck();
  if (!s) return fail(LM_ERR_OOM, "lm_sc_new: host allocation failed");
  s->ctx = c;
  s->n_vars = n_vars;
  const size_t w_bytes = ((size_t)5 << n_vars) * sizeof(uint32_t);
  cudaError_t e = s->ctx->pool.alloc(&s->d_w, w_bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_w, 0, w_bytes, c->stream);
  if (e == cudaSuccess) e = s->ctx->pool.alloc(&s->d_out10, 16 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    lm_sc_free(s);
    return cuda_fail(e, "lm_sc_new");
  }
  int rc = s->ensure_scratch(lm::prod_round_scratch_words());
  if (rc != LM_OK) {
    lm_sc_free(s);
    return rc;
  }
  *out = s;
  return LM_OK;
}

int lm_sc_new_from_tree(lm_tree* t, lm_sumcheck** out) {
  if (!t) return fail(LM_ERR_INVALID, "lm_sc_new_from_tree: tree is null");
  if (!t->d_evals) return fail(LM_ERR_INVALID, "lm_sc_new_from_tree: the polynomial was not retained by this commit");
  int rc = sc_new_common(t->ctx, t->n_vars, out);
  if (rc != LM_OK) return rc;
  (*out)->d_p = t->d_evals;
  (*out)->p_dim = t->elem_dim;
  (*out)->p_live = t->actual_len;
  return LM_OK;
}

int lm_sc_new(lm_ctx* c, const uint32_t* evals, uint32_t n_vars, uint32_t dim, uint64_t live_len, lm_sumcheck** out) {
  if (!evals && live_len) return fail(LM_ERR_INVALID, "lm_sc_new: evals is null");
  if (dim != 1 && dim != 5) return fail(LM_ERR_INVALID, "lm_sc_new: elem_dim must be 1 or 5");
  if (n_vars > 32 || live_len > ((uint64_t)1 << n_vars)) return fail(LM_ERR_INVALID, "lm_sc_new: live_len exceeds 2^n_vars");
  int rc = sc_new_common(c, n_vars, out);
  if (rc != LM_OK) return rc;
  lm_sumcheck* s = *out;
  cudaError_t e = s->ctx->pool.alloc(&s->d_p_owned, (live_len ? live_len : 1) * dim * sizeof(uint32_t));
  if (e == cudaSuccess && live_len)
    e = cudaMemcpyAsync(s->d_p_owned, evals, live_len * dim * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    lm_sc_free(s);
    *out = nullptr;
    return cuda_fail(e, "lm_sc_new upload");
  }
  s->d_p = s->d_p_owned;
  s->p_dim = dim;
  s->p_live = live_len;
  return LM_OK;
}

int lm_sc_new_from_dev(lm_ctx* c, const uint32_t* d_poly, const uint32_t* d_weights, uint32_t n_vars, lm_sumcheck** out) {
  if (!d_poly || !d_weights) return fail(LM_ERR_INVALID, "lm_sc_new_from_dev: null argument");
  int rc = sc_new_common(c, n_vars, out);
  if (rc != LM_OK) return rc;
  lm_sumcheck* s = *out;
  const size_t bytes = ((size_t)5 << n_vars) * sizeof(uint32_t);
  cudaError_t e = s->ctx->pool.alloc(&s->d_p_owned, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_p_owned, d_poly, bytes, cudaMemcpyDeviceToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_w, d_weights, bytes, cudaMemcpyDeviceToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    lm_sc_free(s);
    *out = nullptr;
    return cuda_fail(e, "lm_sc_new_from_dev");
  }
  s->d_p = s->d_p_owned;
  s->p_dim = 5;
  s->p_live = (uint64_t)1 << n_vars;
  return LM_OK;
}

int lm_sc_export_dev(lm_sumcheck* s, uint32_t* d_poly_out, uint32_t* d_weights_out) {
  if (!s || !d_poly_out || !d_weights_out) return fail(LM_ERR_INVALID, "lm_sc_export_dev: null argument");
  if (s->p_dim != 5) return fail(LM_ERR_INVALID, "lm_sc_export_dev: the polynomial has not been folded yet (base field)");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  const size_t bytes = ((size_t)5 << s->n_vars) * sizeof(uint32_t);
  CU(cudaMemcpyAsync(d_poly_out, s->d_p, bytes, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaMemcpyAsync(d_weights_out, s->d_w, bytes, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

static int sc_upload_point(lm_sumcheck* s, const uint32_t* point, uint32_t words) {
  lm_ctx* c = s->ctx;
  if (words > 64 * 5) return fail(LM_ERR_INVALID, "point too long");
  if (words) CU(cudaMemcpyAsync(c->d_point, point, words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  return LM_OK;
}

int lm_sc_add_eq(lm_sumcheck* s, uint64_t selector, const uint32_t* point, uint32_t m, const uint32_t scalar[5]) {
  if (!s || !scalar || (m && !point)) return fail(LM_ERR_INVALID, "lm_sc_add_eq: null argument");
  if (m > s->n_vars || (selector >> (s->n_vars - m)) != 0)
    return fail(LM_ERR_INVALID, "lm_sc_add_eq: selector %llu does not fit %u - %u variables", (unsigned long long)selector,
                s->n_vars, m);
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  int rc = s->ensure_scratch(lm::weights_add_eq_scratch_words(m) + lm::prod_round_scratch_words());
  if (rc != LM_OK) return rc;
  if ((rc = sc_upload_point(s, point, 5 * m)) != LM_OK) return rc;
  CU(lm::weights_add_eq(c->stream, s->d_w, selector, c->d_point, m, scalar, s->d_scratch));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_sc_add_eq_batch(lm_sumcheck* s, uint64_t selector, const uint32_t* points, uint32_t m, const uint32_t* scalars, uint32_t n_st) {
  if (!s || !scalars || !points) return fail(LM_ERR_INVALID, "lm_sc_add_eq_batch: null argument");
  if (m < 1 || m > s->n_vars || (selector >> (s->n_vars - m)) != 0 || n_st == 0)
    return fail(LM_ERR_INVALID, "lm_sc_add_eq_batch: bad selector / point length / count");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  const uint32_t k_max = lm::weights_add_eq_batch_max(m);  // statements per pass (table footprint / accumulator bounds)
  for (uint32_t k0 = 0; k0 < n_st; k0 += k_max) {
    const uint32_t K = n_st - k0 < k_max ? n_st - k0 : k_max;
    const size_t tab = lm::weights_add_eq_batch_scratch_words(m, K);
    int rc = s->ensure_scratch(tab + (size_t)K * m * 5 + lm::prod_round_scratch_words());
    if (rc != LM_OK) return rc;
    uint32_t* d_pts = s->d_scratch + tab;
    CU(cudaMemcpyAsync(d_pts, points + (size_t)k0 * m * 5, (size_t)K * m * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    CU(lm::weights_add_eq_batch(c->stream, s->d_w, selector, d_pts, m, scalars + 5 * k0, K, s->d_scratch));
    CU(cudaStreamSynchronize(c->stream));
  }
  return LM_OK;
}

int lm_sc_add_next(lm_sumcheck* s, uint64_t selector, const uint32_t* point, uint32_t m, const uint32_t scalar[5]) {
  if (!s || !scalar || !point) return fail(LM_ERR_INVALID, "lm_sc_add_next: null argument");
  if (m < 1 || m > s->n_vars || (selector >> (s->n_vars - m)) != 0)
    return fail(LM_ERR_INVALID, "lm_sc_add_next: bad selector / point length");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  int rc = sc_upload_point(s, point, 5 * m);
  if (rc != LM_OK) return rc;
  CU(lm::weights_add_next(c->stream, s->d_w, selector, c->d_point, m, scalar));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_sc_add_strided_eq(lm_sumcheck* s, uint64_t base, uint32_t shift, uint64_t offset, const uint32_t* point, uint32_t pre,
                         const uint32_t scalar[5]) {
  if (!s || !scalar || (pre && !point)) return fail(LM_ERR_INVALID, "lm_sc_add_strided_eq: null argument");
  const uint64_t n = (uint64_t)1 << s->n_vars;
  if (pre > s->n_vars || shift > 63 || pre + shift > 63 || offset >= n || base >= n ||
      base + ((((uint64_t)1 << pre) - 1) << shift) + offset >= n)
    return fail(LM_ERR_INVALID, "lm_sc_add_strided_eq: index set leaves the table");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  if (pre) {
    int rc = sc_upload_point(s, point, 5 * pre);
    if (rc != LM_OK) return rc;
  }
  CU(lm::weights_add_strided_eq(c->stream, s->d_w, base, shift, offset, c->d_point, pre, scalar));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_sc_add_base_eq(lm_sumcheck* s, const uint32_t* points, uint32_t n_q, const uint32_t* scalars) {
  if (!s || (n_q && (!points || !scalars))) return fail(LM_ERR_INVALID, "lm_sc_add_base_eq: null argument");
  if (n_q == 0) return LM_OK;
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  const uint32_t m = s->n_vars;
  const size_t tab_words = lm::weights_add_base_eq_scratch_words(m, n_q);
  int rc = s->ensure_scratch(tab_words + (size_t)n_q * (m + 5) + lm::prod_round_scratch_words());
  if (rc != LM_OK) return rc;
  uint32_t* d_pts = s->d_scratch + tab_words;
  uint32_t* d_sc = d_pts + (size_t)n_q * m;
  CU(cudaMemcpyAsync(d_pts, points, (size_t)n_q * m * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_sc, scalars, (size_t)n_q * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(lm::weights_add_base_eq(c->stream, s->d_w, m, d_pts, n_q, d_sc, s->d_scratch));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

static int sc_fetch_round(lm_sumcheck* s, uint32_t c0[5], uint32_t c2[5]) {
  uint32_t h[10];
  CU(cudaMemcpyAsync(h, s->d_out10, sizeof(h), cudaMemcpyDeviceToHost, s->ctx->stream));
  CU(cudaStreamSynchronize(s->ctx->stream));
  memcpy(c0, h, 5 * sizeof(uint32_t));
  memcpy(c2, h + 5, 5 * sizeof(uint32_t));
  return LM_OK;
}

int lm_sc_round(lm_sumcheck* s, uint32_t c0[5], uint32_t c2[5]) {
  if (!s || !c0 || !c2) return fail(LM_ERR_INVALID, "lm_sc_round: null argument");
  if (s->n_vars < 1) return fail(LM_ERR_INVALID, "lm_sc_round: no variables left");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  CU(lm::prod_round(c->stream, s->d_p, s->p_dim, s->p_live, s->d_w, (uint64_t)1 << s->n_vars, s->d_scratch, s->d_out10));
  return sc_fetch_round(s, c0, c2);
}

// make room for the EF polynomial table the first time a base-field polynomial is folded
static int sc_prepare_fold_target(lm_sumcheck* s, uint32_t** p_out) {
  const uint64_t half = (uint64_t)1 << (s->n_vars - 1);
  if (s->p_dim == 5 && s->d_p == s->d_p_owned) {
    *p_out = s->d_p_owned;  // EF table owned by us: in place
    return LM_OK;
  }
  uint32_t* fresh = nullptr;
  CU(s->ctx->pool.alloc(&fresh, half * 5 * sizeof(uint32_t)));
  *p_out = fresh;
  return LM_OK;
}
static void sc_commit_fold(lm_sumcheck* s, uint32_t* p_out) {
  if (p_out != s->d_p_owned) {
    if (s->d_p_owned) {
      cudaStreamSynchronize(s->ctx->stream);
      s->ctx->pool.free(s->d_p_owned);
    }
    s->d_p_owned = p_out;
  }
  s->d_p = s->d_p_owned;
  s->p_dim = 5;
  s->n_vars -= 1;
  s->p_live = (uint64_t)1 << s->n_vars;
}

int lm_sc_fold(lm_sumcheck* s, const uint32_t r[5]) {
  if (!s || !r) return fail(LM_ERR_INVALID, "lm_sc_fold: null argument");
  if (s->n_vars < 1) return fail(LM_ERR_INVALID, "lm_sc_fold: no variables left");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  uint32_t* p_out = nullptr;
  int rc = sc_prepare_fold_target(s, &p_out);
  if (rc != LM_OK) return rc;
  const uint64_t n = (uint64_t)1 << s->n_vars;
  CU(lm::fold_msb(c->stream, s->d_p, n, s->p_dim, s->p_live, r, p_out));
  CU(lm::fold_msb(c->stream, s->d_w, n, 5, n, r, s->d_w));
  sc_commit_fold(s, p_out);
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_sc_fold_round(lm_sumcheck* s, const uint32_t r[5], uint32_t c0[5], uint32_t c2[5]) {
  if (!s || !r || !c0 || !c2) return fail(LM_ERR_INVALID, "lm_sc_fold_round: null argument");
  if (s->n_vars < 2) return fail(LM_ERR_INVALID, "lm_sc_fold_round: needs at least two variables");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  uint32_t* p_out = nullptr;
  int rc = sc_prepare_fold_target(s, &p_out);
  if (rc != LM_OK) return rc;
  const uint64_t n = (uint64_t)1 << s->n_vars;
  CU(lm::prod_fold_round(c->stream, s->d_p, s->p_dim, s->p_live, s->d_w, n, r, p_out, s->d_w, s->d_scratch, s->d_out10));
  sc_commit_fold(s, p_out);
  return sc_fetch_round(s, c0, c2);
}

int lm_sc_num_vars(const lm_sumcheck* s, uint32_t* n_vars, uint32_t* poly_dim) {
  if (!s) return fail(LM_ERR_INVALID, "lm_sc_num_vars: null argument");
  if (n_vars) *n_vars = s->n_vars;
  if (poly_dim) *poly_dim = s->p_dim;
  return LM_OK;
}

int lm_sc_read(lm_sumcheck* s, uint32_t* out_poly, uint32_t* out_weights) {
  if (!s) return fail(LM_ERR_INVALID, "lm_sc_read: null argument");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  const uint64_t n = (uint64_t)1 << s->n_vars;
  if (out_poly) {
    memset(out_poly, 0, n * s->p_dim * sizeof(uint32_t));
    CU(cudaMemcpyAsync(out_poly, s->d_p, s->p_live * s->p_dim * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  if (out_weights) CU(cudaMemcpyAsync(out_weights, s->d_w, n * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_sc_eval_poly(lm_sumcheck* s, const uint32_t* point, uint32_t out[5]) {
  if (!s || !point || !out) return fail(LM_ERR_INVALID, "lm_sc_eval_poly: null argument");
  lm_ctx* c = s->ctx;
  CU(cudaSetDevice(c->device));
  int rc = sc_upload_point(s, point, 5 * s->n_vars);
  if (rc != LM_OK) return rc;
  // mle_eval reads whole rows of min(2^n, 1024) entries: only valid when the table is fully materialised
  const uint64_t n = (uint64_t)1 << s->n_vars;
  const uint64_t row = n < 1024 ? n : 1024;
  if (s->p_live % row != 0 && s->p_live != n)
    return fail(LM_ERR_INVALID, "lm_sc_eval_poly: live prefix must be a multiple of %llu", (unsigned long long)row);
  rc = lm_dev_mle_eval(c, s->d_p, s->n_vars, s->p_dim, s->p_live, c->d_point, c->d_small);
  if (rc != LM_OK) return rc;
  CU(cudaMemcpyAsync(out, c->d_small, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_sc_commit_poly(lm_sumcheck* s, uint32_t folding_factor, uint32_t log_inv_rate, lm_tree** out_tree, uint32_t out_root[8]) {
  if (!s) return fail(LM_ERR_INVALID, "lm_sc_commit_poly: null argument");
  return commit_impl(s->ctx, s->d_p, true, s->n_vars, s->p_dim, s->p_live, folding_factor, log_inv_rate, false, out_tree,
                     out_root);
}

// ------------------------------------------------------------------------------------------ AIR sumcheck session
int lm_air_free(lm_air* a) {
  if (!a) return LM_OK;
  if (a->ctx) {
    cudaSetDevice(a->ctx->device);
    cudaStreamSynchronize(a->ctx->stream);
  }
  // the table-sized buffers go back to the context's size-keyed cache: the next session of the same shape reuses them
  // (cudaMalloc / cudaFree of GiB buffers cost milliseconds each and used to sit inside the first rounds)
  auto give_back = [&](uint32_t* p, size_t words) {
    if (!p) return;
    if (a->ctx) a->ctx->pool.put(words * sizeof(uint32_t), p);
    else cudaFree(p);
  };
  if (!a->borrowed) give_back(a->d_cols, a->cols_words);
  give_back(a->d_spare, a->spare_words);
  give_back(a->d_ef[0], a->ef_words[0]);
  give_back(a->d_ef[1], a->ef_words[1]);
  if (a->d_dev) cudaFree(a->d_dev);
  if (a->d_eq_tab) cudaFree(a->d_eq_tab);
  if (a->d_scratch) cudaFree(a->d_scratch);
  if (a->d_out) cudaFree(a->d_out);
  delete a;
  return LM_OK;
}

// cols: base-field host columns (the shifted ones are derived, their last row taken from halo_next_row when the
// session covers a row range that is not the last one), or cols_ef: all n_cols + n_shift columns already in EF.
static int air_new_impl(lm_ctx* c, uint32_t table_id, const uint32_t* const* cols, const uint32_t* cols_ef, uint32_t n_cols,
                        uint32_t log_rows, const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha,
                        const uint32_t* logup_alphas_eq, uint32_t n_la, const uint32_t bus_beta[5],
                        const uint32_t* halo_next_row, const uint32_t* eq_scale, lm_air** out, const uint32_t* d_cols_in = nullptr) {
  if (!c || (!cols && !cols_ef && !d_cols_in) || !eq_factor || !alpha_powers || !bus_beta || !out || (!logup_alphas_eq && n_la))
    return fail(LM_ERR_INVALID, "lm_air_new: null argument");
  *out = nullptr;
  uint32_t t_cols, t_shift, t_deg, t_maxc;
  if (!lm::air_table_shape(table_id, &t_cols, &t_shift, &t_deg, &t_maxc) || (table_id & ~0x1ffu) ||
      ((table_id & 0x100u) && (table_id & 0xffu) == 0))
    return fail(LM_ERR_INVALID, "lm_air_new: unknown table id 0x%x (0 execution, 1 extension_op, 2 poseidon16, | 0x100 no bus)",
                table_id);
  const bool bus = !(table_id & 0x100u);
  if (n_cols != t_cols) return fail(LM_ERR_INVALID, "lm_air_new: table %u has %u columns, got %u", table_id & 0xffu, t_cols, n_cols);
  if (n_alpha < t_maxc - (bus ? 0 : 1))
    return fail(LM_ERR_INVALID, "lm_air_new: need >= %u alpha powers, got %u", t_maxc - (bus ? 0 : 1), n_alpha);
  if (bus && n_la < 5) return fail(LM_ERR_INVALID, "lm_air_new: need >= 5 logup alphas, got %u", n_la);
  if (log_rows < 1 || log_rows > 30) return fail(LM_ERR_INVALID, "lm_air_new: log_rows %u out of range", log_rows);
  CU(cudaSetDevice(c->device));
  lm_air* a = new (std::nothrow) lm_air();
  if (!a) return fail(LM_ERR_OOM, "lm_air_new: host allocation failed");
  a->ctx = c;
  a->table_id = table_id;
  a->n_cols = t_cols;
  a->n_shift = t_shift;
  a->degree = t_deg;
  a->n_vars0 = a->log_n = log_rows;
  a->exec = (table_id & 0xffu) == 0;
  a->alpha.assign(alpha_powers, alpha_powers + 5 * (size_t)n_alpha);
  a->la.assign(logup_alphas_eq, logup_alphas_eq + 5 * (size_t)n_la);
  memcpy(a->beta, bus_beta, sizeof(a->beta));
  const uint64_t n = (uint64_t)1 << log_rows;
  const uint32_t all = a->n_cols + a->n_shift;
  a->dim = cols_ef ? 5 : 1;
  cudaError_t e = cudaSuccess;
  if (cols_ef) {
    // [all][n][5] from the host -> planes; for the execution table this is the extension-field source table
    uint32_t* tmp = nullptr;
    const size_t words = (size_t)all * n * 5;
    e = cudaMalloc(&tmp, words * sizeof(uint32_t));
    uint32_t* planes = nullptr;
    if (e == cudaSuccess) e = c->pool.get(words * sizeof(uint32_t), reinterpret_cast<void**>(&planes));
    if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, cols_ef, words * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = lm::air_aos_to_planes(c->stream, tmp, all, n, planes);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (tmp) cudaFree(tmp);
    if (a->exec) {
      a->d_ef[0] = planes, a->ef_words[0] = words;
      a->src_is_base = false, a->src_buf = 0, a->src_log_rows = log_rows, a->pending = 0;
    } else {
      a->d_cols = planes, a->cols_words = words;
    }
  } else if (d_cols_in) {
    // device-resident base columns: borrowed by the execution table, copied (next to their shifts) by the wide tables
    if (a->exec) {
      a->d_cols = const_cast<uint32_t*>(d_cols_in);
      a->borrowed = true;
      a->cols_words = (size_t)a->n_cols * n;
      uint32_t last[2] = {0, 0};
      for (uint32_t k = 0; e == cudaSuccess && k < 2; k++)
        e = cudaMemcpyAsync(&last[k], d_cols_in + (size_t)k * n + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      a->halo[0] = last[0], a->halo[1] = last[1];
      a->src_is_base = true, a->src_log_rows = log_rows, a->pending = 0;
      for (int b = 0; b < 2 && e == cudaSuccess; b++) {
        if (log_rows < 3u + b) break;
        const size_t words = (size_t)22 * 5 * (n >> (2 + b));
        e = c->pool.get(words * sizeof(uint32_t), reinterpret_cast<void**>(&a->d_ef[b]));
        if (e == cudaSuccess) a->ef_words[b] = words;
      }
    } else {
      a->cols_words = (size_t)all * n;
      e = c->pool.get(a->cols_words * sizeof(uint32_t), reinterpret_cast<void**>(&a->d_cols));
      if (e != cudaSuccess) a->cols_words = 0;
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(a->d_cols, d_cols_in, (size_t)a->n_cols * n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream);
      for (uint32_t k = 0; e == cudaSuccess && k < a->n_shift; k++)
        e = lm::air_shift_column(c->stream, a->d_cols + (size_t)k * n, n, a->d_cols + (size_t)(a->n_cols + k) * n);
    }
  } else {
    for (uint32_t k = 0; k < a->n_cols; k++)
      if (!cols[k]) {
        lm_air_free(a);
        return fail(LM_ERR_INVALID, "lm_air_new: column %u is null", k);
      }
    const uint32_t stored = a->exec ? a->n_cols : all;  // the execution table derives its shifted columns on the fly
    a->cols_words = (size_t)stored * n;
    e = c->pool.get(a->cols_words * sizeof(uint32_t), reinterpret_cast<void**>(&a->d_cols));
    if (e != cudaSuccess) a->cols_words = 0;
    if (a->exec && e == cudaSuccess) {
      // both extension tables of the fused rounds (a quarter and an eighth of the rows), so that no round allocates
      for (int b = 0; b < 2 && e == cudaSuccess; b++) {
        if (log_rows < 3u + b) break;
        const size_t words = (size_t)22 * 5 * (n >> (2 + b));
        e = c->pool.get(words * sizeof(uint32_t), reinterpret_cast<void**>(&a->d_ef[b]));
        if (e == cudaSuccess) a->ef_words[b] = words;
      }
    }
    for (uint32_t k = 0; e == cudaSuccess && k < a->n_cols; k++)
      e = cudaMemcpyAsync(a->d_cols + (size_t)k * n, cols[k], n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (a->exec) {
      for (uint32_t k = 0; k < 2; k++) a->halo[k] = halo_next_row ? halo_next_row[k] : cols[k][n - 1];
      a->src_is_base = true, a->src_log_rows = log_rows, a->pending = 0;
    } else {
      // shifted copies of the first n_shift columns (compute_shifted_columns, air_sumcheck.rs:683-694)
      for (uint32_t k = 0; e == cudaSuccess && k < a->n_shift; k++) {
        e = lm::air_shift_column(c->stream, a->d_cols + (size_t)k * n, n, a->d_cols + (size_t)(a->n_cols + k) * n);
        if (e == cudaSuccess && halo_next_row)
          e = cudaMemcpyAsync(a->d_cols + (size_t)(a->n_cols + k) * n + (n - 1), halo_next_row + k, sizeof(uint32_t),
                              cudaMemcpyHostToDevice, c->stream);
      }
    }
  }
  uint32_t* d_eq = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_eq, (size_t)log_rows * 5 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_eq, eq_factor, (size_t)log_rows * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMalloc(&a->d_eq_tab, lm::eqtab_words(log_rows) * sizeof(uint32_t));
  if (e == cudaSuccess) e = lm::air_build_eq_tables(c->stream, d_eq, log_rows, eq_scale, a->d_eq_tab);
  if (e == cudaSuccess) e = cudaMalloc(&a->d_scratch, lm::air_round_scratch_words() * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&a->d_out, 64 * 5 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&a->d_dev, sizeof(lm::AirDev));
  if (e == cudaSuccess) e = cudaMemsetAsync(a->d_dev, 0, sizeof(lm::AirDev), c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (d_eq) cudaFree(d_eq);
  if (e != cudaSuccess) {
    lm_air_free(a);
    return cuda_fail(e, "lm_air_new");
  }
  *out = a;
  return LM_OK;
}

int lm_air_new(lm_ctx* c, uint32_t table_id, const uint32_t* const* cols, uint32_t n_cols, uint32_t log_rows,
               const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha, const uint32_t* logup_alphas_eq,
               uint32_t n_la, const uint32_t bus_beta[5], lm_air** out) {
  if (!cols) return fail(LM_ERR_INVALID, "lm_air_new: null argument");
  return air_new_impl(c, table_id, cols, nullptr, n_cols, log_rows, eq_factor, alpha_powers, n_alpha, logup_alphas_eq, n_la,
                      bus_beta, nullptr, nullptr, out);
}

int lm_air_new_dev(lm_ctx* c, uint32_t table_id, const uint32_t* d_cols, uint32_t n_cols, uint32_t log_rows,
                   const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha, const uint32_t* logup_alphas_eq,
                   uint32_t n_la, const uint32_t bus_beta[5], lm_air** out) {
  if (!d_cols) return fail(LM_ERR_INVALID, "lm_air_new_dev: null argument");
  return air_new_impl(c, table_id, nullptr, nullptr, n_cols, log_rows, eq_factor, alpha_powers, n_alpha, logup_alphas_eq, n_la,
                      bus_beta, nullptr, nullptr, out, d_cols);
}

int lm_air_new_shard(lm_ctx* c, uint32_t table_id, const uint32_t* const* cols, uint32_t n_cols, uint32_t log_rows,
                     const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha,
                     const uint32_t* logup_alphas_eq, uint32_t n_la, const uint32_t bus_beta[5],
                     const uint32_t* halo_next_row, const uint32_t eq_scale[5], lm_air** out) {
  if (!cols) return fail(LM_ERR_INVALID, "lm_air_new_shard: null argument");
  return air_new_impl(c, table_id, cols, nullptr, n_cols, log_rows, eq_factor, alpha_powers, n_alpha, logup_alphas_eq, n_la,
                      bus_beta, halo_next_row, eq_scale, out);
}

int lm_air_new_folded(lm_ctx* c, uint32_t table_id, const uint32_t* cols_ef, uint32_t n_cols_total, uint32_t log_rows,
                      const uint32_t* eq_factor, const uint32_t* alpha_powers, uint32_t n_alpha,
                      const uint32_t* logup_alphas_eq, uint32_t n_la, const uint32_t bus_beta[5], lm_air** out) {
  if (!cols_ef) return fail(LM_ERR_INVALID, "lm_air_new_folded: null argument");
  uint32_t t_cols, t_shift, t_deg, t_maxc;
  if (!lm::air_table_shape(table_id, &t_cols, &t_shift, &t_deg, &t_maxc))
    return fail(LM_ERR_INVALID, "lm_air_new_folded: unknown table id 0x%x", table_id);
  if (n_cols_total != t_cols + t_shift)
    return fail(LM_ERR_INVALID, "lm_air_new_folded: table %u has %u columns incl. shifted ones, got %u", table_id & 0xffu,
                t_cols + t_shift, n_cols_total);
  return air_new_impl(c, table_id, nullptr, cols_ef, t_cols, log_rows, eq_factor, alpha_powers, n_alpha, logup_alphas_eq,
                      n_la, bus_beta, nullptr, nullptr, out);
}

int lm_dev_poseidon16_fill_trace(lm_ctx* c, uint32_t* d_cols, uint64_t n_rows) {
  if (!c || !d_cols) return fail(LM_ERR_INVALID, "lm_dev_poseidon16_fill_trace: null argument");
  CU(cudaSetDevice(c->device));
  CU(lm::poseidon16_fill_trace(c->stream, d_cols, n_rows));
  return LM_OK;
}

int lm_poseidon16_fill_trace(lm_ctx* c, uint32_t* const* cols, uint64_t n_rows) {
  if (!c || !cols) return fail(LM_ERR_INVALID, "lm_poseidon16_fill_trace: null argument");
  for (int k = 0; k < 109; k++)
    if (!cols[k]) return fail(LM_ERR_INVALID, "lm_poseidon16_fill_trace: column %d is null", k);
  if (n_rows == 0) return LM_OK;
  CU(cudaSetDevice(c->device));
  uint32_t* d = nullptr;
  CU(c->pool.alloc(&d, (size_t)109 * n_rows * sizeof(uint32_t)));
  cudaError_t e = cudaSuccess;
  // inputs: flag_permute (column 8) and the 16 input lanes (columns 9..24)
  for (int k = 8; e == cudaSuccess && k < 25; k++)
    e = cudaMemcpyAsync(d + (size_t)k * n_rows, cols[k], n_rows * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = lm::poseidon16_fill_trace(c->stream, d, n_rows);
  for (int k = 25; e == cudaSuccess && k < 109; k++)
    e = cudaMemcpyAsync(cols[k], d + (size_t)k * n_rows, n_rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  c->pool.free(d);
  if (e != cudaSuccess) return cuda_fail(e, "lm_poseidon16_fill_trace");
  return LM_OK;
}

int lm_air_info(const lm_air* a, uint32_t* n_vars, uint32_t* degree, uint32_t* n_cols_total) {
  if (!a) return fail(LM_ERR_INVALID, "lm_air_info: null argument");
  if (n_vars) *n_vars = a->log_n;
  if (degree) *degree = a->degree;
  if (n_cols_total) *n_cols_total = a->n_cols + a->n_shift;
  return LM_OK;
}

// execution table: which kernel variant reads the current source table, and what it needs (air.h AirExecMode)
static int air_exec_args(lm_air* a, lm::AirExecArgs* A, int* mode, bool for_final) {
  const uint32_t need = for_final ? 0u : 1u;  // a round needs one unbound variable on top of the pending ones
  if (a->src_log_rows < a->pending + need) return fail(LM_ERR_INVALID, "lm_air: no variables left");
  *A = lm::AirExecArgs{};
  A->k = a->n_vars0;
  A->eq_tab = a->d_eq_tab;
  A->partial = a->d_scratch;
  A->d = a->d_dev;
  A->m = for_final ? 0 : a->src_log_rows - a->pending - 1;
  if (a->src_is_base) {
    A->base = a->d_cols;
    A->n_base = (uint64_t)1 << a->src_log_rows;
    A->halo[0] = a->halo[0], A->halo[1] = a->halo[1];
    *mode = a->pending == 0 ? lm::AIR_B0 : (a->pending == 1 ? lm::AIR_B1 : lm::AIR_B2);
  } else {
    A->src = a->d_ef[a->src_buf];
    *mode = a->pending == 0 ? lm::AIR_E0 : lm::AIR_E1;
  }
  if (!for_final && (*mode == lm::AIR_B2 || *mode == lm::AIR_E1)) {
    const int dst = a->src_is_base ? 0 : (a->src_buf ^ 1);
    const size_t words = (size_t)22 * 5 * ((size_t)2 << A->m);
    if (a->ef_words[dst] < words) {
      if (a->d_ef[dst]) a->ctx->pool.put(a->ef_words[dst] * sizeof(uint32_t), a->d_ef[dst]);
      a->d_ef[dst] = nullptr, a->ef_words[dst] = 0;
      CU(a->ctx->pool.get(words * sizeof(uint32_t), reinterpret_cast<void**>(&a->d_ef[dst])));
      a->ef_words[dst] = words;
    }
    A->dst = a->d_ef[dst];
  }
  return LM_OK;
}

int lm_air_round(lm_air* a, uint32_t* out_evals) {
  if (!a || !out_evals) return fail(LM_ERR_INVALID, "lm_air_round: null argument");
  if (a->log_n < 1) return fail(LM_ERR_INVALID, "lm_air_round: no variables left");
  if (a->round_done) return fail(LM_ERR_INVALID, "lm_air_round: the previous round has not been folded (lm_air_fold)");
  lm_ctx* c = a->ctx;
  CU(cudaSetDevice(c->device));
  if (a->exec) {
    lm::AirExecArgs A;
    int mode = 0;
    if (int rc = air_exec_args(a, &A, &mode, false)) return rc;
    CU(lm::air_exec_round(c->stream, mode, A, a->alpha.data(), a->la.data(), (uint32_t)(a->la.size() / 5), a->beta, a->r_host));
    if (mode == lm::AIR_B2 || mode == lm::AIR_E1) {  // the round materialised the folded table: it is the source from now on
      a->src_buf = a->src_is_base ? 0 : (a->src_buf ^ 1);
      a->src_log_rows -= a->pending;
      a->src_is_base = false;
      a->pending = 0;
    }
    CU(cudaMemcpyAsync(out_evals, reinterpret_cast<uint8_t*>(a->d_dev) + offsetof(lm::AirDev, out),
                       (size_t)a->degree * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  } else {
    CU(lm::air_generic_round(c->stream, a->table_id, a->d_cols, a->dim, a->log_n, a->d_eq_tab, a->n_vars0, a->alpha.data(),
                             (uint32_t)(a->alpha.size() / 5), a->la.data(), (uint32_t)(a->la.size() / 5), a->beta, a->d_scratch,
                             a->d_out));
    CU(cudaMemcpyAsync(out_evals, a->d_out, (size_t)a->degree * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  a->round_done = true;
  return LM_OK;
}

int lm_air_fold(lm_air* a, const uint32_t r[5]) {
  if (!a || !r) return fail(LM_ERR_INVALID, "lm_air_fold: null argument");
  if (a->log_n < 1) return fail(LM_ERR_INVALID, "lm_air_fold: no variables left");
  lm_ctx* c = a->ctx;
  CU(cudaSetDevice(c->device));
  if (a->exec) {
    // deferred: the next round (or lm_air_final) applies the challenge while it reads the table
    if (!a->round_done) return fail(LM_ERR_INVALID, "lm_air_fold: call lm_air_round first (the fold is fused into the next round)");
    uint8_t* d = reinterpret_cast<uint8_t*>(a->d_dev) + offsetof(lm::AirDev, r);
    CU(cudaMemcpyAsync(d + sizeof(lm::Ef), d, sizeof(lm::Ef), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(d, r, sizeof(lm::Ef), cudaMemcpyHostToDevice, c->stream));
    memcpy(a->r_host, r, sizeof(a->r_host));
    a->pending += 1;
  } else {
    const uint64_t n = (uint64_t)1 << a->log_n;
    const uint32_t all = a->n_cols + a->n_shift;
    const size_t need = (size_t)all * (n / 2) * 5;
    if (a->spare_words < need) {
      if (a->d_spare) c->pool.put(a->spare_words * sizeof(uint32_t), a->d_spare);
      a->d_spare = nullptr;
      a->spare_words = 0;
      CU(c->pool.get(need * sizeof(uint32_t), reinterpret_cast<void**>(&a->d_spare)));
      a->spare_words = need;
    }
    CU(lm::air_fold_lsb(c->stream, a->d_cols, a->dim, n, all, r, a->d_spare));
    std::swap(a->d_cols, a->d_spare);
    std::swap(a->cols_words, a->spare_words);
    a->dim = 5;
  }
  a->log_n -= 1;
  a->round_done = false;
  return LM_OK;
}

int lm_air_final(lm_air* a, uint32_t* out) {
  if (!a || !out) return fail(LM_ERR_INVALID, "lm_air_final: null argument");
  if (a->log_n != 0) return fail(LM_ERR_INVALID, "lm_air_final: %u variables are still unbound", a->log_n);
  lm_ctx* c = a->ctx;
  CU(cudaSetDevice(c->device));
  const size_t bytes = (size_t)(a->n_cols + a->n_shift) * 5 * sizeof(uint32_t);
  if (a->exec) {
    lm::AirExecArgs A;
    int mode = 0;
    if (int rc = air_exec_args(a, &A, &mode, true)) return rc;
    CU(lm::air_exec_final(c->stream, mode, A, a->d_out));
    CU(cudaMemcpyAsync(out, a->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  } else {
    if (a->dim != 5) return fail(LM_ERR_INVALID, "lm_air_final: the table was never folded");
    CU(cudaMemcpyAsync(out, a->d_cols, bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

// ------------------------------------------------------------------------------------------ Logup / quotient GKR
int lm_gkr_free(lm_gkr* g) {
  if (!g) return LM_OK;
  if (g->ctx) {
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
  }
  // layer 0 is owned separately (uploaded by lm_gkr_new or handed over by lm_logup_finish); the rest is one arena
  if (!g->nums.empty()) g->ctx->pool.free(g->nums[0]);
  if (!g->dens.empty()) g->ctx->pool.free(g->dens[0]);
  if (g->arena) g->ctx->pool.free(g->arena);
  delete g;
  return LM_OK;
}

// takes ownership of d_n (2^n_vars words) and d_d (five planes of 2^n_vars words), both already filled on [0, active_len)
static int gkr_from_device(lm_ctx* c, uint32_t* d_n, uint32_t* d_d, uint64_t active_len, uint32_t n_vars, lm_gkr** out,
                           uint32_t top_vars = LM_GKR_TOP_VARS) {
  lm_gkr* g = new (std::nothrow) lm_gkr();
  if (!g) {
    c->pool.free(d_n), c->pool.free(d_d);
    return fail(LM_ERR_OOM, "lm_gkr: host allocation failed");
  }
  g->ctx = c;
  g->n_vars = n_vars;
  g->top_vars = top_vars;
  g->nums.push_back(d_n);
  g->dens.push_back(d_d);
  const uint64_t n = (uint64_t)1 << n_vars;
  // one allocation for everything but layer 0 (40 cudaMalloc calls of up to hundreds of MiB cost more than the up pass)
  auto align = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t w_words = n_vars >= 2 ? ((size_t)1 << (n_vars - 2)) * 20 : 20;
  const size_t w1_words = w_words / 2 > 40 ? w_words / 2 : 40;
  g->tr_cap = 10 * n_vars * n_vars + 40 * n_vars + 256;
  size_t total = 0;
  for (uint32_t l = 1; l <= n_vars - top_vars; l++) total += 2 * align((n >> l) * 5 * sizeof(uint32_t));
  total += align(w_words * sizeof(uint32_t)) + align(w1_words * sizeof(uint32_t)) + align(lm::gkr_eq_table_words(n_vars) * sizeof(uint32_t)) +
           align((size_t)lm::GKR_MAX_BLOCKS * 10 * sizeof(uint32_t)) + align(sizeof(lm::GkrDev)) + align(sizeof(lm::DevFs)) +
           align((size_t)g->tr_cap * sizeof(uint32_t));
  cudaError_t e = g->ctx->pool.alloc(&g->arena, total);
  uint8_t* cur = static_cast<uint8_t*>(g->arena);
  auto carve = [&](size_t bytes) {
    uint32_t* p = reinterpret_cast<uint32_t*>(cur);
    cur += align(bytes);
    return p;
  };
  if (e == cudaSuccess) e = lm::gkr_pad(c->stream, d_n, d_d, active_len, n);
  // up pass (mod.rs:52-62): halve until 2^top_vars fractions remain
  for (uint32_t l = 1; e == cudaSuccess && l <= n_vars - top_vars; l++) {
    const uint64_t m = n >> l;
    uint32_t* nn = carve(m * 5 * sizeof(uint32_t));
    uint32_t* dd = carve(m * 5 * sizeof(uint32_t));
    g->nums.push_back(nn);
    g->dens.push_back(dd);
    e = lm::gkr_layer_up(c->stream, g->nums[l - 1], l == 1 ? 1 : 5, g->dens[l - 1], m * 2, nn, dd);
  }
  if (e == cudaSuccess) {
    // working tables: the first fused fold of the largest layer produces 2^(n_vars - 2) rows of 20 words
    g->d_w[0] = carve(w_words * sizeof(uint32_t));
    g->d_w[1] = carve(w1_words * sizeof(uint32_t));
    g->d_eq_tab = carve(lm::gkr_eq_table_words(n_vars) * sizeof(uint32_t));
    g->d_partial = carve((size_t)lm::GKR_MAX_BLOCKS * 10 * sizeof(uint32_t));
    g->d_g = reinterpret_cast<lm::GkrDev*>(carve(sizeof(lm::GkrDev)));
    g->d_fs = reinterpret_cast<lm::DevFs*>(carve(sizeof(lm::DevFs)));
    g->d_tr = carve((size_t)g->tr_cap * sizeof(uint32_t));
    e = cudaMemsetAsync(g->d_g, 0, sizeof(lm::GkrDev), c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    lm_gkr_free(g);
    return cuda_fail(e, "lm_gkr_new");
  }
  *out = g;
  return LM_OK;
}

// host AoS denominators (active_len x 5) -> device coefficient planes of stride n
static cudaError_t upload_dens_planes(lm_ctx* c, const uint32_t* dens, uint64_t active_len, uint64_t n, uint32_t* d_planes) {
  if (active_len == 0) return cudaSuccess;
  uint32_t* tmp = nullptr;
  cudaError_t e = c->pool.alloc(&tmp, active_len * 5 * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(tmp, dens, active_len * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = lm::gkr_aos_to_planes(c->stream, tmp, active_len, n, d_planes);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  c->pool.free(tmp);
  return e;
}

static int gkr_vars_for(uint64_t active_len, uint32_t* n_vars_out, const char* who) {
  if (active_len < 2) return fail(LM_ERR_INVALID, "%s: need at least two fractions", who);
  uint32_t n_vars = 0;
  while (((uint64_t)1 << n_vars) < active_len) n_vars++;
  if (n_vars <= LM_GKR_TOP_VARS) return fail(LM_ERR_INVALID, "%s: needs more than 2^%u fractions", who, LM_GKR_TOP_VARS);
  if (n_vars > 31) return fail(LM_ERR_INVALID, "%s: too many fractions", who);
  *n_vars_out = n_vars;
  return LM_OK;
}

int lm_gkr_new(lm_ctx* c, const uint32_t* nums, const uint32_t* dens, uint64_t active_len, lm_gkr** out) {
  if (!c || !nums || !dens || !out) return fail(LM_ERR_INVALID, "lm_gkr_new: null argument");
  *out = nullptr;
  uint32_t n_vars = 0;
  if (int rc = gkr_vars_for(active_len, &n_vars, "lm_gkr_new")) return rc;
  CU(cudaSetDevice(c->device));
  const uint64_t n = (uint64_t)1 << n_vars;
  uint32_t *d_n = nullptr, *d_d = nullptr;
  cudaError_t e = c->pool.alloc(&d_n, n * sizeof(uint32_t));
  if (e == cudaSuccess) e = c->pool.alloc(&d_d, n * 5 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, nums, active_len * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = upload_dens_planes(c, dens, active_len, n, d_d);
  if (e != cudaSuccess) {
    c->pool.free(d_n), c->pool.free(d_d);
    return cuda_fail(e, "lm_gkr_new");
  }
  return gkr_from_device(c, d_n, d_d, active_len, n_vars, out);
}

int lm_gkr_new_dev(lm_ctx* c, const uint32_t* d_nums_in, const uint32_t* d_dens_in, uint64_t active_len, lm_gkr** out) {
  if (!c || !d_nums_in || !d_dens_in || !out) return fail(LM_ERR_INVALID, "lm_gkr_new_dev: null argument");
  *out = nullptr;
  uint32_t n_vars = 0;
  if (int rc = gkr_vars_for(active_len, &n_vars, "lm_gkr_new_dev")) return rc;
  CU(cudaSetDevice(c->device));
  const uint64_t n = (uint64_t)1 << n_vars;
  uint32_t *d_n = nullptr, *d_d = nullptr;
  cudaError_t e = c->pool.alloc(&d_n, n * sizeof(uint32_t));
  if (e == cudaSuccess) e = c->pool.alloc(&d_d, n * 5 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, d_nums_in, active_len * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream);
  if (e == cudaSuccess) e = lm::gkr_aos_to_planes(c->stream, d_dens_in, active_len, n, d_d);
  if (e != cudaSuccess) {
    c->pool.free(d_n), c->pool.free(d_d);
    return cuda_fail(e, "lm_gkr_new_dev");
  }
  return gkr_from_device(c, d_n, d_d, active_len, n_vars, out);
}

int lm_gkr_new_shard(lm_ctx* c, const uint32_t* nums, const uint32_t* dens, uint64_t active_len, uint32_t n_vars,
                     uint32_t top_vars, lm_gkr** out) {
  if (!c || !out || (active_len && (!nums || !dens))) return fail(LM_ERR_INVALID, "lm_gkr_new_shard: null argument");
  *out = nullptr;
  if (n_vars < 2 || n_vars > 31 || top_vars < 1 || top_vars >= n_vars)
    return fail(LM_ERR_INVALID, "lm_gkr_new_shard: n_vars %u / top_vars %u out of range", n_vars, top_vars);
  const uint64_t n = (uint64_t)1 << n_vars;
  if (active_len > n) return fail(LM_ERR_INVALID, "lm_gkr_new_shard: active_len exceeds 2^n_vars");
  CU(cudaSetDevice(c->device));
  uint32_t *d_n = nullptr, *d_d = nullptr;
  cudaError_t e = c->pool.alloc(&d_n, n * sizeof(uint32_t));
  if (e == cudaSuccess) e = c->pool.alloc(&d_d, n * 5 * sizeof(uint32_t));
  if (e == cudaSuccess && active_len)
    e = cudaMemcpyAsync(d_n, nums, active_len * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && active_len) e = upload_dens_planes(c, dens, active_len, n, d_d);
  if (e != cudaSuccess) {
    c->pool.free(d_n), c->pool.free(d_d);
    return cuda_fail(e, "lm_gkr_new_shard");
  }
  return gkr_from_device(c, d_n, d_d, active_len, n_vars, out, top_vars);
}

// evaluation of a device-resident base/extension polynomial at a host point, result to the host
static int mle_eval_on_device(lm_ctx* c, const uint32_t* d_evals, uint32_t n_vars, uint32_t dim, uint64_t live_len,
                              const uint32_t* point, uint32_t out[5]) {
  if (n_vars > 64) return fail(LM_ERR_INVALID, "mle_eval: n_vars too large");
  if (n_vars) CU(cudaMemcpyAsync(c->d_point, point, (size_t)n_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  int rc = lm_dev_mle_eval(c, d_evals, n_vars, dim, live_len, c->d_point, c->d_small);
  if (rc != LM_OK) return rc;
  CU(cudaMemcpyAsync(out, c->d_small, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

// ---- Logup table builder (prove_generic_logup's table assembly, logup.rs:52-211) ---------------------------------
struct lm_logup {
  lm_ctx* ctx = nullptr;
  uint64_t total = 0, offset = 0;
  uint32_t n_vars = 0;
  uint32_t *d_nums = nullptr, *d_dens = nullptr;
  lm::Ef c;
  std::vector<uint32_t> alphas;  // n x 5
  std::map<std::pair<const void*, uint64_t>, uint32_t*> cache;  // host array -> device copy
  // the ~150 column copies of a proof are carved from a few large chunks: one cudaMalloc per column cost more than its copy
  std::vector<void*> chunks;
  uint8_t* arena_cur = nullptr;
  size_t arena_left = 0;
  uint32_t* d_batch_out = nullptr;  // results of lm_logup_col_eval_batch (256 x 5 words)

  int arena_get(size_t bytes, uint32_t** out) {
    bytes = (bytes + 255) / 256 * 256;
    if (bytes > arena_left) {
      const size_t chunk = bytes > ((size_t)256 << 20) ? bytes : ((size_t)256 << 20);
      void* p = nullptr;
      CU(ctx->pool.alloc(&p, chunk));
      chunks.push_back(p);
      arena_cur = static_cast<uint8_t*>(p);
      arena_left = chunk;
    }
    *out = reinterpret_cast<uint32_t*>(arena_cur);
    arena_cur += bytes;
    arena_left -= bytes;
    return LM_OK;
  }

  int device_copy(const uint32_t* host, uint64_t len, uint32_t** out) {
    auto key = std::make_pair((const void*)host, len);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return LM_OK;
    }
    uint32_t* d = nullptr;
    const uint64_t padded = (len + 1023) / 1024 * 1024 + 1024;  // lm_dev_mle_eval reads whole 2^10-element chunks
    if (int rc = arena_get(padded * sizeof(uint32_t), &d)) return rc;
    cache[key] = d;
    CU(cudaMemcpyAsync(d, host, len * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d + len, 0, (padded - len) * sizeof(uint32_t), ctx->stream));
    *out = d;
    return LM_OK;
  }
};

static uint32_t to_monty_u32(uint64_t canonical) { return (uint32_t)(((canonical % lm::KB_P) << 32) % lm::KB_P); }

int lm_logup_new(lm_ctx* c, uint64_t total_active_len, const uint32_t cc[5], const uint32_t* alphas_eq_poly, uint32_t n_alphas,
                 lm_logup** out) {
  if (!c || !cc || !alphas_eq_poly || !out) return fail(LM_ERR_INVALID, "lm_logup_new: null argument");
  *out = nullptr;
  if (n_alphas < 2) return fail(LM_ERR_INVALID, "lm_logup_new: need at least 2 alphas");
  uint32_t n_vars = 0;
  if (int rc = gkr_vars_for(total_active_len, &n_vars, "lm_logup_new")) return rc;
  CU(cudaSetDevice(c->device));
  lm_logup* L = new (std::nothrow) lm_logup();
  if (!L) return fail(LM_ERR_OOM, "lm_logup_new: host allocation failed");
  L->ctx = c;
  L->total = total_active_len;
  L->n_vars = n_vars;
  for (int k = 0; k < 5; k++) L->c.c[k] = cc[k];
  L->alphas.assign(alphas_eq_poly, alphas_eq_poly + 5 * (size_t)n_alphas);
  const uint64_t n = (uint64_t)1 << n_vars;
  cudaError_t e = c->pool.alloc(&L->d_nums, n * sizeof(uint32_t));
  if (e == cudaSuccess) e = c->pool.alloc(&L->d_dens, n * 5 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    lm_logup_free(L);
    return cuda_fail(e, "lm_logup_new");
  }
  *out = L;
  return LM_OK;
}

int lm_logup_section(lm_logup* L, uint64_t n_rows, uint32_t num_mode, const uint32_t* num_col, int32_t den_sign,
                     uint32_t domainsep, const lm_logup_data* data, uint32_t n_data) {
  if (!L || (n_data && !data)) return fail(LM_ERR_INVALID, "lm_logup_section: null argument");
  if (L->offset + n_rows > L->total)
    return fail(LM_ERR_INVALID, "lm_logup_section: %llu + %llu rows exceed total_active_len %llu", (unsigned long long)L->offset,
                (unsigned long long)n_rows, (unsigned long long)L->total);
  if (num_mode > 3) return fail(LM_ERR_INVALID, "lm_logup_section: numerator mode %u", num_mode);
  if ((num_mode == 1 || num_mode == 2) && !num_col) return fail(LM_ERR_INVALID, "lm_logup_section: numerator column is null");
  const size_t n_al = L->alphas.size() / 5;
  if (n_data >= n_al || n_data > (uint32_t)lm::LOGUP_MAX_DATA)
    return fail(LM_ERR_INVALID, "lm_logup_section: %u data columns need more than %zu alphas (max %d)", n_data, n_al, lm::LOGUP_MAX_DATA);
  lm_ctx* c = L->ctx;
  CU(cudaSetDevice(c->device));
  lm::LogupSection S{};
  S.n_rows = n_rows;
  S.num_mode = (int)num_mode;
  S.den_sign = den_sign > 0 ? 1 : (den_sign < 0 ? -1 : 0);
  S.n_data = (int)n_data;
  S.c = L->c;
  if (num_col) {
    uint32_t* d = nullptr;
    if (int rc = L->device_copy(num_col, n_rows, &d)) return rc;
    S.num_col = d;
  }
  // contrib = alphas.last * domainsep  (logup.rs:57-60)
  const uint32_t ds = to_monty_u32(domainsep);
  for (int k = 0; k < 5; k++) S.contrib.c[k] = lm::kb_mul(L->alphas[5 * (n_al - 1) + k], ds);
  for (uint32_t i = 0; i < n_data; i++) {
    for (int k = 0; k < 5; k++) S.alphas[i].c[k] = L->alphas[5 * i + k];
    lm::LogupData& d = S.data[i];
    d.kind = data[i].kind;
    d.add = to_monty_u32(data[i].value);
    d.offset = data[i].offset;
    d.stride = data[i].stride ? data[i].stride : 1;
    d.col = nullptr;
    if (d.kind == lm::LOGUP_DATA_COL) {
      if (!data[i].col) return fail(LM_ERR_INVALID, "lm_logup_section: data column %u is null", i);
      if (n_rows && d.offset + (n_rows - 1) * d.stride >= data[i].len)
        return fail(LM_ERR_INVALID, "lm_logup_section: data column %u is too short", i);
      uint32_t* dd = nullptr;
      if (int rc = L->device_copy(data[i].col, data[i].len, &dd)) return rc;
      d.col = dd;
    } else if (d.kind > 2) {
      return fail(LM_ERR_INVALID, "lm_logup_section: data kind %u", d.kind);
    }
  }
  CU(lm::logup_fill_section(c->stream, S, L->d_nums + L->offset, L->d_dens + L->offset, (uint64_t)1 << L->n_vars));
  L->offset += n_rows;
  return LM_OK;
}

int lm_logup_col_eval(lm_logup* L, const uint32_t* col, uint64_t len, uint32_t n_vars, const uint32_t* point, uint32_t out[5]) {
  if (!L || !col || !point || !out) return fail(LM_ERR_INVALID, "lm_logup_col_eval: null argument");
  if (n_vars > 40 || len > ((uint64_t)1 << n_vars)) return fail(LM_ERR_INVALID, "lm_logup_col_eval: column longer than 2^n_vars");
  lm_ctx* c = L->ctx;
  CU(cudaSetDevice(c->device));
  uint32_t* d = nullptr;
  if (int rc = L->device_copy(col, len, &d)) return rc;
  return mle_eval_on_device(c, d, n_vars, 1, len, point, out);
}

int lm_logup_col_eval_batch(lm_logup* L, const uint32_t* const* cols, const uint64_t* lens, uint32_t n_cols, uint32_t n_vars,
                            const uint32_t* point, uint32_t* out) {
  if (!L || (n_cols && (!cols || !lens || !out)) || (n_vars && !point)) return fail(LM_ERR_INVALID, "lm_logup_col_eval_batch: null argument");
  if (n_cols == 0) return LM_OK;
  if (n_cols > 256) return fail(LM_ERR_INVALID, "lm_logup_col_eval_batch: at most 256 columns per call, got %u", n_cols);
  if (n_vars > 40) return fail(LM_ERR_INVALID, "lm_logup_col_eval_batch: n_vars %u too large", n_vars);
  lm_ctx* c = L->ctx;
  CU(cudaSetDevice(c->device));
  std::vector<const uint32_t*> d_cols(n_cols);
  for (uint32_t k = 0; k < n_cols; k++) {
    if (!cols[k]) return fail(LM_ERR_INVALID, "lm_logup_col_eval_batch: column %u is null", k);
    if (lens[k] > ((uint64_t)1 << n_vars)) return fail(LM_ERR_INVALID, "lm_logup_col_eval_batch: column %u longer than 2^n_vars", k);
    uint32_t* d = nullptr;
    if (int rc = L->device_copy(cols[k], lens[k], &d)) return rc;
    d_cols[k] = d;
  }
  if (!L->d_batch_out) CU(c->pool.alloc(&L->d_batch_out, 256 * 5 * sizeof(uint32_t)));
  if (int rc = c->ensure_scratch(lm::mle_eval_scratch_words(n_vars))) return rc;
  if (n_vars) CU(cudaMemcpyAsync(c->d_point, point, (size_t)n_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(lm::mle_eval_batch(c->stream, d_cols.data(), lens, n_cols, n_vars, c->d_point, c->d_scratch, L->d_batch_out));
  CU(cudaMemcpyAsync(out, L->d_batch_out, (size_t)n_cols * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LM_OK;
}

int lm_logup_read(lm_logup* L, uint32_t* out_nums, uint32_t* out_dens) {
  if (!L) return fail(LM_ERR_INVALID, "lm_logup_read: null argument");
  lm_ctx* c = L->ctx;
  CU(cudaSetDevice(c->device));
  if (out_nums) CU(cudaMemcpyAsync(out_nums, L->d_nums, L->offset * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (out_dens && L->offset) {  // the device holds coefficient planes; the caller gets [F; 5] per row
    std::vector<uint32_t> plane(L->offset);
    for (int k = 0; k < 5; k++) {
      CU(cudaMemcpy(plane.data(), L->d_dens + ((size_t)k << L->n_vars), L->offset * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      for (uint64_t i = 0; i < L->offset; i++) out_dens[5 * i + k] = plane[i];
    }
  }
  return LM_OK;
}

int lm_logup_finish(lm_logup* L, lm_gkr** out) {
  if (!L || !out) return fail(LM_ERR_INVALID, "lm_logup_finish: null argument");
  *out = nullptr;
  if (L->offset != L->total)
    return fail(LM_ERR_INVALID, "lm_logup_finish: %llu of %llu rows filled", (unsigned long long)L->offset, (unsigned long long)L->total);
  if (!L->d_nums) return fail(LM_ERR_INVALID, "lm_logup_finish: already finished");
  CU(cudaSetDevice(L->ctx->device));
  uint32_t *d_n = L->d_nums, *d_d = L->d_dens;
  L->d_nums = L->d_dens = nullptr;  // ownership moves to the GKR session
  return gkr_from_device(L->ctx, d_n, d_d, L->total, L->n_vars, out);
}

int lm_logup_free(lm_logup* L) {
  if (!L) return LM_OK;
  if (L->ctx) cudaSetDevice(L->ctx->device);
  if (L->d_nums) L->ctx->pool.free(L->d_nums);
  if (L->d_dens) L->ctx->pool.free(L->d_dens);
  if (L->ctx) cudaStreamSynchronize(L->ctx->stream);
  for (void* p : L->chunks) L->ctx->pool.free(p);
  if (L->d_batch_out) L->ctx->pool.free(L->d_batch_out);
  delete L;
  return LM_OK;
}

int lm_gkr_num_vars(const lm_gkr* g, uint32_t* n_vars) {
  if (!g || !n_vars) return fail(LM_ERR_INVALID, "lm_gkr_num_vars: null argument");
  *n_vars = g->n_vars;
  return LM_OK;
}

int lm_gkr_top_vars(const lm_gkr* g, uint32_t* top_vars) {
  if (!g || !top_vars) return fail(LM_ERR_INVALID, "lm_gkr_top_vars: null argument");
  *top_vars = g->top_vars;
  return LM_OK;
}

int lm_gkr_top(lm_gkr* g, uint32_t* top_nums, uint32_t* top_dens) {
  if (!g || !top_nums || !top_dens) return fail(LM_ERR_INVALID, "lm_gkr_top: null argument");
  lm_ctx* c = g->ctx;
  CU(cudaSetDevice(c->device));
  const size_t m = (size_t)1 << g->top_vars;
  const bool base_nums = g->nums.size() == 1;  // no up pass at all (shards): numerators are still base field
  std::vector<uint32_t> pn(base_nums ? m : 5 * m), pd(5 * m);
  CU(cudaMemcpyAsync(pn.data(), g->nums.back(), pn.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(pd.data(), g->dens.back(), pd.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < m; i++)
    for (int k = 0; k < 5; k++) {
      top_nums[5 * i + k] = base_nums ? (k == 0 ? pn[i] : 0u) : pn[k * m + i];
      top_dens[5 * i + k] = pd[k * m + i];
    }
  return LM_OK;
}

static int gkr_layer_begin_impl(lm_gkr* g, uint32_t claim_vars, const uint32_t* point, const uint32_t alpha[5],
                                const uint32_t* eq_scale) {
  if (!g || !point || !alpha) return fail(LM_ERR_INVALID, "lm_gkr_layer_begin: null argument");
  if (claim_vars < g->top_vars || claim_vars >= g->n_vars || claim_vars > (uint32_t)lm::GKR_MAX_VARS)
    return fail(LM_ERR_INVALID, "lm_gkr_layer_begin: claim over %u variables, expected %u..%u", claim_vars, g->top_vars,
                g->n_vars - 1);
  lm_ctx* c = g->ctx;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(g->d_g) + offsetof(lm::GkrDev, point), point,
                     (size_t)claim_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(g->d_g) + offsetof(lm::GkrDev, alpha), alpha, 5 * sizeof(uint32_t),
                     cudaMemcpyHostToDevice, c->stream));
  g->cur_layer = (int)(g->n_vars - (claim_vars + 1));  // the layer below the claim has claim_vars + 1 variables
  g->cur_k = claim_vars;                               // its even/odd halves have claim_vars variables
  g->cur_rnd = 0;
  g->round_done = false;
  CU(lm::gkr_begin(c->stream, g->layer_args(g->cur_layer, g->cur_k, false), eq_scale, false));
  return LM_OK;
}

int lm_gkr_layer_begin(lm_gkr* g, uint32_t claim_vars, const uint32_t* point, const uint32_t alpha[5]) {
  return gkr_layer_begin_impl(g, claim_vars, point, alpha, nullptr);
}

int lm_gkr_layer_begin_shard(lm_gkr* g, uint32_t claim_vars, const uint32_t* point, const uint32_t alpha[5],
                             const uint32_t eq_scale[5]) {
  if (!eq_scale) return fail(LM_ERR_INVALID, "lm_gkr_layer_begin_shard: null argument");
  return gkr_layer_begin_impl(g, claim_vars, point, alpha, eq_scale);
}

// The fold of lm_gkr_fold is deferred: the round that follows folds the previous table with the stored challenge while it
// reads it (one pass over the table per round), the last one is applied by lm_gkr_layer_end.
int lm_gkr_round(lm_gkr* g, uint32_t c0[5], uint32_t c2[5]) {
  if (!g || !c0 || !c2) return fail(LM_ERR_INVALID, "lm_gkr_round: null argument");
  if (g->cur_layer < 0 || g->cur_rnd >= g->cur_k) return fail(LM_ERR_INVALID, "lm_gkr_round: no layer sumcheck in progress");
  if (g->round_done) return fail(LM_ERR_INVALID, "lm_gkr_round: the previous round has not been folded");
  lm_ctx* c = g->ctx;
  CU(cudaSetDevice(c->device));
  CU(lm::gkr_round(c->stream, g->layer_args(g->cur_layer, g->cur_k, false), g->cur_rnd));
  uint32_t h[10];
  CU(cudaMemcpyAsync(h, reinterpret_cast<uint8_t*>(g->d_g) + offsetof(lm::GkrDev, out), sizeof(h), cudaMemcpyDeviceToHost,
                     c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(c0, h, 5 * sizeof(uint32_t));
  memcpy(c2, h + 5, 5 * sizeof(uint32_t));
  g->round_done = true;
  return LM_OK;
}

int lm_gkr_fold(lm_gkr* g, const uint32_t r[5]) {
  if (!g || !r) return fail(LM_ERR_INVALID, "lm_gkr_fold: null argument");
  if (g->cur_layer < 0 || g->cur_rnd >= g->cur_k) return fail(LM_ERR_INVALID, "lm_gkr_fold: no layer sumcheck in progress");
  lm_ctx* c = g->ctx;
  CU(cudaSetDevice(c->device));
  if (!g->round_done) return fail(LM_ERR_INVALID, "lm_gkr_fold: call lm_gkr_round first (the fold is fused into the next round)");
  CU(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(g->d_g) + offsetof(lm::GkrDev, r), r, 5 * sizeof(uint32_t), cudaMemcpyHostToDevice,
                     c->stream));
  g->cur_rnd += 1;
  g->round_done = false;
  return LM_OK;
}

int lm_gkr_layer_end(lm_gkr* g, uint32_t* inner_evals) {
  if (!g || !inner_evals) return fail(LM_ERR_INVALID, "lm_gkr_layer_end: null argument");
  if (g->cur_layer < 0 || g->cur_rnd != g->cur_k)
    return fail(LM_ERR_INVALID, "lm_gkr_layer_end: %u variables are still unbound", g->cur_k - g->cur_rnd);
  lm_ctx* c = g->ctx;
  CU(cudaSetDevice(c->device));
  CU(lm::gkr_end(c->stream, g->layer_args(g->cur_layer, g->cur_k, false)));
  CU(cudaMemcpyAsync(inner_evals, reinterpret_cast<uint8_t*>(g->d_g) + offsetof(lm::GkrDev, inner), 20 * sizeof(uint32_t),
                     cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  g->cur_layer = -1;
  return LM_OK;
}

}  // extern "C"

// Every layer sumcheck of prove_gkr_quotient with the challenger on the device (spine.cu's lm_gkr_prove is the caller): the
// sponge state goes in, all launches are enqueued back to back, and the state, the transcript words the device appended,
// the final point and the two claims come back with ONE synchronisation.
int lm_internal_gkr_device_layers(lm_gkr* g, uint32_t state[16], int* rate_fresh, const uint32_t* point, const uint32_t claim_num[5],
                                  const uint32_t claim_den[5], std::vector<uint32_t>* transcript, uint32_t* out_point,
                                  uint32_t out_claim_num[5], uint32_t out_claim_den[5]) {
  lm_ctx* c = g->ctx;
  CU(cudaSetDevice(c->device));
  if (g->n_vars > (uint32_t)lm::GKR_MAX_VARS) return fail(LM_ERR_INVALID, "lm_gkr_prove: too many variables");
  std::vector<uint8_t> host(sizeof(lm::GkrDev), 0);
  lm::GkrDev* hg = reinterpret_cast<lm::GkrDev*>(host.data());
  memcpy(hg->point, point, (size_t)g->top_vars * 5 * sizeof(uint32_t));
  memcpy(hg->claim_num.c, claim_num, 5 * sizeof(uint32_t));
  memcpy(hg->claim_den.c, claim_den, 5 * sizeof(uint32_t));
  hg->k = g->top_vars;
  lm::DevFs hf{};
  memcpy(hf.state, state, sizeof(hf.state));
  hf.rate_fresh = *rate_fresh ? 1u : 0u;
  hf.n_words = 0;
  hf.cap_words = g->tr_cap;
  hf.error = 0;
  CU(cudaMemcpyAsync(g->d_g, hg, sizeof(lm::GkrDev), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(g->d_fs, &hf, sizeof(hf), cudaMemcpyHostToDevice, c->stream));
  CU(lm::gkr_begin(c->stream, g->layer_args((int)(g->n_vars - (g->top_vars + 1)), g->top_vars, true), nullptr, true));
  for (uint32_t k = g->top_vars; k < g->n_vars; k++) {
    const lm::GkrLayerArgs a = g->layer_args((int)(g->n_vars - (k + 1)), k, true);
    if (k + 1 < g->n_vars) {
      const lm::GkrLayerArgs nx = g->layer_args((int)(g->n_vars - (k + 2)), k + 1, true);
      CU(lm::gkr_layer_device(c->stream, a, &nx));
    } else {
      CU(lm::gkr_layer_device(c->stream, a, nullptr));
    }
  }
  CU(cudaMemcpyAsync(hg, g->d_g, sizeof(lm::GkrDev), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&hf, g->d_fs, sizeof(hf), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (hf.error) return fail(LM_ERR_INVALID, "lm_gkr_prove: device challenger error 0x%x (1 stale rate, 2 transcript overflow, 4 zero eq coordinate)", hf.error);
  if (hf.n_words) {
    const size_t at = transcript->size();
    transcript->resize(at + hf.n_words);
    CU(cudaMemcpy(transcript->data() + at, g->d_tr, (size_t)hf.n_words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  memcpy(state, hf.state, sizeof(hf.state));
  *rate_fresh = hf.rate_fresh ? 1 : 0;
  memcpy(out_point, hg->point, (size_t)g->n_vars * 5 * sizeof(uint32_t));
  memcpy(out_claim_num, hg->claim_num.c, 5 * sizeof(uint32_t));
  memcpy(out_claim_den, hg->claim_den.c, 5 * sizeof(uint32_t));
  return LM_OK;
}

extern "C" {

int lm_finger_print(lm_ctx* c, const uint32_t* data, uint64_t n_rows, uint32_t n_data, const uint32_t* alphas,
                    const uint32_t cc[5], uint32_t* out) {
  if (!c || !data || !alphas || !cc || !out) return fail(LM_ERR_INVALID, "lm_finger_print: null argument");
  if (n_rows == 0) return LM_OK;
  CU(cudaSetDevice(c->device));
  uint32_t *d_data = nullptr, *d_al = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_data, n_rows * n_data * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_al, (size_t)n_data * 5 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, n_rows * 5 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_data, data, n_rows * n_data * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_al, alphas, (size_t)n_data * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = lm::finger_print(c->stream, d_data, n_rows, n_data, d_al, cc, d_out);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, n_rows * 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream);
  if (d_data) cudaFree(d_data);
  if (d_al) cudaFree(d_al);
  if (d_out) cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "lm_finger_print");
  return LM_OK;
}

int lm_mle_eval(lm_ctx* c, const uint32_t* evals, uint32_t n_vars, uint32_t dim, uint64_t live_len, const uint32_t* point,
                uint32_t out[5]) {
  if (!c || !evals || !point || !out) return fail(LM_ERR_INVALID, "lm_mle_eval: null argument");
  if (dim != 1 && dim != 5) return fail(LM_ERR_INVALID, "elem_dim must be 1 or 5, got %u", dim);
  if (n_vars > 34) return fail(LM_ERR_INVALID, "lm_mle_eval: n_vars too large");
  const uint64_t len = (uint64_t)1 << n_vars;
  if (live_len > len) return fail(LM_ERR_INVALID, "lm_mle_eval: live_len exceeds 2^n_vars");
  CU(cudaSetDevice(c->device));
  uint32_t* d_evals = nullptr;
  // rows are read in 2^10-element chunks: round the live prefix up and zero the tail
  const uint64_t chunk = n_vars < 10 ? len : 1024;
  uint64_t padded = (live_len + chunk - 1) / chunk * chunk;
  if (padded == 0) padded = chunk;
  CU(cudaMalloc(&d_evals, padded * dim * sizeof(uint32_t)));
  cudaError_t e = cudaMemcpyAsync(d_evals, evals, live_len * dim * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && padded > live_len)
    e = cudaMemsetAsync(d_evals + live_len * dim, 0, (padded - live_len) * dim * sizeof(uint32_t), c->stream);
  if (e == cudaSuccess && n_vars)
    e = cudaMemcpyAsync(c->d_point, point, (size_t)n_vars * 5 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
  int rc = LM_OK;
  if (e != cudaSuccess) rc = cuda_fail(e, "lm_mle_eval upload");
  if (rc == LM_OK) rc = lm_dev_mle_eval(c, d_evals, n_vars, dim, live_len, c->d_point, c->d_small);
  if (rc == LM_OK) {
    e = cudaMemcpyAsync(out, c->d_small, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "lm_mle_eval download");
  }
  cudaStreamSynchronize(c->stream);
  cudaFree(d_evals);
  return rc;
}

}  // extern "C"

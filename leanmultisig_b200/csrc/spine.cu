// Host spine in C++: the sequential, transcript-bound half of the prover that the reference runs in Rust around the
// data-parallel steps — built ABOVE the C ABI (it only calls the public lm_* entry points) so that the GKR and AIR
// sumcheck drivers do not pay an interpreter round trip per round.
//
//   lm_fs_*               ProverState / Challenger      crates/backend/fiat-shamir/src/prover.rs:28-178, challenger.rs:8-76
//   lm_gkr_prove          prove_gkr_quotient            crates/sub_protocols/src/quotient_gkr/mod.rs:31-141
//                         build_bare_from_coeffs        quotient_gkr/sumcheck_utils.rs:491-503
//   lm_air_prove_batched  prove_batched_air_sumcheck    crates/sub_protocols/src/air_sumcheck.rs:636-681
//                         compute_bare_round_poly / process_challenge   air_sumcheck.rs:225-287
//                         expand_bare_to_full           crates/backend/fiat-shamir/src/utils.rs:30-41
//                         lagrange_interpolation        crates/backend/poly/src/dense_poly.rs:33
//
// In a Rust integration this file is not needed (the Rust ProverState and drivers stay); it is the C++ mirror the task
// asks for where the reference's host code is compiled code, and the Python classes of the same names delegate to it.
// Field elements are Montgomery-form u32 throughout (kb.cuh's host paths).
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>
#include "../../include/leanmultisig_b200.h"
#include "kb.cuh"
#include "merkle.h"

int lm_internal_fail(int code, const char* msg);  // capi.cu: sets lm_last_error
// capi.cu: all layer sumchecks of the quotient GKR with the challenger on the device (devfs.cuh)
int lm_internal_gkr_device_layers(lm_gkr* g, uint32_t state[16], int* rate_fresh, const uint32_t* point, const uint32_t claim_num[5],
                                  const uint32_t claim_den[5], std::vector<uint32_t>* transcript, uint32_t* out_point,
                                  uint32_t out_claim_num[5], uint32_t out_claim_den[5]);

namespace {
using lm::Ef;
using lm::KB_P;
using lm::KB_R1;

const Ef EF_ZERO = {{0, 0, 0, 0, 0}};
const Ef EF_ONE = {{KB_R1, 0, 0, 0, 0}};

uint32_t to_monty(uint64_t canonical) { return (uint32_t)(((canonical % KB_P) << 32) % KB_P); }
uint32_t from_monty(uint32_t m) { return lm::kb_canon(lm::kb_redc_lazy((uint64_t)m)); }

uint32_t kb_pow(uint32_t a, uint64_t e) {
  uint32_t r = KB_R1;
  for (; e; e >>= 1) {
    if (e & 1) r = lm::kb_mul(r, a);
    a = lm::kb_mul(a, a);
  }
  return r;
}
uint32_t kb_inv(uint32_t a) { return kb_pow(a, KB_P - 2); }

// X^(p i), i = 1..4 (quintic_extension/mod.rs:19-48), canonical residues; the same table as leanmultisig_b200/field.py
const uint32_t FROBENIUS_CANON[4][5] = {
    {1576402667, 1173144480, 1567662457, 1206866823, 2428146},
    {1680345488, 1381986, 615237464, 1380104858, 295431824},
    {441230756, 323126830, 704986542, 1445620072, 503505220},
    {1364444097, 1144738982, 2008416047, 143367062, 1027410849},
};
Ef ef_frobenius(const Ef& a) {
  static uint32_t fm[4][5];
  static bool ready = false;
  if (!ready) {
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 5; k++) fm[i][k] = to_monty(FROBENIUS_CANON[i][k]);
    ready = true;
  }
  Ef out = {{a.c[0], 0, 0, 0, 0}};
  for (int i = 1; i < 5; i++)
    for (int k = 0; k < 5; k++) out.c[k] = lm::kb_add(out.c[k], lm::kb_mul(a.c[i], fm[i - 1][k]));
  return out;
}
// a^-1 = (a^p a^(p^2) a^(p^3) a^(p^4)) / Norm(a)   (quintic_extension/extension.rs:585-613)
bool ef_inv(const Ef& a, Ef* out) {
  const Ef f1 = ef_frobenius(a);
  const Ef f12 = ef_frobenius(lm::ef_mul(a, f1));
  const Ef conj = lm::ef_mul(f12, ef_frobenius(ef_frobenius(f12)));
  const Ef norm = lm::ef_mul(a, conj);
  if (norm.c[0] == 0) return false;
  *out = lm::ef_mul_base(conj, kb_inv(norm.c[0]));
  return true;
}
Ef ef_load(const uint32_t* p) {
  Ef e;
  memcpy(e.c, p, sizeof(e.c));
  return e;
}
void ef_store(uint32_t* p, const Ef& e) { memcpy(p, e.c, sizeof(e.c)); }
Ef poly_eval(const std::vector<Ef>& coeffs, const Ef& x) {
  Ef acc = EF_ZERO;
  for (size_t i = coeffs.size(); i-- > 0;) acc = lm::ef_add(lm::ef_mul(acc, x), coeffs[i]);
  return acc;
}
// eq(alpha, r) = (1 - alpha)(1 - r) + alpha r
Ef eq1(const Ef& alpha, const Ef& r) {
  return lm::ef_add(lm::ef_mul(lm::ef_sub(EF_ONE, alpha), lm::ef_sub(EF_ONE, r)), lm::ef_mul(alpha, r));
}
// coefficients of the polynomial of degree < n with p(i) = values[i], i = 0..n-1
std::vector<Ef> lagrange_at_integers(const std::vector<Ef>& values) {
  const int n = (int)values.size();
  std::vector<Ef> coeffs(n, EF_ZERO);
  for (int i = 0; i < n; i++) {
    std::vector<uint32_t> num(1, KB_R1);  // prod_{j != i} (X - j), Montgomery base-field coefficients
    uint32_t denom = KB_R1;
    for (int j = 0; j < n; j++) {
      if (j == i) continue;
      const uint32_t jm = to_monty((uint64_t)j);
      std::vector<uint32_t> next(num.size() + 1, 0);
      for (size_t k = 0; k < num.size(); k++) {
        next[k + 1] = lm::kb_add(next[k + 1], num[k]);
        next[k] = lm::kb_sub(next[k], lm::kb_mul(jm, num[k]));
      }
      num.swap(next);
      denom = lm::kb_mul(denom, i >= j ? to_monty((uint64_t)(i - j)) : lm::kb_neg(to_monty((uint64_t)(j - i))));
    }
    const uint32_t dinv = kb_inv(denom);
    for (int k = 0; k < n; k++) coeffs[k] = lm::ef_add(coeffs[k], lm::ef_mul_base(values[i], lm::kb_mul(num[k], dinv)));
  }
  return coeffs;
}
// full(X) = ((1 - alpha) + (2 alpha - 1) X) * bare(X)
std::vector<Ef> expand_bare_to_full(const std::vector<Ef>& bare, const Ef& alpha) {
  const Ef c0 = lm::ef_sub(EF_ONE, alpha), c1 = lm::ef_sub(lm::ef_add(alpha, alpha), EF_ONE);
  std::vector<Ef> full(bare.size() + 1, EF_ZERO);
  for (size_t i = 0; i < bare.size(); i++) {
    full[i] = lm::ef_add(full[i], lm::ef_mul(c0, bare[i]));
    full[i + 1] = lm::ef_add(full[i + 1], lm::ef_mul(c1, bare[i]));
  }
  return full;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ ProverState
struct lm_fs {
  lm_ctx* ctx = nullptr;
  uint32_t state[16] = {0};
  bool rate_fresh = false;
  std::vector<uint32_t> transcript;

  void observe(const uint32_t chunk[8]) {
    memcpy(state + 8, chunk, 8 * sizeof(uint32_t));
    lm::poseidon1_permute_host(state);
    rate_fresh = true;
  }
  void observe_many(const uint32_t* s, size_t n) {
    for (size_t i = 0; i < n; i += 8) {
      uint32_t chunk[8] = {0};
      memcpy(chunk, s + i, (n - i < 8 ? n - i : 8) * sizeof(uint32_t));
      observe(chunk);
    }
  }
  void duplex() {
    const uint32_t zeros[8] = {0};
    observe(zeros);
  }
  bool sample_many(size_t n_blocks, std::vector<uint32_t>* out) {
    for (size_t i = 0; i < n_blocks; i++) {
      if (i) duplex();
      if (!rate_fresh) return false;
      rate_fresh = false;
      out->insert(out->end(), state + 8, state + 16);
    }
    return true;
  }
  bool sample_ef(size_t n, std::vector<Ef>* out) {
    std::vector<uint32_t> fes;
    if (!sample_many((n * 5 + 7) / 8, &fes)) return false;
    for (size_t i = 0; i < n; i++) out->push_back(ef_load(fes.data() + 5 * i));
    return true;
  }
  void add_scalars(const uint32_t* s, size_t n) {
    observe_many(s, n);
    transcript.insert(transcript.end(), s, s + n);
  }
  void add_ef(const std::vector<Ef>& v) {
    std::vector<uint32_t> flat;
    for (const Ef& e : v) flat.insert(flat.end(), e.c, e.c + 5);
    add_scalars(flat.data(), flat.size());
  }
  // prover.rs:105-128: everything is absorbed (the full polynomial when eq_alpha is given), the constant coefficient is not sent
  void add_sumcheck_polynomial(const std::vector<Ef>& coeffs, const Ef* eq_alpha) {
    const std::vector<Ef> absorbed = eq_alpha ? expand_bare_to_full(coeffs, *eq_alpha) : coeffs;
    std::vector<uint32_t> flat;
    for (const Ef& e : absorbed) flat.insert(flat.end(), e.c, e.c + 5);
    observe_many(flat.data(), flat.size());
    for (size_t i = 1; i < coeffs.size(); i++) transcript.insert(transcript.end(), coeffs[i].c, coeffs[i].c + 5);
  }
};

extern "C" {

int lm_fs_new(lm_ctx* ctx, lm_fs** out) {
  if (!out) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_new: null argument");
  lm_fs* fs = new (std::nothrow) lm_fs();
  if (!fs) return lm_internal_fail(LM_ERR_OOM, "lm_fs_new: host allocation failed");
  fs->ctx = ctx;
  *out = fs;
  return LM_OK;
}
int lm_fs_free(lm_fs* fs) {
  delete fs;
  return LM_OK;
}
int lm_fs_add_scalars(lm_fs* fs, const uint32_t* words, uint64_t n) {
  if (!fs || (n && !words)) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_add_scalars: null argument");
  fs->add_scalars(words, n);
  return LM_OK;
}
int lm_fs_observe(lm_fs* fs, const uint32_t* words, uint64_t n) {
  if (!fs || (n && !words)) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_observe: null argument");
  fs->observe_many(words, n);
  return LM_OK;
}
int lm_fs_duplex(lm_fs* fs) {
  if (!fs) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_duplex: null argument");
  fs->duplex();
  return LM_OK;
}
int lm_fs_add_sumcheck_polynomial(lm_fs* fs, const uint32_t* coeffs, uint32_t n_coeffs, const uint32_t* eq_alpha) {
  if (!fs || !coeffs || n_coeffs == 0) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_add_sumcheck_polynomial: null argument");
  std::vector<Ef> c;
  for (uint32_t i = 0; i < n_coeffs; i++) c.push_back(ef_load(coeffs + 5 * i));
  if (eq_alpha) {
    const Ef a = ef_load(eq_alpha);
    fs->add_sumcheck_polynomial(c, &a);
  } else {
    fs->add_sumcheck_polynomial(c, nullptr);
  }
  return LM_OK;
}
int lm_fs_sample(lm_fs* fs, uint32_t n, uint32_t* out) {
  if (!fs || (n && !out)) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_sample: null argument");
  std::vector<Ef> v;
  if (!fs->sample_ef(n, &v)) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_sample: stale rate, insert a duplex() before");
  for (uint32_t i = 0; i < n; i++) ef_store(out + 5 * i, v[i]);
  return LM_OK;
}
int lm_fs_sample_in_range(lm_fs* fs, uint32_t bits, uint32_t n, uint64_t* out) {
  if (!fs || (n && !out) || bits > 31) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_sample_in_range: bad argument");
  std::vector<uint32_t> fes;
  if (!fs->sample_many((n + 7) / 8, &fes)) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_sample_in_range: stale rate");
  for (uint32_t i = 0; i < n; i++) out[i] = from_monty(fes[i]) & (((uint64_t)1 << bits) - 1);
  return LM_OK;
}
int lm_fs_pow_grinding(lm_fs* fs, uint32_t bits) {
  if (!fs) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_pow_grinding: null argument");
  if (bits == 0) return LM_OK;
  if (!fs->ctx) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_pow_grinding: the transcript was created without a device context");
  uint64_t w = 0;
  if (int rc = lm_pow_grind(fs->ctx, fs->state, bits, &w)) return rc;
  const uint32_t wm = to_monty(w);
  fs->observe_many(&wm, 1);
  if (from_monty(fs->state[8]) & (((uint64_t)1 << bits) - 1))
    return lm_internal_fail(LM_ERR_CUDA, "lm_fs_pow_grinding: device PoW witness rejected by the host sponge");
  fs->transcript.push_back(wm);
  return LM_OK;
}
int lm_fs_transcript_len(const lm_fs* fs, uint64_t* n) {
  if (!fs || !n) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_transcript_len: null argument");
  *n = fs->transcript.size();
  return LM_OK;
}
int lm_fs_transcript(const lm_fs* fs, uint32_t* out) {
  if (!fs || (!out && !fs->transcript.empty())) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_transcript: null argument");
  if (!fs->transcript.empty()) memcpy(out, fs->transcript.data(), fs->transcript.size() * sizeof(uint32_t));
  return LM_OK;
}
int lm_fs_set_state(lm_fs* fs, const uint32_t state[16], int rate_fresh) {
  if (!fs || !state) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_set_state: null argument");
  memcpy(fs->state, state, sizeof(fs->state));
  fs->rate_fresh = rate_fresh != 0;
  return LM_OK;
}
int lm_fs_state(const lm_fs* fs, uint32_t state[16], int* rate_fresh) {
  if (!fs || !state) return lm_internal_fail(LM_ERR_INVALID, "lm_fs_state: null argument");
  memcpy(state, fs->state, sizeof(fs->state));
  if (rate_fresh) *rate_fresh = fs->rate_fresh ? 1 : 0;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ WHIR round bookkeeping
int lm_whir_stir_update(lm_sumcheck* sc, const uint64_t* idx, uint32_t n_q, uint32_t gen, uint32_t num_variables,
                        const uint32_t comb[5], const uint32_t* ood_ys, const uint32_t* ood_answers, uint32_t n_ood,
                        const uint32_t* stir_evals, uint32_t total_io[5]) {
  if (!sc || !comb || !total_io || (n_q && (!idx || !stir_evals)) || (n_ood && (!ood_ys || !ood_answers)))
    return lm_internal_fail(LM_ERR_INVALID, "lm_whir_stir_update: null argument");
  const Ef c = ef_load(comb);
  Ef total = ef_load(total_io), pw = EF_ONE;
  std::vector<uint32_t> pt(5 * (size_t)num_variables);
  for (uint32_t k = 0; k < n_ood; k++) {
    // MultilinearPoint::expand_from_univariate (poly/src/point.rs:51-61): y, y^2, y^4, ...
    Ef cur = ef_load(ood_ys + 5 * k);
    for (uint32_t v = 0; v < num_variables; v++) {
      ef_store(pt.data() + 5 * v, cur);
      cur = lm::ef_mul(cur, cur);
    }
    if (int rc = lm_sc_add_eq(sc, 0, pt.data(), num_variables, pw.c)) return rc;
    total = lm::ef_add(total, lm::ef_mul(pw, ef_load(ood_answers + 5 * k)));
    pw = lm::ef_mul(pw, c);
  }
  if (n_q) {
    // in-domain points gen^idx expanded to (x, x^2, x^4, ...) in the base field, one scalar comb^(n_ood + q) each
    std::vector<uint32_t> pts((size_t)n_q * num_variables), scal(5 * (size_t)n_q);
    for (uint32_t q = 0; q < n_q; q++) {
      uint32_t y = kb_pow(gen, idx[q]);
      for (uint32_t v = 0; v < num_variables; v++) {
        pts[(size_t)q * num_variables + v] = y;
        y = lm::kb_mul(y, y);
      }
      ef_store(scal.data() + 5 * q, pw);
      total = lm::ef_add(total, lm::ef_mul(pw, ef_load(stir_evals + 5 * q)));
      pw = lm::ef_mul(pw, c);
    }
    if (int rc = lm_sc_add_base_eq(sc, pts.data(), n_q, scal.data())) return rc;
  }
  ef_store(total_io, total);
  return LM_OK;
}

// sumcheck_prove_many_rounds for the product sumcheck of the opening (sumcheck/src/prove.rs:86-151 over
// product_computation.rs:37-125), the whole phase without returning to the caller between rounds: per round the device pass
// (first round: lm_sc_round; later: the fold by the previous challenge fused with the round, lm_sc_fold_round), c1 from the
// running sum, absorb (c0, c1, c2), PoW, sample r, sum <- c0 + r (c1 + r c2); the last challenge is folded at the end.
int lm_whir_sumcheck_rounds(lm_sumcheck* sc, lm_fs* fs, uint32_t n_rounds, uint32_t pow_bits, uint32_t total_io[5],
                            uint32_t* out_challenges) {
  if (!sc || !fs || !total_io || (n_rounds && !out_challenges))
    return lm_internal_fail(LM_ERR_INVALID, "lm_whir_sumcheck_rounds: null argument");
  Ef total = ef_load(total_io), pending = EF_ZERO;
  for (uint32_t i = 0; i < n_rounds; i++) {
    uint32_t c0w[5], c2w[5];
    if (int rc = i ? lm_sc_fold_round(sc, pending.c, c0w, c2w) : lm_sc_round(sc, c0w, c2w)) return rc;
    const Ef c0 = ef_load(c0w), c2 = ef_load(c2w);
    const Ef c1 = lm::ef_sub(lm::ef_sub(total, lm::ef_add(c0, c0)), c2);  // h(0) + h(1) = sum, h(1) = c0 + c1 + c2
    fs->add_sumcheck_polynomial({c0, c1, c2}, nullptr);
    if (int rc = lm_fs_pow_grinding(fs, pow_bits)) return rc;
    std::vector<Ef> r;
    if (!fs->sample_ef(1, &r)) return lm_internal_fail(LM_ERR_INVALID, "lm_whir_sumcheck_rounds: stale rate");
    pending = r[0];
    total = lm::ef_add(c0, lm::ef_mul(pending, lm::ef_add(c1, lm::ef_mul(pending, c2))));
    ef_store(out_challenges + 5 * i, pending);
  }
  if (n_rounds)
    if (int rc = lm_sc_fold(sc, pending.c)) return rc;
  ef_store(total_io, total);
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ quotient GKR
int lm_gkr_prove(lm_gkr* gkr, lm_fs* fs, uint32_t out_quotient[5], uint32_t* out_point, uint32_t out_claim_num[5],
                 uint32_t out_claim_den[5]) {
  if (!gkr || !fs || !out_quotient || !out_point || !out_claim_num || !out_claim_den)
    return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: null argument");
  uint32_t n_vars = 0;
  if (int rc = lm_gkr_num_vars(gkr, &n_vars)) return rc;
  const uint32_t TOP = 5;  // N_VARS_TO_SEND_GKR_COEFFS
  uint32_t top_vars = 0;
  if (int rc = lm_gkr_top_vars(gkr, &top_vars)) return rc;
  if (top_vars != TOP) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: shard sessions are driven by the caller (sharded.py)");
  std::vector<uint32_t> tn(32 * 5), td(32 * 5);
  if (int rc = lm_gkr_top(gkr, tn.data(), td.data())) return rc;
  fs->add_scalars(tn.data(), tn.size());
  fs->add_scalars(td.data(), td.size());
  std::vector<Ef> top_n, top_d;
  Ef quotient = EF_ZERO;
  for (int i = 0; i < 32; i++) {
    top_n.push_back(ef_load(tn.data() + 5 * i));
    top_d.push_back(ef_load(td.data() + 5 * i));
    Ef inv;
    if (!ef_inv(top_d[i], &inv)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: a top-layer denominator is zero");
    quotient = lm::ef_add(quotient, lm::ef_mul(top_n[i], inv));
  }
  std::vector<Ef> point;
  if (!fs->sample_ef(TOP, &point)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: stale rate");
  auto mle_small = [&](std::vector<Ef> cur) {
    for (const Ef& x : point) {
      const size_t half = cur.size() / 2;
      for (size_t i = 0; i < half; i++) cur[i] = lm::ef_add(cur[i], lm::ef_mul(x, lm::ef_sub(cur[i + half], cur[i])));
      cur.resize(half);
    }
    return cur[0];
  };
  Ef claim_num = mle_small(top_n), claim_den = mle_small(top_d);

  // prove_gkr_layer for every layer (mod.rs:64-75) on the device: no host round trip per round
  std::vector<uint32_t> pt(5 * (size_t)TOP), out_pt(5 * (size_t)n_vars);
  for (uint32_t i = 0; i < TOP; i++) ef_store(pt.data() + 5 * i, point[i]);
  int fresh = fs->rate_fresh ? 1 : 0;
  if (int rc = lm_internal_gkr_device_layers(gkr, fs->state, &fresh, pt.data(), claim_num.c, claim_den.c, &fs->transcript,
                                             out_pt.data(), out_claim_num, out_claim_den))
    return rc;
  fs->rate_fresh = fresh != 0;
  ef_store(out_quotient, quotient);
  memcpy(out_point, out_pt.data(), out_pt.size() * sizeof(uint32_t));
  return LM_OK;
}

// The same driver with the round loop and the transcript on the host (one synchronisation per round): what a caller that
// keeps its own ProverState does through lm_gkr_layer_begin / round / fold / layer_end; kept to cross-check lm_gkr_prove.
int lm_gkr_prove_hostloop(lm_gkr* gkr, lm_fs* fs, uint32_t out_quotient[5], uint32_t* out_point, uint32_t out_claim_num[5],
                 uint32_t out_claim_den[5]) {
  if (!gkr || !fs || !out_quotient || !out_point || !out_claim_num || !out_claim_den)
    return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove_hostloop: null argument");
  uint32_t n_vars = 0;
  if (int rc = lm_gkr_num_vars(gkr, &n_vars)) return rc;
  const uint32_t TOP = 5;  // N_VARS_TO_SEND_GKR_COEFFS
  uint32_t top_vars = 0;
  if (int rc = lm_gkr_top_vars(gkr, &top_vars)) return rc;
  if (top_vars != TOP) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove_hostloop: shard sessions are driven by the caller (sharded.py)");
  std::vector<uint32_t> tn(32 * 5), td(32 * 5);
  if (int rc = lm_gkr_top(gkr, tn.data(), td.data())) return rc;
  fs->add_scalars(tn.data(), tn.size());
  fs->add_scalars(td.data(), td.size());
  std::vector<Ef> top_n, top_d;
  Ef quotient = EF_ZERO;
  for (int i = 0; i < 32; i++) {
    top_n.push_back(ef_load(tn.data() + 5 * i));
    top_d.push_back(ef_load(td.data() + 5 * i));
    Ef inv;
    if (!ef_inv(top_d[i], &inv)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove_hostloop: a top-layer denominator is zero");
    quotient = lm::ef_add(quotient, lm::ef_mul(top_n[i], inv));
  }
  std::vector<Ef> point;
  if (!fs->sample_ef(TOP, &point)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove_hostloop: stale rate");
  auto mle_small = [&](std::vector<Ef> cur) {
    for (const Ef& x : point) {
      const size_t half = cur.size() / 2;
      for (size_t i = 0; i < half; i++) cur[i] = lm::ef_add(cur[i], lm::ef_mul(x, lm::ef_sub(cur[i + half], cur[i])));
      cur.resize(half);
    }
    return cur[0];
  };
  Ef claim_num = mle_small(top_n), claim_den = mle_small(top_d);

  for (uint32_t k = TOP; k < n_vars; k++) {
    // prove_gkr_layer (mod.rs:80-141): alpha after a duplex
    fs->duplex();
    std::vector<Ef> tmp;
    if (!fs->sample_ef(1, &tmp)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: stale rate");
    const Ef alpha = tmp[0];
    Ef s = lm::ef_add(claim_num, lm::ef_mul(alpha, claim_den)), mmf = EF_ONE;
    std::vector<uint32_t> pt(5 * (size_t)k);
    for (uint32_t i = 0; i < k; i++) ef_store(pt.data() + 5 * i, point[i]);
    if (int rc = lm_gkr_layer_begin(gkr, k, pt.data(), alpha.c)) return rc;
    std::vector<Ef> q;
    for (uint32_t rnd = 0; rnd < k; rnd++) {
      uint32_t c0r[5], c2r[5];
      if (int rc = lm_gkr_round(gkr, c0r, c2r)) return rc;
      const Ef eq_alpha = point[k - 1 - rnd];
      // build_bare_from_coeffs (sumcheck_utils.rs:491-503)
      const Ef c0 = lm::ef_mul(ef_load(c0r), mmf), c2 = lm::ef_mul(ef_load(c2r), mmf);
      Ef ainv;
      if (!ef_inv(eq_alpha, &ainv)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: eq point coordinate is zero");
      const Ef h1 = lm::ef_mul(lm::ef_sub(s, lm::ef_mul(lm::ef_sub(EF_ONE, eq_alpha), c0)), ainv);
      const std::vector<Ef> bare = {c0, lm::ef_sub(lm::ef_sub(h1, c0), c2), c2};
      fs->add_sumcheck_polynomial(bare, &eq_alpha);
      tmp.clear();
      if (!fs->sample_ef(1, &tmp)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: stale rate");
      const Ef r = tmp[0];
      const Ef eq_eval = eq1(eq_alpha, r);
      s = lm::ef_mul(eq_eval, poly_eval(bare, r));
      mmf = lm::ef_mul(mmf, eq_eval);
      if (int rc = lm_gkr_fold(gkr, r.c)) return rc;
      q.push_back(r);
    }
    uint32_t inner[20];
    if (int rc = lm_gkr_layer_end(gkr, inner)) return rc;
    fs->add_scalars(inner, 20);
    tmp.clear();
    if (!fs->sample_ef(1, &tmp)) return lm_internal_fail(LM_ERR_INVALID, "lm_gkr_prove: stale rate");
    const Ef beta = tmp[0], omb = lm::ef_sub(EF_ONE, beta);
    const Ef nl = ef_load(inner), nr = ef_load(inner + 5), dl = ef_load(inner + 10), dr = ef_load(inner + 15);
    claim_num = lm::ef_add(lm::ef_mul(omb, nl), lm::ef_mul(beta, nr));
    claim_den = lm::ef_add(lm::ef_mul(omb, dl), lm::ef_mul(beta, dr));
    point.assign(q.rbegin(), q.rend());
    point.push_back(beta);
  }
  ef_store(out_quotient, quotient);
  for (uint32_t i = 0; i < n_vars; i++) ef_store(out_point + 5 * i, point[i]);
  ef_store(out_claim_num, claim_num);
  ef_store(out_claim_den, claim_den);
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ batched AIR sumcheck
int lm_air_prove_batched(lm_air* const* airs, uint32_t n_sessions, const uint32_t* eq_factors, const uint32_t* sums,
                         const uint32_t eta[5], lm_fs* fs, uint32_t* out_challenges, uint32_t* out_n_rounds) {
  if (!airs || !n_sessions || !eq_factors || !sums || !eta || !fs || !out_challenges)
    return lm_internal_fail(LM_ERR_INVALID, "lm_air_prove_batched: null argument");
  struct Sess {
    lm_air* h;
    uint32_t n_vars, degree;
    std::vector<Ef> eq;  // the last element belongs to the variable bound next
    Ef sum, mmf, k;
  };
  std::vector<Sess> S(n_sessions);
  uint32_t n_rounds = 0, max_full_degree = 0;
  const uint32_t* eqp = eq_factors;
  for (uint32_t i = 0; i < n_sessions; i++) {
    uint32_t tot = 0;
    S[i].h = airs[i];
    if (!airs[i]) return lm_internal_fail(LM_ERR_INVALID, "lm_air_prove_batched: null session");
    if (int rc = lm_air_info(airs[i], &S[i].n_vars, &S[i].degree, &tot)) return rc;
    for (uint32_t v = 0; v < S[i].n_vars; v++, eqp += 5) S[i].eq.push_back(ef_load(eqp));
    S[i].sum = ef_load(sums + 5 * i);
    S[i].mmf = S[i].k = EF_ONE;
    if (S[i].n_vars > n_rounds) n_rounds = S[i].n_vars;
    if (S[i].degree + 1 > max_full_degree) max_full_degree = S[i].degree + 1;
  }
  std::vector<Ef> eta_pow(n_sessions, EF_ONE);
  for (uint32_t i = 1; i < n_sessions; i++) eta_pow[i] = lm::ef_mul(eta_pow[i - 1], ef_load(eta));
  std::vector<uint32_t> raw;
  for (uint32_t rnd = 0; rnd < n_rounds; rnd++) {
    std::vector<Ef> combined(max_full_degree + 1, EF_ZERO);
    std::vector<std::vector<Ef>> bare(n_sessions);
    for (uint32_t i = 0; i < n_sessions; i++) {
      Sess& s = S[i];
      const uint32_t join_round = n_rounds - s.n_vars;
      const Ef w = lm::ef_mul(eta_pow[i], s.k);
      if (rnd < join_round) {
        combined[1] = lm::ef_add(combined[1], lm::ef_mul(w, s.sum));
        continue;
      }
      // compute_bare_round_poly (air_sumcheck.rs:225-266)
      raw.resize(5 * (size_t)s.degree);
      if (int rc = lm_air_round(s.h, raw.data())) return rc;
      const Ef alpha = s.eq.back();
      std::vector<Ef> p_evals;
      for (uint32_t z = 0; z < s.degree; z++) p_evals.push_back(lm::ef_mul(ef_load(raw.data() + 5 * z), s.mmf));
      Ef ainv;
      if (!ef_inv(alpha, &ainv)) return lm_internal_fail(LM_ERR_INVALID, "lm_air_prove_batched: eq coordinate is zero");
      const Ef p1 = lm::ef_mul(lm::ef_sub(s.sum, lm::ef_mul(lm::ef_sub(EF_ONE, alpha), p_evals[0])), ainv);
      p_evals.insert(p_evals.begin() + 1, p1);
      bare[i] = lagrange_at_integers(p_evals);
      const std::vector<Ef> full = expand_bare_to_full(bare[i], alpha);
      for (size_t c = 0; c < full.size(); c++) combined[c] = lm::ef_add(combined[c], lm::ef_mul(w, full[c]));
    }
    fs->add_sumcheck_polynomial(combined, nullptr);
    std::vector<Ef> tmp;
    if (!fs->sample_ef(1, &tmp)) return lm_internal_fail(LM_ERR_INVALID, "lm_air_prove_batched: stale rate");
    const Ef ch = tmp[0];
    ef_store(out_challenges + 5 * rnd, ch);
    for (uint32_t i = 0; i < n_sessions; i++) {
      Sess& s = S[i];
      if (rnd < n_rounds - s.n_vars) {
        s.k = lm::ef_mul(s.k, ch);
        continue;
      }
      // process_challenge (air_sumcheck.rs:268-287)
      const Ef alpha = s.eq.back();
      const Ef eq_eval = eq1(alpha, ch);
      s.sum = lm::ef_mul(poly_eval(bare[i], ch), eq_eval);
      s.mmf = lm::ef_mul(s.mmf, eq_eval);
      if (int rc = lm_air_fold(s.h, ch.c)) return rc;
      s.eq.pop_back();
    }
  }
  if (out_n_rounds) *out_n_rounds = n_rounds;
  return LM_OK;
}

}  // extern "C"

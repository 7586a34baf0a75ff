"""ORACLE — TEST INFRASTRUCTURE ONLY.

Host spine of WHIR restated from the reference so that the prove -> verify loop the reference's own tests rely on
(crates/whir/tests/run_whir.rs:21-141) can be closed here:

  Challenger / ProverState / VerifierState   crates/backend/fiat-shamir/src/{challenger,prover,verifier,utils}.rs
  WhirConfig.new (parameter derivation)       crates/whir/src/config.rs:146-617
  commit / prove (CPU, on the oracle kernels) crates/whir/src/commit.rs:64-99, open.rs:37-446,518-584
  verify                                      crates/whir/src/verify.rs:83-435

Merkle-path pruning (fiat-shamir/src/merkle_pruning.rs) is a wire-format compression and is not restated: paths
are carried in full.  PoW grinding returns the SMALLEST valid witness (the reference's rayon find_any returns an
arbitrary one, prover.rs:135-167), which makes transcripts reproducible.
Field elements are Montgomery-form numpy uint32 at every interface; per-value algebra runs on Python ints.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field

import numpy as np

import oracle as O

P = O.P
_R = (1 << 32) % P
_RINV = pow(_R, -1, P)

# ------------------------------------------------------------------------------------------ tiny EF algebra (ints)
ZERO, ONE = (0, 0, 0, 0, 0), (1, 0, 0, 0, 0)


def fm(v):
    return tuple(int(x) * _RINV % P for x in np.asarray(v, dtype=np.uint64).reshape(-1))


def tm(e):
    return np.array([x % P * _R % P for x in e], dtype=np.uint32)


def add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


def sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


def mul(a, b):
    d = [0] * 9
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                d[i + j] += x * y
    return ((d[0] + d[5] - d[8]) % P, (d[1] + d[6]) % P, (d[2] - d[5] + d[7] + d[8]) % P, (d[3] - d[6] + d[8]) % P,
            (d[4] - d[7]) % P)


def scal(a, k):
    return tuple(x * k % P for x in a)


def inv(a):
    return fm(O.ef_inv(tm(a)))


def peval(coeffs, x):
    acc = ZERO
    for c in reversed(coeffs):
        acc = add(mul(acc, x), c)
    return acc


def eq_outside(p, q):
    acc = ONE
    for a, b in zip(p, q):
        acc = mul(acc, add(mul(a, b), mul(sub(ONE, a), sub(ONE, b))))
    return acc


def expand_from_univariate(y, n):
    out, cur = [], y
    for _ in range(n):
        out.append(cur)
        cur = mul(cur, cur)
    return out


def mle_eval_small(values, point):
    cur = list(values)
    for x in point:
        h = len(cur) // 2
        cur = [add(cur[i], mul(x, sub(cur[i + h], cur[i]))) for i in range(h)]
    return cur[0]


# ------------------------------------------------------------------------------------------ Fiat-Shamir
RATE, WIDTH, CAPACITY = 8, 16, 8


class Challenger:
    """Duplex sponge over the Poseidon1 PERMUTATION (challenger.rs:8-76)."""

    def __init__(self):
        self.state = np.zeros(WIDTH, dtype=np.uint32)
        self.rate_fresh = False

    def observe(self, value):
        self.state[CAPACITY:] = value
        self.state = O.poseidon1_permute(self.state)
        self.rate_fresh = True

    def observe_many(self, scalars):
        s = np.asarray(scalars, dtype=np.uint32).reshape(-1)
        for i in range(0, s.size, RATE):
            buf = np.zeros(RATE, dtype=np.uint32)
            chunk = s[i:i + RATE]
            buf[: chunk.size] = chunk
            self.observe(buf)

    def duplex(self):
        self.observe(np.zeros(RATE, dtype=np.uint32))

    def sample(self):
        assert self.rate_fresh, "stale rate. insert a duplex() before."
        self.rate_fresh = False
        return self.state[CAPACITY:].copy()

    def sample_many(self, n):
        out = []
        for i in range(n):
            if i:
                self.duplex()
            out.append(self.sample())
        return out

    def sample_in_range(self, bits, n_samples):
        fes = np.concatenate(self.sample_many(-(-n_samples // RATE))) if n_samples else np.zeros(0, dtype=np.uint32)
        return [int(O.from_monty(fe)) & ((1 << bits) - 1) for fe in fes[:n_samples]]


def _sample_vec(ch: Challenger, n):
    fes = np.concatenate(ch.sample_many(-(-(n * 5) // RATE)))[: n * 5] if n else np.zeros(0, dtype=np.uint32)
    return [fes[5 * i:5 * i + 5].copy() for i in range(n)]


def expand_bare_to_full(bare, alpha):
    a = alpha
    c0, c1 = sub(ONE, a), sub(add(a, a), ONE)
    full = [ZERO] * (len(bare) + 1)
    for i, b in enumerate(bare):
        full[i] = add(full[i], mul(c0, b))
        full[i + 1] = add(full[i + 1], mul(c1, b))
    return full


def grind(state, bits, start=0):
    """smallest canonical witness w >= start with low `bits` bits of lane 8 of permute(capacity | w | 0..) zero"""
    batch = 1 << 12
    base = start
    while True:
        st = np.zeros((batch, WIDTH), dtype=np.uint32)
        st[:, :CAPACITY] = state[:CAPACITY]
        st[:, CAPACITY] = O.to_monty(np.arange(base, base + batch, dtype=np.uint64))
        out = O.poseidon1_permute(st)
        hits = np.nonzero((O.from_monty(out[:, CAPACITY]) & ((1 << bits) - 1)) == 0)[0]
        if hits.size:
            return base + int(hits[0])
        base += batch


class ProverState:
    """prover.rs:28-178 (transcript recording + challenger); `grinder(state, bits) -> witness` may be replaced."""

    def __init__(self, grinder=None):
        self.challenger = Challenger()
        self.transcript: list[int] = []
        self.merkle_paths: list[list] = []
        self.grinder = grinder or grind

    def add_base_scalars(self, scalars):
        s = np.asarray(scalars, dtype=np.uint32).reshape(-1)
        self.challenger.observe_many(s)
        self.transcript.extend(int(x) for x in s)

    def add_extension_scalars(self, scalars):
        self.add_base_scalars(np.asarray(scalars, dtype=np.uint32).reshape(-1))

    def observe_scalars(self, scalars):
        self.challenger.observe_many(scalars)

    def duplex(self):
        self.challenger.duplex()

    def sample_vec(self, n):
        return _sample_vec(self.challenger, n)

    def sample(self):
        return self.sample_vec(1)[0]

    def sample_in_range(self, bits, n):
        return self.challenger.sample_in_range(bits, n)

    def add_sumcheck_polynomial(self, coeffs, eq_alpha=None):
        c = np.asarray(coeffs, dtype=np.uint32).reshape(-1, 5)
        if eq_alpha is None:
            self.challenger.observe_many(c.reshape(-1))
            self.transcript.extend(int(x) for x in c[1:].reshape(-1))  # c0 is reconstructed by the verifier
        else:
            full = expand_bare_to_full([fm(x) for x in c], fm(eq_alpha))
            self.challenger.observe_many(np.concatenate([tm(x) for x in full]))
            self.transcript.extend(int(x) for x in c[1:].reshape(-1))

    def hint_merkle_paths(self, paths):
        """paths: list of (leaf_data words, sibling digests, leaf_index)"""
        self.merkle_paths.append(paths)

    def pow_grinding(self, bits):
        if bits == 0:
            return
        w = self.grinder(self.challenger.state.copy(), bits)
        wm = O.to_monty(np.array([w], dtype=np.uint64))
        self.challenger.observe_many(wm)
        assert int(O.from_monty(self.challenger.state[CAPACITY])) & ((1 << bits) - 1) == 0
        self.transcript.append(int(wm[0]))


class ProofError(Exception):
    pass


class VerifierState:
    """verifier.rs:14-196 without the pruning / raw-transcript bookkeeping."""

    def __init__(self, transcript, merkle_paths):
        self.challenger = Challenger()
        self.transcript = list(transcript)
        self.off = 0
        self.openings = [p for group in merkle_paths for p in group]
        self.open_idx = 0

    def _read(self, n):
        if self.off + n > len(self.transcript):
            raise ProofError("ExceededTranscript")
        out = np.array(self.transcript[self.off:self.off + n], dtype=np.uint32)
        self.off += n
        return out

    def next_base_scalars_vec(self, n):
        s = self._read(n)
        self.challenger.observe_many(s)
        return s

    def next_extension_scalars_vec(self, n):
        return self.next_base_scalars_vec(5 * n).reshape(n, 5)

    def duplex(self):
        self.challenger.duplex()

    def sample_vec(self, n):
        return _sample_vec(self.challenger, n)

    def sample(self):
        return self.sample_vec(1)[0]

    def sample_in_range(self, bits, n):
        return self.challenger.sample_in_range(bits, n)

    def next_merkle_opening(self):
        if self.open_idx >= len(self.openings):
            raise ProofError("ExceededTranscript")
        o = self.openings[self.open_idx]
        self.open_idx += 1
        return o

    def check_pow_grinding(self, bits):
        if bits == 0:
            return
        w = self._read(1)
        self.challenger.observe_many(w)
        if int(O.from_monty(self.challenger.state[CAPACITY])) & ((1 << bits) - 1) != 0:
            raise ProofError("InvalidGrindingWitness")

    def next_sumcheck_polynomial(self, n_coeffs, claimed_sum, eq_alpha=None):
        if eq_alpha is None:
            rest = self._read((n_coeffs - 1) * 5).reshape(-1, 5)
            rest_c = [fm(x) for x in rest]
            tot = ZERO
            for c in rest_c:
                tot = add(tot, c)
            c0 = scal(sub(claimed_sum, tot), pow(2, -1, P))
            self.challenger.observe_many(np.concatenate([tm(c0), rest.reshape(-1)]))
            return [c0] + rest_c
        rest = self._read((n_coeffs - 2) * 5).reshape(-1, 5)
        rest_b = [fm(x) for x in rest]
        tot = ZERO
        for c in rest_b:
            tot = add(tot, c)
        h0 = sub(claimed_sum, mul(eq_alpha, tot))
        full = expand_bare_to_full([h0] + rest_b, eq_alpha)
        self.challenger.observe_many(np.concatenate([tm(x) for x in full]))
        return full


# ------------------------------------------------------------------------------------------ parameters
def _log_eta(log_inv_rate, log_c):
    return -(0.5 * log_inv_rate + log_c)  # Johnson bound


def _list_size_bits(log_inv_rate, log_c):
    return log_inv_rate / 2.0 - (1.0 + _log_eta(log_inv_rate, log_c))


def _prox_gaps_error(log_degree, log_inv_rate, field_bits, num_functions, log_c):
    eta = 2.0 ** _log_eta(log_inv_rate, log_c)
    rho = 1.0 / float(1 << log_inv_rate)
    rho_sqrt = math.sqrt(rho)
    gamma = 1.0 - rho_sqrt - eta
    n = float(1 << (log_degree + log_inv_rate))
    m = max(math.ceil(rho_sqrt / (2.0 * eta)), 3.0)
    num_1 = (2.0 * (m + 0.5) ** 5 + 3.0 * (m + 0.5) * gamma * rho) * n
    den_1 = 3.0 * rho * rho_sqrt
    error = math.log2(num_1 / den_1 + (m + 0.5) / rho_sqrt)
    return field_bits - (error + math.log2(num_functions - 1.0))


def _log_1_delta(log_inv_rate, log_c):
    eta = 2.0 ** _log_eta(log_inv_rate, log_c)
    rate = 1.0 / float(1 << log_inv_rate)
    return math.log2(1.0 - (1.0 - math.sqrt(rate) - eta))


def _queries(level, log_inv_rate, log_c):
    return math.ceil(-level / _log_1_delta(log_inv_rate, log_c))


def _queries_error(log_inv_rate, nq, log_c):
    return -nq * _log_1_delta(log_inv_rate, log_c)


def _ood_samples(level, log_degree, log_inv_rate, field_bits, log_c):
    for s in range(1, 64):
        err = s * field_bits + 1.0 - (2.0 * _list_size_bits(log_inv_rate, log_c) + log_degree * s)
        if err >= level:
            return s
    raise AssertionError("Could not find an appropriate number of OOD samples")


def _folding_pow_bits(level, field_bits, num_variables, log_inv_rate, log_c):
    prox = _prox_gaps_error(num_variables, log_inv_rate, field_bits, 2, log_c)
    sumcheck = field_bits - (_list_size_bits(log_inv_rate, log_c) + 1.0)
    return max(0.0, level - min(prox, sumcheck))


@dataclass
class RoundConfig:
    query_pow_bits: int
    folding_pow_bits: int
    num_queries: int
    ood_samples: int
    log_inv_rate: int
    num_variables: int
    folding_factor: int
    domain_size: int
    folded_domain_gen: int  # Montgomery u32


@dataclass
class WhirConfig:
    """WhirConfig::new with SecurityAssumption::JohnsonBound (config.rs:186-334)."""

    num_variables: int
    security_level: int = 124
    pow_bits: int = 16
    first_folding: int = 7
    subsequent_folding: int = 5
    rs_domain_initial_reduction_factor: int = 5
    max_num_variables_to_send_coeffs: int = 8
    starting_log_inv_rate: int = 1
    field_bits: int = 155  # EF::bits(): bit length of p^5
    round_parameters: list = dc_field(default_factory=list)

    def folding_at(self, r):
        return self.first_folding if r == 0 else self.subsequent_folding

    def total_folding(self, n_rounds):
        return self.first_folding + self.subsequent_folding * n_rounds

    def rs_reduction_factor(self, r):
        return self.rs_domain_initial_reduction_factor if r == 0 else 1

    def _optimal_log_c(self, num_variables, log_inv_rate):
        level = max(self.security_level - self.pow_bits, 0)
        best_m, best_q = 3, None
        for m in range(3, 101):
            log_c = math.log2(2.0 * m)
            if math.ceil(_folding_pow_bits(self.security_level, self.field_bits, num_variables, log_inv_rate, log_c)) > self.pow_bits:
                break
            q = _queries(level, log_inv_rate, log_c)
            if best_q is None or q < best_q:
                best_q, best_m = q, m
        return math.log2(2.0 * best_m)

    def __post_init__(self):
        nv = self.num_variables
        assert 0 < self.first_folding <= nv and 0 < self.subsequent_folding <= nv
        assert self.rs_domain_initial_reduction_factor <= self.first_folding
        level = max(self.security_level - self.pow_bits, 0)
        log_inv_rate = self.starting_log_inv_rate
        domain_size = 1 << (nv + log_inv_rate)
        assert nv + log_inv_rate - self.first_folding <= 24, "Increase folding_factor_0"
        rest = nv - self.first_folding
        if rest < self.max_num_variables_to_send_coeffs:
            num_rounds, self.final_sumcheck_rounds = 0, rest
        else:
            num_rounds = -(-(rest - self.max_num_variables_to_send_coeffs) // self.subsequent_folding)
            self.final_sumcheck_rounds = rest - num_rounds * self.subsequent_folding
        log_c_old = self._optimal_log_c(nv, log_inv_rate)
        self.commitment_ood_samples = _ood_samples(self.security_level, nv, log_inv_rate, self.field_bits, log_c_old)
        self.starting_folding_pow_bits = math.ceil(
            _folding_pow_bits(self.security_level, self.field_bits, nv, log_inv_rate, log_c_old))
        self.round_parameters = []
        nvm = nv - self.first_folding
        for rnd in range(num_rounds):
            rs_red = self.rs_reduction_factor(rnd)
            next_rate = log_inv_rate + (self.folding_at(rnd) - rs_red)
            log_c_new = self._optimal_log_c(nvm, next_rate)
            num_queries = _queries(level, log_inv_rate, log_c_old)
            ood = _ood_samples(self.security_level, nvm, next_rate, self.field_bits, log_c_new)
            query_error = _queries_error(log_inv_rate, num_queries, log_c_old)
            comb_error = self.field_bits - (math.log2(ood + num_queries) + _list_size_bits(next_rate, log_c_new) + 1.0)
            query_pow = max(0.0, self.security_level - min(query_error, comb_error))
            fold_pow = _folding_pow_bits(self.security_level, self.field_bits, nvm, next_rate, log_c_new)
            ff = self.folding_at(rnd)
            gen = O.two_adic_generator(domain_size.bit_length() - 1 - ff)
            self.round_parameters.append(RoundConfig(math.ceil(query_pow), math.ceil(fold_pow), num_queries, ood, log_inv_rate,
                                                     nvm, ff, domain_size, gen))
            nvm -= self.folding_at(rnd + 1)
            log_inv_rate = next_rate
            domain_size >>= rs_red
            log_c_old = log_c_new
        self.final_queries = _queries(level, log_inv_rate, log_c_old)
        self.final_query_pow_bits = math.ceil(max(0.0, self.security_level - _queries_error(log_inv_rate, self.final_queries, log_c_old)))
        self.final_log_inv_rate = log_inv_rate

    @property
    def n_rounds(self):
        return len(self.round_parameters)

    def starting_domain_size(self):
        return 1 << (self.num_variables + self.starting_log_inv_rate)

    def n_vars_of_final_polynomial(self):
        return self.num_variables - self.total_folding(self.n_rounds)

    def final_round_config(self):
        assert self.round_parameters, "no WHIR round (config.rs:422)"
        last = self.round_parameters[-1]
        rs_red = self.rs_reduction_factor(self.n_rounds - 1)
        ff = self.folding_at(self.n_rounds)
        domain_size = last.domain_size >> rs_red
        return RoundConfig(self.final_query_pow_bits, 0, self.final_queries, last.ood_samples, last.log_inv_rate,
                           last.num_variables - ff, ff, domain_size, O.two_adic_generator(domain_size.bit_length() - 1 - ff))


# ------------------------------------------------------------------------------------------ statements
@dataclass
class SparseStatement:
    """crates/whir/src/lib.rs:31-95: point over the inner (low) variables, values = [(selector, value)]"""

    total_num_variables: int
    point: list  # list of EF tuples (canonical ints)
    values: list  # list of (selector, EF tuple)
    is_next: bool = False

    @property
    def inner(self):
        return len(self.point)

    @property
    def selector_vars(self):
        return self.total_num_variables - len(self.point)

    @staticmethod
    def dense(point, value):
        return SparseStatement(len(point), list(point), [(0, value)])


# ------------------------------------------------------------------------------------------ CPU prover (oracle kernels)
class CpuWitness:
    def __init__(self, codeword, layers, full_width, ood_points, ood_answers, dim):
        self.codeword, self.layers, self.full_width = codeword, layers, full_width
        self.ood_points, self.ood_answers, self.dim = ood_points, ood_answers, dim

    def open(self, index):
        return O.merkle_open(self.codeword, self.full_width, self.layers, index)


def _sample_ood(ps, n_samples, n_vars, evaluate):
    pts, answers = [], []
    if n_samples:
        pts = [fm(x) for x in ps.sample_vec(n_samples)]
        for y in pts:
            answers.append(evaluate(expand_from_univariate(y, n_vars)))
        ps.add_extension_scalars(np.concatenate([tm(a) for a in answers]))
    return pts, answers


def _pts(point):
    return np.stack([tm(x) for x in point]) if point else np.zeros((0, 5), dtype=np.uint32)


def cpu_commit(cfg: WhirConfig, ps: ProverState, poly: np.ndarray, actual_len: int) -> CpuWitness:
    """commit.rs:64-99 on a base-field polynomial"""
    nv = cfg.num_variables
    n_blocks = 1 << cfg.first_folding
    block = (1 << nv) // n_blocks
    eff_cols = -(-actual_len // block)
    cw = O.reorder_and_dft(poly, nv, 1, cfg.first_folding, cfg.starting_log_inv_rate, max(eff_cols, 1))
    layers = O.merkle_tree(cw, n_blocks, eff_cols)
    ps.add_base_scalars(layers[-1])
    pts, answers = _sample_ood(ps, cfg.commitment_ood_samples, nv, lambda pt: fm(O.mle_eval(poly, _pts(pt))))
    return CpuWitness(cw, layers, n_blocks, pts, answers, 1)


def combine_statement_cpu(statements, gamma, n_vars):
    """open.rs:518-584 -> (weights table 2^n x 5, combined sum)"""
    w = np.zeros((1 << n_vars, 5), dtype=np.uint32)
    total, gp = ZERO, ONE
    for smt in statements:
        for sel, val in smt.values:
            if smt.is_next:
                O.weights_add_next(w, sel, _pts(smt.point), tm(gp))
            else:
                O.weights_add_eq(w, sel, _pts(smt.point), tm(gp))
            total = add(total, mul(val, gp))
            gp = mul(gp, gamma)
    return w, total


class CpuSumcheck:
    """SumcheckSingle on numpy tables (open.rs:323-446) with the round loop of product_computation.rs / prove.rs."""

    def __init__(self, evals, weights, total):
        self.evals, self.weights, self.sum = evals, weights, total

    def rounds(self, ps, n_rounds, pow_bits):
        chals = []
        for _ in range(n_rounds):
            c0, c2 = O.prod_round(self.evals, self.weights)
            c0, c2 = fm(c0), fm(c2)
            c1 = sub(sub(self.sum, add(c0, c0)), c2)
            ps.add_sumcheck_polynomial(np.stack([tm(c0), tm(c1), tm(c2)]))
            ps.pow_grinding(pow_bits)
            r = ps.sample()
            rr = fm(r)
            self.sum = peval([c0, c1, c2], rr)
            self.evals, self.weights = O.fold_msb(self.evals, r), O.fold_msb(self.weights, r)
            chals.append(rr)
        return chals

    def add_eq(self, point, scalar):
        O.weights_add_eq(self.weights, 0, _pts(point), tm(scalar))

    def add_base_eq(self, points_base, scalars):
        O.weights_add_base_eq(self.weights, points_base, np.stack([tm(s) for s in scalars]))


def _stir_points(gen, indexes, n_vars):
    """expand_from_univariate(gen^i) in the base field: (len(indexes) x n_vars) Montgomery words"""
    pts = np.empty((len(indexes), n_vars), dtype=np.uint32)
    for q, i in enumerate(indexes):
        y, b, e = int(O.to_monty(1)), gen, i
        while e:
            if e & 1:
                y = O.kb_mul(y, b)
            b = O.kb_mul(b, b)
            e >>= 1
        for k in range(n_vars):
            pts[q, k] = y
            y = O.kb_mul(y, y)
    return pts


def _eval_leaf(leaf_words, dim, folding_randomness):
    vals = [fm(leaf_words[5 * i:5 * i + 5]) for i in range(len(leaf_words) // 5)] if dim == 5 else \
        [(int(O.from_monty(x)), 0, 0, 0, 0) for x in leaf_words]
    return mle_eval_small(vals, folding_randomness)


def prove_rounds(cfg: WhirConfig, ps: ProverState, sc, witness, randomness_vec, commit_round, open_witness):
    """open.rs:58-248, shared by the CPU oracle prover and (through the callbacks) by tests of the GPU prover.
    sc: object with rounds/add_eq/add_base_eq and .evals-like accessors supplied by the callbacks:
      commit_round(sc, folding_factor_next, log_inv_rate) -> new witness (root appended by the caller here)
      open_witness(witness, indexes) -> list of (leaf words, path)"""
    domain_size = cfg.starting_domain_size()
    next_domain_gen = O.two_adic_generator(domain_size.bit_length() - 1 - cfg.first_folding)
    for round_index in range(cfg.n_rounds + 1):
        num_variables = cfg.num_variables - cfg.total_folding(round_index)
        if round_index == cfg.n_rounds:
            coeffs = O.evals_to_coeffs(sc.read_evals())
            ps.add_extension_scalars(coeffs.reshape(-1))
            ps.pow_grinding(cfg.final_query_pow_bits)
            idx = ps.sample_in_range((domain_size >> cfg.folding_at(round_index)).bit_length() - 1, cfg.final_queries)
            ps.hint_merkle_paths([(leaf, path, i) for (leaf, path), i in zip(open_witness(witness, idx), idx)])
            if cfg.final_sumcheck_rounds:
                randomness_vec.extend(sc.rounds(ps, cfg.final_sumcheck_rounds, 0))
            return randomness_vec
        rp = cfg.round_parameters[round_index]
        ff_next = cfg.folding_at(round_index + 1)
        new_domain_size = domain_size >> cfg.rs_reduction_factor(round_index)
        inv_rate = new_domain_size >> num_variables
        new_witness = commit_round(sc, ff_next, inv_rate.bit_length() - 1)
        ps.add_base_scalars(new_witness.root)
        ood_points, ood_answers = _sample_ood(ps, rp.ood_samples, num_variables, lambda pt: sc.eval_poly(pt))
        ps.pow_grinding(rp.query_pow_bits)
        idx = ps.sample_in_range((domain_size >> cfg.folding_at(round_index)).bit_length() - 1, rp.num_queries)
        stir_pts = _stir_points(next_domain_gen, idx, num_variables)
        fr = randomness_vec[len(randomness_vec) - cfg.folding_at(round_index):]
        opened = open_witness(witness, idx)
        ps.hint_merkle_paths([(leaf, path, i) for (leaf, path), i in zip(opened, idx)])
        stir_evals = [_eval_leaf(leaf, witness.dim, fr) for leaf, _ in opened]
        ps.duplex()
        gen = fm(ps.sample())
        powers = [ONE]
        for _ in range(len(ood_points) + len(idx)):
            powers.append(mul(powers[-1], gen))
        for k, (y, ans) in enumerate(zip(ood_points, ood_answers)):
            sc.add_eq(expand_from_univariate(y, num_variables), powers[k])
            sc.sum = add(sc.sum, mul(powers[k], ans))
        stir_rand = powers[len(ood_points):len(ood_points) + len(idx)]
        sc.add_base_eq(stir_pts, stir_rand)
        for rnd, ev in zip(stir_rand, stir_evals):
            sc.sum = add(sc.sum, mul(rnd, ev))
        randomness_vec.extend(sc.rounds(ps, ff_next, rp.folding_pow_bits))
        domain_size = new_domain_size
        next_domain_gen = O.two_adic_generator(new_domain_size.bit_length() - 1 - ff_next)
        witness = new_witness
    return randomness_vec


class _CpuSc(CpuSumcheck):
    def read_evals(self):
        return self.evals

    def eval_poly(self, pt):
        return fm(O.mle_eval(self.evals, _pts(pt)))


class _CpuRoundWitness(CpuWitness):
    root = None


def cpu_prove(cfg: WhirConfig, ps: ProverState, statements, witness: CpuWitness, poly: np.ndarray):
    """WhirConfig::prove on the oracle kernels (open.rs:37-56, 448-516)."""
    nv = cfg.num_variables
    stm = [SparseStatement.dense(expand_from_univariate(y, nv), a) for y, a in zip(witness.ood_points, witness.ood_answers)]
    stm += list(statements)
    ps.duplex()
    gamma = fm(ps.sample())
    weights, total = combine_statement_cpu(stm, gamma, nv)
    sc = _CpuSc(poly, weights, total)
    randomness = sc.rounds(ps, cfg.first_folding, cfg.starting_folding_pow_bits)

    def commit_round(s, ff_next, log_inv_rate):
        n = int(math.log2(s.evals.shape[0]))
        cw = O.reorder_and_dft(s.evals, n, 5, ff_next, log_inv_rate, 1 << ff_next)
        layers = O.merkle_tree(cw, 5 << ff_next, 5 << ff_next)
        w = _CpuRoundWitness(cw, layers, 5 << ff_next, [], [], 5)
        w.root = layers[-1]
        return w

    def open_witness(w, idx):
        return [w.open(i) for i in idx]

    return prove_rounds(cfg, ps, sc, witness, randomness, commit_round, open_witness)


# ------------------------------------------------------------------------------------------ verifier
def _parse_commitment(vs, n_vars, ood_samples):
    root = vs.next_base_scalars_vec(8)
    pts, answers = [], []
    if ood_samples:
        pts = [fm(x) for x in vs.sample_vec(ood_samples)]
        answers = [fm(x) for x in vs.next_extension_scalars_vec(ood_samples)]
    return dict(n_vars=n_vars, root=root, ood_points=pts, ood_answers=answers)


def parse_commitment(cfg, vs):
    return _parse_commitment(vs, cfg.num_variables, cfg.commitment_ood_samples)


def _oods(c):
    return [SparseStatement.dense(expand_from_univariate(y, c["n_vars"]), a) for y, a in zip(c["ood_points"], c["ood_answers"])]


def _combine(vs, claimed, constraints):
    gen = fm(vs.sample())
    rand = [ONE]
    for smt in constraints:
        for _, val in smt.values:
            claimed = add(claimed, mul(rand[-1], val))
            rand.append(mul(rand[-1], gen))
    rand.pop()
    return rand, claimed


def _verify_sumcheck_rounds(vs, claimed, rounds, pow_bits):
    rs = []
    for _ in range(rounds):
        coeffs = vs.next_sumcheck_polynomial(3, claimed)
        vs.check_pow_grinding(pow_bits)
        r = fm(vs.sample())
        claimed = peval(coeffs, r)
        rs.append(r)
    return rs, claimed


def _next_mle(x, y):
    n = len(x)
    eq_prefix = [ONE]
    for i in range(n):
        eq_prefix.append(mul(eq_prefix[i], add(mul(x[i], y[i]), mul(sub(ONE, x[i]), sub(ONE, y[i])))))
    low = [ONE] * (n + 1)
    for i in range(n - 1, -1, -1):
        low[i] = mul(mul(low[i + 1], x[i]), sub(ONE, y[i]))
    s = ZERO
    for arr in range(n):
        s = add(s, mul(mul(eq_prefix[arr], mul(sub(ONE, x[arr]), y[arr])), low[arr + 1]))
    allp = ONE
    for v in list(x) + list(y):
        allp = mul(allp, v)
    return add(s, allp)


def _verify_stir(cfg, vs, params: RoundConfig, commitment, folding_randomness, round_index):
    vs.check_pow_grinding(params.query_pow_bits)
    height = params.domain_size >> params.folding_factor
    idx = vs.sample_in_range(height.bit_length() - 1, params.num_queries)
    dim = 1 if round_index == 0 else 5
    folds = []
    for i in idx:
        leaf, path, _ = vs.next_merkle_opening()
        if not O.merkle_verify(commitment["root"], height.bit_length() - 1, i, leaf, path):
            raise ProofError("InvalidProof: merkle")
        if len(leaf) != dim << params.folding_factor:
            raise ProofError("InvalidProof: leaf width")
        folds.append(_eval_leaf(leaf, dim, folding_randomness))
    out = []
    for i, v in zip(idx, folds):
        pt = _stir_points(params.folded_domain_gen, [i], params.num_variables)[0]
        out.append(SparseStatement.dense([(int(O.from_monty(x)), 0, 0, 0, 0) for x in pt], v))
    return out


def verify(cfg: WhirConfig, vs: VerifierState, commitment, statements):
    """verify.rs:83-219"""
    round_constraints, round_rand = [], []
    claimed = ZERO
    prev = commitment
    vs.duplex()
    constraints = _oods(prev) + list(statements)
    rand, claimed = _combine(vs, claimed, constraints)
    round_constraints.append((rand, constraints))
    fr, claimed = _verify_sumcheck_rounds(vs, claimed, cfg.first_folding, cfg.starting_folding_pow_bits)
    round_rand.append(fr)
    for r in range(cfg.n_rounds):
        rp = cfg.round_parameters[r]
        new = _parse_commitment(vs, rp.num_variables, rp.ood_samples)
        stir = _verify_stir(cfg, vs, rp, prev, round_rand[-1], r)
        constraints = _oods(new) + stir
        vs.duplex()
        rand, claimed = _combine(vs, claimed, constraints)
        round_constraints.append((rand, constraints))
        fr, claimed = _verify_sumcheck_rounds(vs, claimed, cfg.folding_at(r + 1), rp.folding_pow_bits)
        round_rand.append(fr)
        prev = new
    n_final = 1 << cfg.n_vars_of_final_polynomial()
    final_coeffs = [fm(x) for x in vs.next_extension_scalars_vec(n_final)]
    stir = _verify_stir(cfg, vs, cfg.final_round_config(), prev, round_rand[-1], cfg.n_rounds)
    for c in stir:
        alpha = c.point[0]
        if peval(final_coeffs, alpha) != c.values[0][1]:
            raise ProofError("InvalidProof: final stir")
    final_r, claimed = _verify_sumcheck_rounds(vs, claimed, cfg.final_sumcheck_rounds, 0)
    round_rand.append(final_r)
    point = [x for rr in round_rand for x in rr]
    # eval_constraints_poly (verify.rs:358-398)
    value = ZERO
    pt = point
    for rnd, (rand, constraints) in enumerate(round_constraints):
        if rnd > 0:
            pt = pt[cfg.folding_at(rnd - 1):]
        i = 0
        for smt in constraints:
            inner_pt = pt[len(pt) - smt.inner:]
            common = _next_mle(smt.point, inner_pt) if smt.is_next else eq_outside(smt.point, inner_pt)
            sv = smt.selector_vars
            for sel, _ in smt.values:
                e = common
                for j in range(sv):
                    e = mul(e, pt[j] if sel & (1 << (sv - 1 - j)) else sub(ONE, pt[j]))
                value = add(value, mul(e, rand[i]))
                i += 1
        assert i == len(rand)
    # final value: multilinear in coefficient form at the reversed point (evals.rs eval_multilinear_coeffs)
    rev = final_r[::-1]
    cur = list(final_coeffs)
    for x in rev:  # coefficients ordered with the first variable as the most significant index bit
        h = len(cur) // 2
        cur = [add(cur[i], mul(x, cur[i + h])) for i in range(h)]
    final_value = cur[0]
    if claimed != mul(value, final_value):
        raise ProofError("InvalidProof: final sumcheck")
    return point


def sumcheck_verify(vs: VerifierState, n_vars: int, degree: int, expected_sum, eq_alphas=None):
    """crates/backend/sumcheck/src/verify.rs:5-30 -> (challenges, final target)"""
    target, chals = expected_sum, []
    for rnd in range(n_vars):
        coeffs = vs.next_sumcheck_polynomial(degree + 1, target, None if eq_alphas is None else eq_alphas[rnd])
        r = fm(vs.sample())
        chals.append(r)
        target = peval(coeffs, r)
    return chals, target

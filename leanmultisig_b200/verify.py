"""Verifier side of the WHIR opening on top of the C ABI: proof bytes -> restored openings -> accept / reject.

Reference call sites mirrored (names kept):
  VerifierState::new / restore_merkle_paths   crates/backend/fiat-shamir/src/verifier.rs:14-106
  PrunedMerklePaths::restore                  crates/backend/fiat-shamir/src/merkle_pruning.rs:86-164
  FSVerifier (next_*, check_pow_grinding, next_sumcheck_polynomial)   verifier.rs:108-196
  WhirConfig::parse_commitment / verify       crates/whir/src/verify.rs:66-219
  verify_stir_challenges                      crates/whir/src/verify.rs:229-345
  eval_constraints_poly                       crates/whir/src/verify.rs:358-398
What runs on the device: every Poseidon1 hash of the verifier.  `restore` needs the leaf digests and the subtree hashes of
the pruned paths — computed level by level for all paths of a query batch at once (`DeviceHasher`: one batched compression
launch per level, lm_dev_poseidon1) — and `verify_stir_challenges` checks all openings of a round against the round's root
and folds every leaf at the folding randomness in ONE call (lm_verify_openings).  The transcript sponge and the field
arithmetic of the (few hundred) scalar checks stay on the host, as in the prover mirror.
There is no CPU path for the hashes: the hasher and the opening check take a `Context` (tests of the host logic on machines
without a GPU inject doubles built on the oracle).
"""
from __future__ import annotations

import numpy as np

from . import field as F
from .fiat_shamir import CAPACITY, RATE, Challenger
from .merkle_pruning import PrunedMerklePaths, lca_level


class ProofError(Exception):
    """fiat-shamir/src/errors.rs: InvalidProof / ExceededTranscript / InvalidGrindingWitness"""


# ------------------------------------------------------------------------------------------ hashes on the device
class DeviceHasher:
    """hash_slice (symetric/src/sponge.rs:7-25) of many equal-width leaves and compress (compression.rs) of many digest
    pairs, each as batched launches of the Poseidon1 compression kernel."""

    def __init__(self, ctx):
        self.ctx = ctx

    def compress_pairs(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        st = np.concatenate([left, right], axis=1).astype(np.uint32)
        return self.ctx.poseidon1(st, compress=True)[:, :8].copy()

    def hash_leaves(self, rows: np.ndarray) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        n, w = rows.shape
        if w < 16 or w % 8:
            raise ProofError("InvalidProof: leaf width")
        st = self.ctx.poseidon1(rows[:, w - 16:], compress=True)
        for c in range(w // 8 - 3, -1, -1):
            st[:, 8:] = rows[:, 8 * c:8 * c + 8]
            st = self.ctx.poseidon1(st, compress=True)
        return st[:, :8].copy()


def restore(pruned: PrunedMerklePaths, hasher):
    """PrunedMerklePaths::restore, the backward pass run level-synchronously: at level l every path that still climbs combines
    its node with the stored sibling or with its right neighbour's level-l subtree hash — one batched compression per level.
    -> list of (leaf_index, row, siblings[height x 8]) in the ORIGINAL query order, or None (malformed)."""
    n, h = len(pruned.paths), pruned.merkle_height
    if h >= 32 or pruned.n_trailing_zeros > 1024 or n == 0 or len(pruned.leaf_data) != n:
        return None if n else []
    width = {len(d) for d in pruned.leaf_data}
    if len(width) != 1:
        return None
    rows = np.zeros((n, width.pop() + pruned.n_trailing_zeros), dtype=np.uint32)
    for i, d in enumerate(pruned.leaf_data):
        rows[i, :len(d)] = np.asarray(d, dtype=np.uint32)
    idx = [int(p[0]) for p in pruned.paths]
    if any(i >= (1 << h) for i in idx) or any(idx[i] == idx[i + 1] for i in range(n - 1)):
        return None
    levels = [h] + [lca_level(idx[i - 1], idx[i]) for i in range(1, n)]
    skip = [lca_level(idx[i], idx[i + 1]) - 1 for i in range(n - 1)] + [None]
    stored = [list(p[1]) for p in pruned.paths]
    for i in range(n):  # a path keeps one sibling per climbed level except the skipped one
        if len(stored[i]) != levels[i] - (1 if skip[i] is not None and skip[i] < levels[i] else 0):
            return None
    try:
        cur = hasher.hash_leaves(rows)
    except ProofError:
        return None
    subtree = [[cur[i]] for i in range(n)]
    taken = [0] * n
    sibs = [[] for _ in range(n)]
    for lvl in range(max(levels)):
        act = [i for i in range(n) if lvl < levels[i]]
        sib = np.empty((len(act), 8), dtype=np.uint32)
        for k, i in enumerate(act):
            if skip[i] == lvl:
                if lvl >= len(subtree[i + 1]):
                    return None
                sib[k] = subtree[i + 1][lvl]
            else:
                sib[k] = np.asarray(stored[i][taken[i]], dtype=np.uint32)
                taken[i] += 1
            sibs[i].append(sib[k].copy())
        node = np.stack([subtree[i][lvl] for i in act])
        right = np.array([(idx[i] >> lvl) & 1 for i in act], dtype=bool)[:, None]
        out = hasher.compress_pairs(np.where(right, sib, node), np.where(right, node, sib))
        for k, i in enumerate(act):
            subtree[i].append(out[k])
    restored = []
    for i in range(n):  # forward pass: the levels above the fork are the previous path's
        s = sibs[i] + (restored[-1][2][levels[i]:] if restored else [])
        if len(s) != h:
            return None
        restored.append((idx[i], rows[i], s))
    out = []
    for pos in pruned.original_order:
        if pos >= n:
            return None
        li, row, s = restored[pos]
        out.append((li, row, np.stack(s) if s else np.zeros((0, 8), dtype=np.uint32)))
    return out


# ------------------------------------------------------------------------------------------ transcript, verifier side
def _expand_bare_to_full(bare, alpha):
    """fiat-shamir/src/utils.rs:30-41: (1 - alpha + (2 alpha - 1) X) * bare(X)"""
    a = F.sub(F.ONE, alpha)
    b = F.sub(F.scal(alpha, 2), F.ONE)
    full = [F.ZERO] * (len(bare) + 1)
    for i, c in enumerate(bare):
        full[i] = F.add(full[i], F.mul(a, c))
        full[i + 1] = F.add(full[i + 1], F.mul(b, c))
    return full


class VerifierState:
    """verifier.rs:14-196 (without the raw-transcript bookkeeping of the recursion program).  `proof`: wire.Proof."""

    def __init__(self, proof, hasher):
        self.challenger = Challenger()
        self.transcript = np.ascontiguousarray(proof.transcript, dtype=np.uint32)
        self.off = 0
        self.openings = []
        for pruned in proof.merkle_paths:
            batch = restore(pruned, hasher)
            if batch is None:
                raise ProofError("InvalidProof: merkle paths do not restore")
            self.openings.extend(batch)
        self.open_idx = 0

    def _read(self, n: int) -> np.ndarray:
        if self.off + n > self.transcript.size:
            raise ProofError("ExceededTranscript")
        out = self.transcript[self.off:self.off + n]
        self.off += n
        return out

    def next_base_scalars_vec(self, n: int) -> np.ndarray:
        s = self._read(n)
        self.challenger.observe_many(s)
        return s

    def next_extension_scalars_vec(self, n: int) -> list:
        return [F.from_monty(x) for x in self.next_base_scalars_vec(5 * n).reshape(n, 5)]

    def duplex(self) -> None:
        self.challenger.duplex()

    def sample_vec(self, n: int) -> list:
        fes = self.challenger.sample_many(-(-(n * 5) // RATE))[: n * 5]
        return [F.from_monty(fes[5 * i:5 * i + 5]) for i in range(n)]

    def sample(self):
        return self.sample_vec(1)[0]

    def sample_in_range(self, bits: int, n: int) -> list:
        return self.challenger.sample_in_range(bits, n)

    def next_merkle_opening(self):
        if self.open_idx >= len(self.openings):
            raise ProofError("ExceededTranscript")
        o = self.openings[self.open_idx]
        self.open_idx += 1
        return o

    def check_pow_grinding(self, bits: int) -> None:
        if bits == 0:
            return
        self.challenger.observe_many(self._read(1))
        if (int(self.challenger.state[CAPACITY]) * F._RINV % F.P) & ((1 << bits) - 1):
            raise ProofError("InvalidGrindingWitness")

    def next_sumcheck_polynomial(self, n_coeffs: int, claimed_sum, eq_alpha=None) -> list:
        if eq_alpha is None:
            rest = self._read((n_coeffs - 1) * 5).reshape(-1, 5)
            rest_c = [F.from_monty(x) for x in rest]
            tot = F.ZERO
            for c in rest_c:
                tot = F.add(tot, c)
            c0 = F.scal(F.sub(claimed_sum, tot), pow(2, -1, F.P))  # h(0) + h(1) = claimed_sum
            self.challenger.observe_many(np.concatenate([F.to_monty(c0), rest.reshape(-1)]))
            return [c0] + rest_c
        rest_b = [F.from_monty(x) for x in self._read((n_coeffs - 2) * 5).reshape(-1, 5)]
        tot = F.ZERO
        for c in rest_b:
            tot = F.add(tot, c)
        full = _expand_bare_to_full([F.sub(claimed_sum, F.mul(eq_alpha, tot))] + rest_b, eq_alpha)
        self.challenger.observe_many(np.concatenate([F.to_monty(x) for x in full]))
        return full


# ------------------------------------------------------------------------------------------ WHIR verify
class Statement:
    """SparseStatement on the verifier (crates/whir/src/lib.rs:31-95): canonical tuples."""

    def __init__(self, total_num_variables: int, point, values, is_next: bool = False):
        self.total_num_variables, self.point, self.values, self.is_next = total_num_variables, list(point), list(values), is_next

    @staticmethod
    def dense(point, value) -> "Statement":
        return Statement(len(point), point, [(0, value)])


def _expand_from_univariate(y, n: int) -> list:
    """poly/src/point.rs:51-61: (y, y^2, y^4, ..., y^(2^(n-1)))"""
    out, cur = [], y
    for _ in range(n):
        out.append(cur)
        cur = F.mul(cur, cur)
    return out


def _eq_outside(p, q):
    acc = F.ONE
    for a, b in zip(p, q):
        ab = F.mul(a, b)
        acc = F.mul(acc, F.sub(F.add(F.scal(ab, 2), F.ONE), F.add(a, b)))  # a b + (1 - a)(1 - b)
    return acc


def _next_mle(x, y):
    """poly/src/next_mle.rs: the multilinear indicator of y = x + 1 (and x = y = all ones)"""
    n = len(x)
    eq_prefix = [F.ONE]
    for i in range(n):
        eq_prefix.append(F.mul(eq_prefix[i], F.add(F.mul(x[i], y[i]), F.mul(F.sub(F.ONE, x[i]), F.sub(F.ONE, y[i])))))
    low = [F.ONE] * (n + 1)
    for i in range(n - 1, -1, -1):
        low[i] = F.mul(F.mul(low[i + 1], x[i]), F.sub(F.ONE, y[i]))
    s = F.ZERO
    for arr in range(n):
        s = F.add(s, F.mul(F.mul(eq_prefix[arr], F.mul(F.sub(F.ONE, x[arr]), y[arr])), low[arr + 1]))
    allp = F.ONE
    for v in list(x) + list(y):
        allp = F.mul(allp, v)
    return F.add(s, allp)


class WhirVerifier:
    def __init__(self, ctx, cfg):
        """ctx: anything with `verify_openings(root, log_height, indices, rows, paths, elem_dim=, fold_point=)` — a Context"""
        self.ctx, self.cfg = ctx, cfg

    # verify.rs:21-58 ParsedCommitment::parse
    @staticmethod
    def _parse(vs: VerifierState, n_vars: int, ood_samples: int) -> dict:
        root = vs.next_base_scalars_vec(8).copy()
        pts, answers = [], []
        if ood_samples:
            pts = vs.sample_vec(ood_samples)
            answers = vs.next_extension_scalars_vec(ood_samples)
        return dict(n_vars=n_vars, root=root, ood_points=pts, ood_answers=answers)

    def parse_commitment(self, vs: VerifierState) -> dict:
        return self._parse(vs, self.cfg.num_variables, self.cfg.commitment_ood_samples)

    @staticmethod
    def _oods(c: dict) -> list:
        return [Statement.dense(_expand_from_univariate(y, c["n_vars"]), a) for y, a in zip(c["ood_points"], c["ood_answers"])]

    @staticmethod
    def _combine(vs, claimed, constraints):
        gen = vs.sample()
        rand = [F.ONE]
        for smt in constraints:
            for _, val in smt.values:
                claimed = F.add(claimed, F.mul(rand[-1], val))
                rand.append(F.mul(rand[-1], gen))
        rand.pop()
        return rand, claimed

    @staticmethod
    def _sumcheck_rounds(vs, claimed, rounds: int, pow_bits: int):
        rs = []
        for _ in range(rounds):
            coeffs = vs.next_sumcheck_polynomial(3, claimed)
            vs.check_pow_grinding(pow_bits)
            r = vs.sample()
            claimed = F.poly_eval(coeffs, r)
            rs.append(r)
        return rs, claimed

    def _verify_stir(self, vs, params, commitment, folding_randomness, round_index: int) -> list:
        vs.check_pow_grinding(params.query_pow_bits)
        log_height = (params.domain_size >> params.folding_factor).bit_length() - 1
        idx = vs.sample_in_range(log_height, params.num_queries)
        dim = 1 if round_index == 0 else 5
        width = dim << params.folding_factor
        rows, paths = [], []
        for _ in idx:
            _, row, sibs = vs.next_merkle_opening()
            if len(row) != width or sibs.shape != (log_height, 8):
                raise ProofError("InvalidProof: opening shape")
            rows.append(row)
            paths.append(sibs)
        if not idx:
            return []
        # all openings of the round against its root + the fold of every leaf at the folding randomness: one device call
        ok, folds = self.ctx.verify_openings(commitment["root"], log_height, idx, np.stack(rows), np.stack(paths), elem_dim=dim,
                                             fold_point=np.stack([F.to_monty(r) for r in folding_randomness]))
        if not bool(np.all(ok)):
            raise ProofError("InvalidProof: merkle")
        g = params.folded_domain_gen * F._RINV % F.P
        out = []
        for i, v in zip(idx, folds):
            y = (pow(g, int(i), F.P), 0, 0, 0, 0)
            out.append(Statement.dense(_expand_from_univariate(y, params.num_variables), F.from_monty(v)))
        return out

    def verify(self, vs: VerifierState, commitment: dict, statements) -> list:
        cfg = self.cfg
        for s in statements:
            assert s.total_num_variables == commitment["n_vars"]
        round_constraints, round_rand = [], []
        claimed, prev = F.ZERO, commitment
        vs.duplex()
        constraints = self._oods(prev) + list(statements)
        rand, claimed = self._combine(vs, claimed, constraints)
        round_constraints.append((rand, constraints))
        fr, claimed = self._sumcheck_rounds(vs, claimed, cfg.first_folding, cfg.starting_folding_pow_bits)
        round_rand.append(fr)
        for r in range(cfg.n_rounds):
            rp = cfg.round_parameters[r]
            new = self._parse(vs, rp.num_variables, rp.ood_samples)
            stir = self._verify_stir(vs, rp, prev, round_rand[-1], r)
            constraints = self._oods(new) + stir
            vs.duplex()
            rand, claimed = self._combine(vs, claimed, constraints)
            round_constraints.append((rand, constraints))
            fr, claimed = self._sumcheck_rounds(vs, claimed, cfg.folding_at(r + 1), rp.folding_pow_bits)
            round_rand.append(fr)
            prev = new
        final_coeffs = vs.next_extension_scalars_vec(1 << cfg.n_vars_of_final_polynomial())
        stir = self._verify_stir(vs, cfg.final_round_config(), prev, round_rand[-1], cfg.n_rounds)
        for c in stir:  # verify_constraint_coeffs: the univariate reading of the final polynomial at the domain point
            if F.poly_eval(final_coeffs, c.point[0]) != c.values[0][1]:
                raise ProofError("InvalidProof: final stir")
        final_r, claimed = self._sumcheck_rounds(vs, claimed, cfg.final_sumcheck_rounds, 0)
        round_rand.append(final_r)
        point = [x for rr in round_rand for x in rr]
        # eval_constraints_poly (verify.rs:358-398)
        value, pt = F.ZERO, point
        for rnd, (rand, constraints) in enumerate(round_constraints):
            if rnd > 0:
                pt = pt[cfg.folding_at(rnd - 1):]
            i = 0
            for smt in constraints:
                inner = len(smt.point)
                inner_pt = pt[len(pt) - inner:]
                common = _next_mle(smt.point, inner_pt) if smt.is_next else _eq_outside(smt.point, inner_pt)
                sv = smt.total_num_variables - inner
                for sel, _ in smt.values:
                    e = common
                    for j in range(sv):
                        e = F.mul(e, pt[j] if sel & (1 << (sv - 1 - j)) else F.sub(F.ONE, pt[j]))
                    value = F.add(value, F.mul(e, rand[i]))
                    i += 1
            assert i == len(rand)
        cur = list(final_coeffs)  # eval_multilinear_coeffs at the reversed point
        for x in final_r[::-1]:
            half = len(cur) // 2
            cur = [F.add(cur[k], F.mul(x, cur[k + half])) for k in range(half)]
        if claimed != F.mul(value, cur[0]):
            raise ProofError("InvalidProof: final sumcheck")
        return point

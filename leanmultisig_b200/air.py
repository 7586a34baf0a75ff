"""Host-side mirror of the AIR sumcheck session on top of the C ABI.

Reference names kept: trait OuterSumcheckSession (crates/sub_protocols/src/air_sumcheck.rs:34-42),
AirSumcheckSession::new (:67-126), compute_bare_round_poly (:225-266), process_challenge (:268-287),
final_column_evals (:289-291), prove_batched_air_sumcheck (:636-681), expand_bare_to_full
(crates/backend/fiat-shamir/src/utils.rs:30-41).  The device computes the round sums and the folds; the few
field operations per round that the reference performs on the host are done here on Python ints.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import field as F
from ._lib import check, lib, u32p

EXECUTION_TABLE = 0


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(u32p)


# (n_cols, n_shift, bare degree) per table_id & 0xff: execution/air.rs:42-55, extension_op/air.rs:44-57, poseidon_16/mod.rs:294-314
AIR_SHAPES = {0: (20, 2, 5), 1: (29, 13, 6), 2: (109, 0, 10)}


class OuterSumcheckHost:
    """The host half of trait OuterSumcheckSession shared by the single-GPU and the sharded session: running sum,
    missing_mul_factor, p(1) and the Lagrange step (air_sumcheck.rs:225-287).  Subclasses provide `_raw_round()`
    (degree x 5 round sums over the whole hypercube) and `_fold(challenge)`."""

    def _init_host(self, eq_factor, sum_, initial_n_vars: int, degree: int):
        eq = _u32(eq_factor).reshape(-1, 5)
        assert eq.shape[0] == initial_n_vars
        self._initial_n_vars, self._degree = initial_n_vars, degree
        self.eq_factor = [F.from_monty(e) for e in eq]  # the last element is removed at each round
        self._sum = F.from_monty(sum_)
        self.missing_mul_factor = F.ONE
        self.rounds_done = 0

    # ---- trait OuterSumcheckSession ------------------------------------------------------------------------
    def initial_n_vars(self) -> int:
        return self._initial_n_vars

    def sum(self) -> np.ndarray:
        return F.to_monty(self._sum)

    def bare_degree(self) -> int:
        return self._degree

    def eq_alpha(self) -> np.ndarray:
        return F.to_monty(self.eq_factor[-1])

    def compute_bare_round_poly(self) -> np.ndarray:
        """coefficients (degree + 1) x 5 of the bare round polynomial"""
        raw = self._raw_round()
        p_evals = [F.mul(F.from_monty(v), self.missing_mul_factor) for v in raw]
        alpha = self.eq_factor[-1]
        # p(1) from the running sum: sum = (1 - alpha) p(0) + alpha p(1)
        p_at_1 = F.mul(F.sub(self._sum, F.mul(F.sub(F.ONE, alpha), p_evals[0])), F.inv(alpha))
        p_evals.insert(1, p_at_1)
        coeffs = F.lagrange_interpolation_at_integers(p_evals)
        return np.stack([F.to_monty(c) for c in coeffs])

    def process_challenge(self, challenge, bare_poly) -> None:
        r = F.from_monty(challenge)
        alpha = self.eq_factor[-1]
        eq_eval = F.add(F.mul(F.sub(F.ONE, alpha), F.sub(F.ONE, r)), F.mul(alpha, r))
        coeffs = [F.from_monty(c) for c in _u32(bare_poly).reshape(-1, 5)]
        self._sum = F.mul(F.poly_eval(coeffs, r), eq_eval)
        self.missing_mul_factor = F.mul(self.missing_mul_factor, eq_eval)
        self._fold(_u32(challenge))
        self.rounds_done += 1
        self.eq_factor.pop()


class AirSumcheckSession(OuterSumcheckHost):
    def __init__(self, ctx, table_id: int, columns, eq_factor, sum_, alpha_powers, logup_alphas_eq_poly, bus_beta, *,
                 halo_next_row=None, eq_scale=None, folded_columns=None, device_columns=None):
        """columns: the table's base-field columns (natural row order).  Keyword forms used by the sharded session
        (leanmultisig_b200/sharded.py): `halo_next_row` / `eq_scale` make this the session of ONE row-range shard
        (lm_air_new_shard), `folded_columns` ((n_cols + n_shift) x rows x 5) starts from already folded EF columns
        (lm_air_new_folded)."""
        eq = _u32(eq_factor).reshape(-1, 5)
        ap, la, beta = _u32(alpha_powers).reshape(-1, 5), _u32(logup_alphas_eq_poly).reshape(-1, 5), _u32(bus_beta)
        h = C.c_void_p()
        if device_columns is not None:
            # (device pointer, n_cols): the table's base columns, contiguous, already on the device (lm_air_new_dev) - e.g.
            # the table's segment of the committed stacked polynomial; the execution table borrows them without a copy
            ptr, n_cols = device_columns
            n_vars = eq.shape[0]
            check(lib().lm_air_new_dev(ctx.handle, table_id, C.c_void_p(ptr), n_cols, n_vars, _p(eq), _p(ap), ap.shape[0], _p(la),
                                       la.shape[0], _p(beta), C.byref(h)))
        elif folded_columns is not None:
            fc = _u32(folded_columns)
            assert fc.ndim == 3 and fc.shape[2] == 5
            n = fc.shape[1]
            n_vars = n.bit_length() - 1
            assert n == 1 << n_vars and eq.shape[0] == n_vars
            check(lib().lm_air_new_folded(ctx.handle, table_id, _p(fc), fc.shape[0], n_vars, _p(eq), _p(ap), ap.shape[0],
                                          _p(la), la.shape[0], _p(beta), C.byref(h)))
        else:
            cols = [_u32(c) for c in columns]
            n = cols[0].size
            n_vars = n.bit_length() - 1
            assert all(c.size == n for c in cols) and n == 1 << n_vars and eq.shape[0] == n_vars
            ptrs = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
            if halo_next_row is None and eq_scale is None:
                check(lib().lm_air_new(ctx.handle, table_id, ptrs, len(cols), n_vars, _p(eq), _p(ap), ap.shape[0],
                                       _p(la), la.shape[0], _p(beta), C.byref(h)))
            else:
                halo = None if halo_next_row is None else _u32(halo_next_row)
                scale = F.to_monty(F.ONE) if eq_scale is None else _u32(eq_scale)
                check(lib().lm_air_new_shard(ctx.handle, table_id, ptrs, len(cols), n_vars, _p(eq), _p(ap), ap.shape[0],
                                             _p(la), la.shape[0], _p(beta), None if halo is None else _p(halo), _p(scale),
                                             C.byref(h)))
        self.handle = h
        nv, deg, tot = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().lm_air_info(h, C.byref(nv), C.byref(deg), C.byref(tot)))
        self._n_cols_total = tot.value
        self._init_host(eq, sum_, n_vars, deg.value)

    def _raw_round(self) -> np.ndarray:
        raw = np.empty((self._degree, 5), dtype=np.uint32)
        check(lib().lm_air_round(self.handle, _p(raw)))
        return raw

    def _fold(self, challenge) -> None:
        check(lib().lm_air_fold(self.handle, _p(challenge)))

    def final_column_evals(self) -> np.ndarray:
        out = np.empty((self._n_cols_total, 5), dtype=np.uint32)
        check(lib().lm_air_final(self.handle, _p(out)))
        return out

    def free(self):
        if self.handle:
            check(lib().lm_air_free(self.handle))
            self.handle = None


def expand_bare_to_full(bare_coeffs, alpha) -> list:
    """full(X) = ((1 - alpha) + (2 alpha - 1) X) * bare(X)   (fiat-shamir/src/utils.rs:30-41)"""
    a = F.from_monty(alpha)
    c0, c1 = F.sub(F.ONE, a), F.sub(F.add(a, a), F.ONE)
    bare = [F.from_monty(c) for c in _u32(bare_coeffs).reshape(-1, 5)]
    full = [F.ZERO] * (len(bare) + 1)
    for i, b in enumerate(bare):
        full[i] = F.add(full[i], F.mul(c0, b))
        full[i + 1] = F.add(full[i + 1], F.mul(c1, b))
    return full


def prove_batched_air_sumcheck(sessions, eta, absorb_and_sample):
    """Back-loaded batching of several sessions (air_sumcheck.rs:636-681).  `absorb_and_sample(coeffs)` stands for
    prover_state.add_sumcheck_polynomial + sample and returns the challenge (5 Montgomery words)."""
    n_rounds = max(s.initial_n_vars() for s in sessions)
    max_full_degree = max(s.bare_degree() + 1 for s in sessions)
    eta_c = F.from_monty(eta)
    eta_powers = [F.power(eta_c, i) for i in range(len(sessions))]
    k = [F.ONE] * len(sessions)
    challenges = []
    for rnd in range(n_rounds):
        combined = [F.ZERO] * (max_full_degree + 1)
        bare_polys = [None] * len(sessions)
        for idx, s in enumerate(sessions):
            join_round = n_rounds - s.initial_n_vars()
            w = F.mul(eta_powers[idx], k[idx])
            if rnd < join_round:
                combined[1] = F.add(combined[1], F.mul(w, F.from_monty(s.sum())))
            else:
                bare = s.compute_bare_round_poly()
                for i, c in enumerate(expand_bare_to_full(bare, s.eq_alpha())):
                    combined[i] = F.add(combined[i], F.mul(w, c))
                bare_polys[idx] = bare
        ch = _u32(absorb_and_sample(np.stack([F.to_monty(c) for c in combined])))
        challenges.append(ch)
        for idx, s in enumerate(sessions):
            join_round = n_rounds - s.initial_n_vars()
            if rnd < join_round:
                k[idx] = F.mul(k[idx], F.from_monty(ch))
            elif bare_polys[idx] is not None:
                s.process_challenge(ch, bare_polys[idx])
    return challenges


def prove_batched_air_sumcheck_native(sessions, eta, native_state):
    """prove_batched_air_sumcheck with the round loop in the library's C++ spine (lm_air_prove_batched): the sessions'
    device handles are driven directly, `native_state` is a fiat_shamir.NativeProverState.  Returns the challenges; the
    sessions' `final_column_evals()` are valid afterwards (their Python-side round bookkeeping is not advanced)."""
    n = len(sessions)
    handles = (C.c_void_p * n)(*[s.handle for s in sessions])
    eqs = np.concatenate([np.stack([F.to_monty(e) for e in s.eq_factor]) for s in sessions]).astype(np.uint32)
    sums = np.stack([s.sum() for s in sessions]).astype(np.uint32)
    n_rounds = max(s.initial_n_vars() for s in sessions)
    out = np.empty((n_rounds, 5), dtype=np.uint32)
    nr = C.c_uint32()
    check(lib().lm_air_prove_batched(handles, n, _p(np.ascontiguousarray(eqs)), _p(np.ascontiguousarray(sums)), _p(_u32(eta)),
                                     native_state.handle, _p(out), C.byref(nr)))
    assert nr.value == n_rounds
    return [out[i].copy() for i in range(n_rounds)]


def fill_trace_poseidon_16(ctx, trace) -> None:
    """fill_trace_poseidon_16 (crates/lean_vm/src/tables/poseidon_16/trace_gen.rs:10-43): `trace` is the list of the 109
    base-field columns (numpy uint32, equal length); columns 25.. are overwritten in place from flag_permute and the inputs."""
    assert len(trace) == 109
    n = trace[0].size
    for c in trace:
        assert c.dtype == np.uint32 and c.size == n and c.flags["C_CONTIGUOUS"] and c.flags["WRITEABLE"]
    ptrs = (C.c_void_p * 109)(*[c.ctypes.data for c in trace])
    check(lib().lm_poseidon16_fill_trace(ctx.handle, ptrs, n))

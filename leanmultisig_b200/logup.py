"""Host-side mirror of the Logup quotient-GKR prover on top of the C ABI.

Reference names kept: prove_gkr_quotient / prove_gkr_layer (crates/sub_protocols/src/quotient_gkr/mod.rs:31-141),
build_bare_from_coeffs (quotient_gkr/sumcheck_utils.rs:491-503), finger_print (crates/utils/src/multilinear.rs:76).
The transcript is abstracted as three callables so the tests can drive it with any challenge source:
  add_scalars(list of EF)           prover_state.add_extension_scalars
  add_sumcheck_poly(coeffs, alpha)  prover_state.add_sumcheck_polynomial(&bare.coeffs, Some(eq_alpha))
  sample() -> EF (5 Montgomery words)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import field as F
from ._lib import check, lib, u32p

N_VARS_TO_SEND_GKR_COEFFS = 5


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(u32p)


def finger_print(ctx, data, alphas, c) -> np.ndarray:
    d, al, cc = _u32(data), _u32(alphas).reshape(-1, 5), _u32(c)
    out = np.empty((d.shape[0], 5), dtype=np.uint32)
    check(lib().lm_finger_print(ctx.handle, _p(d), d.shape[0], d.shape[1], _p(al), _p(cc), _p(out)))
    return out


def _mle_eval_small(values, point):
    """values: list of EF tuples (2^k), point: list of EF tuples (x_0 = MSB)"""
    cur = list(values)
    for x in point:
        half = len(cur) // 2
        cur = [F.add(cur[i], F.mul(x, F.sub(cur[i + half], cur[i]))) for i in range(half)]
    return cur[0]


def build_bare_from_coeffs(c0_raw, c2_raw, eq_alpha, sum_, mmf):
    c0 = F.mul(c0_raw, mmf)
    c2 = F.mul(c2_raw, mmf)
    h1 = F.mul(F.sub(sum_, F.mul(F.sub(F.ONE, eq_alpha), c0)), F.inv(eq_alpha))
    return [c0, F.sub(F.sub(h1, c0), c2), c2]


class GkrQuotientProver:
    def __init__(self, ctx, nums, dens):
        n_, d_ = _u32(nums).reshape(-1), _u32(dens).reshape(-1, 5)
        assert n_.size == d_.shape[0]
        h = C.c_void_p()
        check(lib().lm_gkr_new(ctx.handle, _p(n_), _p(d_), n_.size, C.byref(h)))
        self.handle = h
        nv = C.c_uint32()
        check(lib().lm_gkr_num_vars(h, C.byref(nv)))
        self.n_vars = nv.value

    def top(self):
        tn, td = np.empty((32, 5), dtype=np.uint32), np.empty((32, 5), dtype=np.uint32)
        check(lib().lm_gkr_top(self.handle, _p(tn), _p(td)))
        return tn, td

    def prove(self, add_scalars, add_sumcheck_poly, sample):
        """returns (quotient, point, claim_num, claim_den) as Montgomery-form arrays"""
        tn, td = self.top()
        add_scalars(tn)
        add_scalars(td)
        top_n, top_d = [F.from_monty(v) for v in tn], [F.from_monty(v) for v in td]
        quotient = F.ZERO
        for a, b in zip(top_n, top_d):
            quotient = F.add(quotient, F.mul(a, F.inv(b)))
        point = [F.from_monty(sample()) for _ in range(N_VARS_TO_SEND_GKR_COEFFS)]
        claim_num, claim_den = _mle_eval_small(top_n, point), _mle_eval_small(top_d, point)
        for k in range(N_VARS_TO_SEND_GKR_COEFFS, self.n_vars):
            point, claim_num, claim_den = self._prove_layer(k, point, claim_num, claim_den, add_scalars, add_sumcheck_poly,
                                                            sample)
        return (F.to_monty(quotient), np.stack([F.to_monty(x) for x in point]), F.to_monty(claim_num),
                F.to_monty(claim_den))

    def _prove_layer(self, k, point, claim_num, claim_den, add_scalars, add_sumcheck_poly, sample):
        alpha_m = _u32(sample())
        alpha = F.from_monty(alpha_m)
        s = F.add(claim_num, F.mul(alpha, claim_den))
        mmf = F.ONE
        pt = np.stack([F.to_monty(x) for x in point])
        check(lib().lm_gkr_layer_begin(self.handle, k, _p(pt), _p(alpha_m)))
        remaining = list(point)
        q = []
        c0, c2 = np.empty(5, dtype=np.uint32), np.empty(5, dtype=np.uint32)
        for _ in range(k):
            check(lib().lm_gkr_round(self.handle, _p(c0), _p(c2)))
            eq_alpha = remaining[-1]
            bare = build_bare_from_coeffs(F.from_monty(c0), F.from_monty(c2), eq_alpha, s, mmf)
            add_sumcheck_poly(np.stack([F.to_monty(c) for c in bare]), F.to_monty(eq_alpha))
            r_m = _u32(sample())
            r = F.from_monty(r_m)
            eq_eval = F.add(F.mul(F.sub(F.ONE, eq_alpha), F.sub(F.ONE, r)), F.mul(eq_alpha, r))
            s = F.mul(eq_eval, F.poly_eval(bare, r))
            mmf = F.mul(mmf, eq_eval)
            check(lib().lm_gkr_fold(self.handle, _p(r_m)))
            q.append(r)
            remaining.pop()
        q.reverse()
        inner = np.empty((4, 5), dtype=np.uint32)
        check(lib().lm_gkr_layer_end(self.handle, _p(inner)))
        add_scalars(inner)
        beta = F.from_monty(sample())
        nl, nr, dl, dr = (F.from_monty(v) for v in inner)
        omb = F.sub(F.ONE, beta)
        return q + [beta], F.add(F.mul(omb, nl), F.mul(beta, nr)), F.add(F.mul(omb, dl), F.mul(beta, dr))

    def free(self):
        if self.handle:
            check(lib().lm_gkr_free(self.handle))
            self.handle = None

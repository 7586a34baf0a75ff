"""Logup quotient GKR: oracle-level structural checks on CPU; GPU prover against the oracle and the verifier identities."""
import numpy as np
import pytest

import oracle as O
from leanmultisig_b200 import field as F


def padded(nums, dens, n_vars):
    n = 1 << n_vars
    pn = np.zeros(n, dtype=np.uint32)
    pn[: nums.size] = nums
    pd = np.zeros((n, 5), dtype=np.uint32)
    pd[:, 0] = int(O.to_monty(1))
    pd[: dens.shape[0]] = dens
    return pn, pd


def oracle_layers(pn, pd, n_vars):
    layers = [(pn, pd)]
    cur_n, cur_d = pn, pd
    for _ in range(n_vars - 5):
        cur_n, cur_d = O.gkr_layer_up(cur_n, cur_d)
        layers.append((cur_n, cur_d))
    return layers


def test_oracle_layer_up_preserves_the_quotient(rng):
    n_vars, active = 7, 100
    nums, dens = O.random_field(rng, active), O.random_field(rng, (active, 5))
    pn, pd = padded(nums, dens, n_vars)
    direct = F.ZERO
    for a, b in zip(nums, dens):
        direct = F.add(direct, F.mul((int(O.from_monty(a)), 0, 0, 0, 0), F.inv(F.from_monty(b))))
    tn, td = oracle_layers(pn, pd, n_vars)[-1]
    top = F.ZERO
    for a, b in zip(tn, td):
        top = F.add(top, F.mul(F.from_monty(a), F.inv(F.from_monty(b))))
    assert top == direct


def test_oracle_finger_print(rng):
    data = O.random_field(rng, (9, 4))
    al, c = O.random_field(rng, (4, 5)), O.random_field(rng, 5)
    fp = O.finger_print(data, al, c)
    for r in range(9):
        s = F.from_monty(c)
        for i in range(4):
            s = F.sub(s, F.mul(F.from_monty(al[i]), (int(O.from_monty(data[r, i])), 0, 0, 0, 0)))
        assert np.array_equal(fp[r], F.to_monty(s))


class Transcript:
    """Records what the prover sends and hands out seeded challenges (stands in for ProverState)."""

    def __init__(self, seed):
        self.rs = np.random.default_rng(seed)
        self.log = []

    def add_scalars(self, v):
        self.log.append(("scalars", np.array(v, copy=True)))

    def add_sumcheck_poly(self, coeffs, alpha):
        self.log.append(("poly", np.array(coeffs, copy=True), np.array(alpha, copy=True)))

    def sample(self):
        v = O.random_field(self.rs, 5)
        self.log.append(("challenge", v))
        return v


def verify_gkr(log, n_vars):
    """verify_gkr_quotient (quotient_gkr/mod.rs:147-203) replayed over the recorded transcript."""
    it = iter(log)

    def nxt(kind):
        e = next(it)
        assert e[0] == kind, (e[0], kind)
        return e

    tn = [F.from_monty(v) for v in nxt("scalars")[1]]
    td = [F.from_monty(v) for v in nxt("scalars")[1]]
    quotient = F.ZERO
    for a, b in zip(tn, td):
        quotient = F.add(quotient, F.mul(a, F.inv(b)))
    point = [F.from_monty(nxt("challenge")[1]) for _ in range(5)]

    def ev(vals, pt):
        cur = list(vals)
        for x in pt:
            h = len(cur) // 2
            cur = [F.add(cur[i], F.mul(x, F.sub(cur[i + h], cur[i]))) for i in range(h)]
        return cur[0]

    cn, cd = ev(tn, point), ev(td, point)
    for k in range(5, n_vars):
        alpha = F.from_monty(nxt("challenge")[1])
        s = F.add(cn, F.mul(alpha, cd))
        eq_alphas_rev = point[::-1]
        q = []
        for i in range(k):
            _, coeffs, a_m = nxt("poly")
            bare = [F.from_monty(c) for c in coeffs]
            a = eq_alphas_rev[i]
            assert F.from_monty(a_m) == a
            assert len(bare) == 3
            # sum = (1 - a) h(0) + a h(1)
            assert F.add(F.mul(F.sub(F.ONE, a), F.poly_eval(bare, F.ZERO)), F.mul(a, F.poly_eval(bare, F.ONE))) == s
            r = F.from_monty(nxt("challenge")[1])
            eq_eval = F.add(F.mul(F.sub(F.ONE, a), F.sub(F.ONE, r)), F.mul(a, r))
            s = F.mul(eq_eval, F.poly_eval(bare, r))
            q.append(r)
        q.reverse()
        inner = [F.from_monty(v) for v in nxt("scalars")[1]]
        ce = F.add(F.mul(alpha, F.mul(inner[2], inner[3])), F.add(F.mul(inner[0], inner[3]), F.mul(inner[1], inner[2])))
        eqv = F.ONE
        for a, x in zip(point, q):
            eqv = F.mul(eqv, F.add(F.mul(a, x), F.mul(F.sub(F.ONE, a), F.sub(F.ONE, x))))
        assert s == F.mul(eqv, ce), f"layer {k}"
        beta = F.from_monty(nxt("challenge")[1])
        omb = F.sub(F.ONE, beta)
        cn = F.add(F.mul(omb, inner[0]), F.mul(beta, inner[1]))
        cd = F.add(F.mul(omb, inner[2]), F.mul(beta, inner[3]))
        point = q + [beta]
    return quotient, point, cn, cd


@pytest.mark.gpu
@pytest.mark.parametrize("n_vars,active", [(6, 64), (7, 100), (9, 300), (12, 4096), (13, 5000), (15, 20001)])
def test_gpu_gkr_prover(rng, n_vars, active):
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    nums, dens = O.random_field(rng, active), O.random_field(rng, (active, 5))
    pn, pd = padded(nums, dens, n_vars)
    layers = oracle_layers(pn, pd, n_vars)
    prover = lm.GkrQuotientProver(ctx, nums, dens)
    assert prover.n_vars == n_vars
    tn, td = prover.top()
    assert np.array_equal(tn, layers[-1][0]) and np.array_equal(td, layers[-1][1])
    tr = Transcript(11)
    quotient, point, cn, cd = prover.prove(tr.add_scalars, tr.add_sumcheck_poly, tr.sample)
    # (1) the verifier accepts the transcript and derives the same outputs
    vq, vpoint, vcn, vcd = verify_gkr(tr.log, n_vars)
    assert np.array_equal(quotient, F.to_monty(vq))
    assert np.array_equal(point, np.stack([F.to_monty(x) for x in vpoint]))
    assert np.array_equal(cn, F.to_monty(vcn)) and np.array_equal(cd, F.to_monty(vcd))
    # (2) ground truth as in the reference's run_gkr_quotient test (mod.rs:293-301): quotient = sum n_i / d_i and the
    #     final claims are the MLEs of the (padded) inputs at the final point
    direct = F.ZERO
    for a, b in zip(nums, dens):
        direct = F.add(direct, F.mul((int(O.from_monty(a)), 0, 0, 0, 0), F.inv(F.from_monty(b))))
    assert np.array_equal(quotient, F.to_monty(direct))
    assert np.array_equal(cn, O.mle_eval(pn, point)) and np.array_equal(cd, O.mle_eval(pd, point))
    # (3) every round's raw coefficients equal the oracle's for the same challenges (top two layers re-derived)
    polys = [e for e in tr.log if e[0] == "poly"]
    chals = [e[1] for e in tr.log if e[0] == "challenge"]
    alpha0 = chals[5]
    lay_n, lay_d = layers[-2]
    nl, nr, dl, dr = lay_n[0::2], lay_n[1::2], lay_d[0::2], lay_d[1::2]
    pt0 = np.stack(chals[:5])
    c0, c2 = O.gkr_round(O.embed(nl) if nl.ndim == 1 else nl, O.embed(nr) if nr.ndim == 1 else nr, dl, dr, pt0[:4], alpha0)
    s = F.add(F.from_monty(O.mle_eval(layers[-1][0], pt0)), F.mul(F.from_monty(alpha0), F.from_monty(O.mle_eval(layers[-1][1], pt0))))
    from leanmultisig_b200.logup import build_bare_from_coeffs
    bare = build_bare_from_coeffs(F.from_monty(c0), F.from_monty(c2), F.from_monty(pt0[4]), s, F.ONE)
    assert np.array_equal(polys[0][1], np.stack([F.to_monty(c) for c in bare]))
    prover.free()
    ctx.close()


@pytest.mark.gpu
def test_gpu_finger_print(rng):
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    data = O.random_field(rng, (5000, 5))
    al, c = O.random_field(rng, (5, 5)), O.random_field(rng, 5)
    assert np.array_equal(lm.finger_print(ctx, data, al, c), O.finger_print(data, al, c))
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_vars,active", [(6, 40), (7, 100), (11, 2000), (14, 16000), (17, 100000)])
def test_gpu_gkr_native_spine_matches_python_spine(rng, n_vars, active):
    """lm_gkr_prove (round loop + transcript in C++, csrc/spine.cu) produces the transcript and outputs of the Python-driven
    prover and of the oracle's CPU prover; the oracle verifier accepts it."""
    import leanmultisig_b200 as lm
    from oracle import logup as OL
    from oracle import whir as W

    ctx = lm.Context(0, 20)
    nums, dens = O.random_field(rng, active), O.random_field(rng, (active, 5))
    ps_o = W.ProverState()
    q_o, pt_o, cn_o, cd_o = OL.prove_gkr_quotient_cpu(ps_o, nums, dens)
    p1 = lm.GkrQuotientProver(ctx, nums, dens)
    ps_py = lm.ProverState(ctx)
    out_py = p1.prove_with_state(ps_py)
    p1.free()
    p2 = lm.GkrQuotientProver(ctx, nums, dens)
    ps_n = lm.NativeProverState(ctx)
    out_n = p2.prove_native(ps_n)
    p2.free()
    assert ps_n.transcript == ps_py.transcript == ps_o.transcript
    for a, b in zip(out_n, out_py):
        assert np.array_equal(a, b)
    assert np.array_equal(out_n[0], W.tm(q_o)) and np.array_equal(out_n[2], W.tm(cn_o)) and np.array_equal(out_n[3], W.tm(cd_o))
    vs = W.VerifierState(ps_n.transcript, [])
    OL.verify_gkr_quotient(vs, n_vars)
    assert vs.off == len(ps_n.transcript)
    ps_n.free()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_vars,active", [(12, 4000), (20, 1000000), (22, (1 << 22) - 77)])
def test_gpu_gkr_device_challenger_matches_host_loop(rng, n_vars, active):
    """The device-resident challenger (lm_gkr_prove: Poseidon1 duplex, add_sumcheck_polynomial and sample inside the round
    kernels, csrc/devfs.cuh) against the host-driven loop (lm_gkr_prove_hostloop: lm_fs sponge on the host, one
    synchronisation per round) at sizes where every kernel variant (multi-CTA rounds, last-block reduction, one-CTA tails)
    runs: identical transcript, sponge state, point and claims; the final claims are the MLEs of the inputs."""
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    nums, dens = O.random_field(rng, active), O.random_field(rng, (active, 5))
    outs, states = [], []
    for device in (True, False):
        p = lm.GkrQuotientProver(ctx, nums, dens)
        ps = lm.NativeProverState(ctx)
        ps.add_base_scalars(np.arange(11, dtype=np.uint32))
        outs.append(p.prove_native(ps) if device else p.prove_native_hostloop(ps))
        states.append((ps.transcript, ps.state_and_freshness()))
        p.free()
        ps.free()
    assert states[0][0] == states[1][0]
    assert np.array_equal(states[0][1][0], states[1][1][0]) and states[0][1][1] == states[1][1][1]
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    pn, pd = padded(nums, dens, n_vars)
    assert np.array_equal(outs[0][2], O.mle_eval(pn, outs[0][1])) and np.array_equal(outs[0][3], O.mle_eval(pd, outs[0][1]))
    ctx.close()

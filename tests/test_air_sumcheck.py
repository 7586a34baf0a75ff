"""AIR sumcheck of the execution table: oracle self-consistency on CPU, GPU session against the oracle."""
import numpy as np
import pytest

import oracle as O
from leanmultisig_b200 import field as F


def make_exec_trace(rng, log_n, non_padded):
    """Random 20-column trace whose rows >= non_padded repeat one padding row (what the reference's tables do)."""
    n = 1 << log_n
    cols = O.random_field(rng, (20, n))
    pad = O.random_field(rng, 20)
    cols[:, non_padded:] = pad[:, None]
    return cols


def all_columns(cols):
    return np.concatenate([cols, np.stack([O.shift_column(cols[0]), O.shift_column(cols[1])])])


def extra(rng):
    alpha = O.random_field(rng, 5)
    ap = [np.array([int(O.to_monty(1)), 0, 0, 0, 0], dtype=np.uint32)]
    for _ in range(13):
        ap.append(O.ef_mul(ap[-1], alpha))
    return np.stack(ap), O.random_field(rng, (8, 5)), O.random_field(rng, 5)


def oracle_session(cols22, eq_factor, ap, la, beta, challenges):
    """Runs every round with the oracle; returns (raw evals per round, final column evals)."""
    cur = cols22
    raws = []
    L = eq_factor.shape[0]
    for r in range(L):
        raws.append(O.air_exec_round(cur, eq_factor[: L - r - 1], ap, la, beta))
        cur = np.stack([O.fold_lsb(cur[c], challenges[r]) for c in range(22)])
    return raws, cur[:, 0, :]


def test_oracle_air_sumcheck_is_a_valid_sumcheck(rng):
    """Structural pin (the reference's own check, prove_poseidon_16.rs / air_sumcheck): with the true hypercube sum,
    every round satisfies sum = (1-a) p(0) + a p(1) and the final claim equals eq(point) * C(column evals)."""
    L = 5
    cols22 = all_columns(make_exec_trace(rng, L, 19))
    eq_factor = O.random_field(rng, (L, 5))
    ap, la, beta = extra(rng)
    n = 1 << L
    eq_full = O.eq_table(eq_factor)
    total = F.ZERO
    for x in range(n):
        pt = np.zeros((22, 5), dtype=np.uint32)
        pt[:, 0] = cols22[:, x]
        total = F.add(total, F.mul(F.from_monty(eq_full[x]), F.from_monty(O.air_exec_eval(pt, ap, la, beta))))
    challenges = O.random_field(rng, (L, 5))
    raws, finals = oracle_session(cols22, eq_factor, ap, la, beta, challenges)
    s, mmf = total, F.ONE
    for r in range(L):
        alpha = F.from_monty(eq_factor[L - 1 - r])
        p_evals = [F.mul(F.from_monty(v), mmf) for v in raws[r]]
        p1 = F.mul(F.sub(s, F.mul(F.sub(F.ONE, alpha), p_evals[0])), F.inv(alpha))
        vals = [p_evals[0], p1] + p_evals[1:]
        coeffs = F.lagrange_interpolation_at_integers(vals)
        assert len(coeffs) == 6
        for i, v in enumerate(vals):
            assert F.poly_eval(coeffs, (i, 0, 0, 0, 0)) == v
        ch = F.from_monty(challenges[r])
        eq_eval = F.add(F.mul(F.sub(F.ONE, alpha), F.sub(F.ONE, ch)), F.mul(alpha, ch))
        s = F.mul(F.poly_eval(coeffs, ch), eq_eval)
        mmf = F.mul(mmf, eq_eval)
    final_c = F.from_monty(O.air_exec_eval(finals, ap, la, beta))
    assert s == F.mul(mmf, final_c)
    # and the column evaluations are the MLEs at the (reversed) challenge point, shift columns = next-row MLEs
    point = challenges[::-1]
    for c in (0, 7, 19, 20, 21):
        assert np.array_equal(finals[c], O.mle_eval(cols22[c], point))


def test_host_field_matches_oracle(rng):
    a, b = O.random_field(rng, 5), O.random_field(rng, 5)
    assert np.array_equal(F.to_monty(F.mul(F.from_monty(a), F.from_monty(b))), O.ef_mul(a, b))
    assert np.array_equal(F.to_monty(F.inv(F.from_monty(a))), O.ef_inv(a))


@pytest.mark.gpu
@pytest.mark.parametrize("L,non_padded", [(1, 2), (2, 3), (6, 40), (11, 1500), (13, 8192), (15, 30000)])
def test_gpu_air_session_matches_oracle(rng, L, non_padded):
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    cols = make_exec_trace(rng, L, non_padded)
    cols22 = all_columns(cols)
    eq_factor = O.random_field(rng, (L, 5))
    ap, la, beta = extra(rng)
    challenges = O.random_field(rng, (L, 5))
    sum0 = O.random_field(rng, 5)  # any claimed sum: kernel outputs do not depend on it
    sess = lm.AirSumcheckSession(ctx, 0, list(cols), eq_factor, sum0, ap, la, beta)
    assert sess.initial_n_vars() == L and sess.bare_degree() == 5
    raws, finals = oracle_session(cols22, eq_factor, ap, la, beta, challenges)
    s, mmf = F.from_monty(sum0), F.ONE
    for r in range(L):
        bare = sess.compute_bare_round_poly()
        alpha = F.from_monty(eq_factor[L - 1 - r])
        p_evals = [F.mul(F.from_monty(v), mmf) for v in raws[r]]
        p1 = F.mul(F.sub(s, F.mul(F.sub(F.ONE, alpha), p_evals[0])), F.inv(alpha))
        exp = F.lagrange_interpolation_at_integers([p_evals[0], p1] + p_evals[1:])
        assert np.array_equal(bare, np.stack([F.to_monty(c) for c in exp])), f"round {r}"
        sess.process_challenge(challenges[r], bare)
        ch = F.from_monty(challenges[r])
        eq_eval = F.add(F.mul(F.sub(F.ONE, alpha), F.sub(F.ONE, ch)), F.mul(alpha, ch))
        s = F.mul(F.poly_eval(exp, ch), eq_eval)
        mmf = F.mul(mmf, eq_eval)
        assert np.array_equal(sess.sum(), F.to_monty(s))
    assert np.array_equal(sess.final_column_evals(), finals)
    sess.free()
    ctx.close()


@pytest.mark.gpu
def test_gpu_batched_air_sumcheck_two_heights(rng):
    """prove_batched_air_sumcheck with a tall and a short table: the final claims must satisfy the verifier identity."""
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    ap, la, beta = extra(rng)
    sessions, data = [], []
    for L in (9, 6):
        cols = make_exec_trace(rng, L, (1 << L) - 3)
        eqf = O.random_field(rng, (L, 5))
        cols22 = all_columns(cols)
        eq_full = O.eq_table(eqf)
        total = F.ZERO
        for x in range(1 << L):
            pt = np.zeros((22, 5), dtype=np.uint32)
            pt[:, 0] = cols22[:, x]
            total = F.add(total, F.mul(F.from_monty(eq_full[x]), F.from_monty(O.air_exec_eval(pt, ap, la, beta))))
        sessions.append(lm.AirSumcheckSession(ctx, 0, list(cols), eqf, F.to_monty(total), ap, la, beta))
        data.append((L, eqf, total))
    eta = O.random_field(rng, 5)
    transcript = []
    rs = np.random.default_rng(5)

    def absorb_and_sample(coeffs):
        transcript.append(coeffs)
        return O.random_field(rs, 5)

    challenges = lm.prove_batched_air_sumcheck(sessions, eta, absorb_and_sample)
    # verifier side: running claim through the combined polynomials
    eta_c = F.from_monty(eta)
    claim = F.ZERO
    for idx, (L, eqf, total) in enumerate(data):
        claim = F.add(claim, F.mul(F.power(eta_c, idx), total))
    for coeffs, ch in zip(transcript, challenges):
        cs = [F.from_monty(c) for c in coeffs]
        assert F.add(F.poly_eval(cs, F.ZERO), F.poly_eval(cs, F.ONE)) == claim
        claim = F.poly_eval(cs, F.from_monty(ch))
    # final: claim = sum_idx eta^idx * k_idx * eq(eq_factor, point) * C(final evals)
    n_rounds = 9
    exp = F.ZERO
    for idx, ((L, eqf, total), s) in enumerate(zip(data, sessions)):
        k = F.ONE
        for ch in challenges[: n_rounds - L]:
            k = F.mul(k, F.from_monty(ch))
        pt = challenges[n_rounds - L:][::-1]  # natural ordering point
        eqv = F.ONE
        for a, x in zip(eqf, pt):
            a, x = F.from_monty(a), F.from_monty(x)
            eqv = F.mul(eqv, F.add(F.mul(a, x), F.mul(F.sub(F.ONE, a), F.sub(F.ONE, x))))
        cval = F.from_monty(O.air_exec_eval(s.final_column_evals(), ap, la, beta))
        exp = F.add(exp, F.mul(F.mul(F.power(eta_c, idx), k), F.mul(eqv, cval)))
        s.free()
    assert claim == exp
    ctx.close()

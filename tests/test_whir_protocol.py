"""WHIR commit -> prove -> verify, the loop of the reference's own test (crates/whir/tests/run_whir.rs:21-141).

CPU tier: the oracle spine (oracle/whir.py) proves and verifies, rejects tampered proofs, and the host logic of
the product (parameter derivation, transcript sponge, final-coefficient transform) agrees with it.
GPU tier: the product prover (leanmultisig_b200.whir.WhirProver on the device sessions) must produce the SAME
transcript as the oracle prover and be accepted by the oracle verifier.
"""
import numpy as np
import pytest

import oracle as O
from oracle import whir as W

from leanmultisig_b200 import field as F
from leanmultisig_b200 import whir_config as WC


def make_statements(rng, poly, nv, n_sparse=3, with_next=False):
    """run_whir.rs:66-104: sparse points with a few selectors each, plus one all-selector (index) statement"""
    specs = [int(rng.integers(0, nv // 2)) for _ in range(n_sparse)] + [nv]
    stm = []
    for k, sel_len in enumerate(specs):
        inner = nv - sel_len
        pt = [tuple(int(x) for x in rng.integers(0, W.P, 5)) for _ in range(inner)]
        sels = []
        for _ in range(int(rng.integers(1, 5))):
            s = int(rng.integers(0, 1 << sel_len))
            if s not in sels:
                sels.append(s)
        is_next = with_next and k == 0 and inner > 0
        vals = []
        for s in sels:
            sub = poly[s << inner:(s + 1) << inner]
            if inner == 0:
                v = (int(O.from_monty(sub[0])), 0, 0, 0, 0)
            elif is_next:
                w = O.next_mle_folded(W._pts(pt))
                acc = W.ZERO
                subc = O.from_monty(sub)
                for i in range(1 << inner):
                    acc = W.add(acc, W.scal(W.fm(w[i]), int(subc[i])))
                v = acc
            else:
                v = W.fm(O.mle_eval(sub, W._pts(pt)))
            vals.append((s, v))
        stm.append(W.SparseStatement(nv, pt, vals, is_next))
    return stm


def oracle_prove(cfg, poly, stm, actual_len=None):
    ps = W.ProverState()
    wit = W.cpu_commit(cfg, ps, poly, poly.size if actual_len is None else actual_len)
    point = W.cpu_prove(cfg, ps, stm, wit, poly)
    return ps, point


def oracle_verify(cfg, transcript, paths, stm):
    vs = W.VerifierState(transcript, paths)
    pc = W.parse_commitment(cfg, vs)
    return W.verify(cfg, vs, pc, stm)


SMALL = dict(security_level=124, pow_bits=10, first_folding=4, subsequent_folding=3,
             rs_domain_initial_reduction_factor=2, max_num_variables_to_send_coeffs=3, starting_log_inv_rate=1)


# ------------------------------------------------------------------------------------------------ CPU tier
def test_schedule_matches_reference_numbers():
    """SURVEY.md section 8d / BASELINE.md: query schedule of the production parameters"""
    c = WC.WhirConfig(22)
    assert [r.num_queries for r in c.round_parameters] == [230, 74] and c.final_queries == 32
    assert [r.log_inv_rate for r in c.round_parameters] == [1, 3] and c.final_sumcheck_rounds == 5
    c = WC.WhirConfig(28)
    assert [r.num_queries for r in c.round_parameters] == [256, 75, 32] and c.final_queries == 21


@pytest.mark.parametrize("nv", [12, 16, 18, 22, 25, 28])
@pytest.mark.parametrize("params", [(16, 7, 5, 5, 8, 1), (18, 7, 4, 5, 9, 2), (10, 4, 3, 2, 3, 1), (0, 5, 5, 3, 6, 3)])
def test_product_config_equals_oracle_config(nv, params):
    pw, ff, sf, rs, ms, rate = params
    if nv + rate - ff > 24 or nv < ff:
        pytest.skip("outside the two-adic range")
    a = WC.WhirConfig(nv, 124, pw, ff, sf, rs, ms, rate)
    b = W.WhirConfig(nv, 124, pw, ff, sf, rs, ms, rate)
    assert [vars(r) for r in a.round_parameters] == [vars(r) for r in b.round_parameters]
    for k in ("commitment_ood_samples", "starting_folding_pow_bits", "final_queries", "final_query_pow_bits",
              "final_sumcheck_rounds", "final_log_inv_rate"):
        assert getattr(a, k) == getattr(b, k)
    assert vars(a.final_round_config()) == vars(b.final_round_config()) if a.n_rounds else True


def test_two_adic_generators():
    for bits in range(25):
        assert WC.two_adic_generator(bits) == O.two_adic_generator(bits)
    g = F.two_adic_generator(24)
    assert pow(g, 1 << 24, F.P) == 1 and pow(g, 1 << 23, F.P) != 1


@pytest.mark.parametrize("with_next", [False, True])
def test_oracle_prove_verify(with_next):
    rng = np.random.default_rng(5)
    nv = 13
    cfg = W.WhirConfig(nv, **SMALL)
    assert cfg.n_rounds == 2
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv, with_next=with_next)
    ps, point = oracle_prove(cfg, poly, stm)
    assert oracle_verify(cfg, ps.transcript, ps.merkle_paths, stm) == point


def test_oracle_prove_verify_zero_padded_polynomial():
    """commit.rs:68-74: trailing zero columns are neither transformed nor hashed"""
    rng = np.random.default_rng(6)
    nv = 13
    cfg = W.WhirConfig(nv, **SMALL)
    live = (1 << nv) * 5 // 16
    poly = O.random_field(rng, 1 << nv)
    poly[live:] = 0
    stm = make_statements(rng, poly, nv)
    ps, _ = oracle_prove(cfg, poly, stm, live)
    oracle_verify(cfg, ps.transcript, ps.merkle_paths, stm)
    # the same proof as with the full-width commit: skipping zero columns is an optimisation, not a format change
    ps2, _ = oracle_prove(cfg, poly, stm)
    assert ps.transcript == ps2.transcript


def test_oracle_verifier_rejects_tampering():
    rng = np.random.default_rng(7)
    nv = 12
    cfg = W.WhirConfig(nv, **SMALL)
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv)
    ps, _ = oracle_prove(cfg, poly, stm)
    oracle_verify(cfg, ps.transcript, ps.merkle_paths, stm)
    for pos in (0, 9, 30, len(ps.transcript) // 2, len(ps.transcript) - 1):
        bad = list(ps.transcript)
        bad[pos] = (bad[pos] + 1) % W.P
        with pytest.raises(W.ProofError):
            oracle_verify(cfg, bad, ps.merkle_paths, stm)
    # a wrong claimed value
    s0 = stm[0]
    wrong = [W.SparseStatement(nv, s0.point, [(s0.values[0][0], W.add(s0.values[0][1], W.ONE))] + s0.values[1:])] + stm[1:]
    with pytest.raises(W.ProofError):
        oracle_verify(cfg, ps.transcript, ps.merkle_paths, wrong)
    # a modified opened leaf and a modified sibling
    for which in (0, 1):
        paths = [[(leaf.copy(), path.copy(), i) for leaf, path, i in grp] for grp in ps.merkle_paths]
        target = paths[-1][3][which]
        target.reshape(-1)[2] ^= 1
        with pytest.raises(W.ProofError):
            oracle_verify(cfg, ps.transcript, paths, stm)
    # truncated proof
    with pytest.raises(W.ProofError):
        oracle_verify(cfg, ps.transcript[:-3], ps.merkle_paths, stm)


def test_product_challenger_matches_oracle():
    """the host permutation exported by the library (no GPU involved) drives the same duplex sponge"""
    from leanmultisig_b200.fiat_shamir import ProverState

    rng = np.random.default_rng(11)
    a, b = ProverState(), W.ProverState()
    for step in range(40):
        kind = step % 5
        if kind == 0:
            v = O.random_field(rng, int(rng.integers(1, 30)))
            a.add_base_scalars(v), b.add_base_scalars(v)
        elif kind == 1:
            n = int(rng.integers(1, 5))
            for x, y in zip(a.sample_vec(n), b.sample_vec(n)):
                assert np.array_equal(x, y)
        elif kind == 2:
            a.duplex(), b.duplex()
            assert a.sample_in_range(13, 37) == b.sample_in_range(13, 37)
        elif kind == 3:
            c = O.random_field(rng, (3, 5))
            a.add_sumcheck_polynomial(c), b.add_sumcheck_polynomial(c)
        else:
            c, al = O.random_field(rng, (4, 5)), O.random_field(rng, 5)
            a.add_sumcheck_polynomial(c, al), b.add_sumcheck_polynomial(c, al)
        assert np.array_equal(a.challenger.state, b.challenger.state)
    assert a.transcript == b.transcript


def test_product_host_helpers_match_oracle():
    from leanmultisig_b200.whir import evals_to_coeffs

    rng = np.random.default_rng(12)
    for log_n in (0, 1, 3, 8):
        e = O.random_field(rng, (1 << log_n, 5))
        assert np.array_equal(evals_to_coeffs(e), O.evals_to_coeffs(e))
    rows = O.random_field(rng, (6, 32, 5))
    pt = [tuple(int(x) for x in rng.integers(0, F.P, 5)) for _ in range(5)]
    got = F.np_mle_eval_rows(F.np_from_monty(rows), pt)
    for q in range(6):
        assert tuple(int(x) for x in got[q]) == W.fm(O.mle_eval(rows[q], W._pts(pt)))


# ------------------------------------------------------------------------------------------------ GPU tier
def to_product_statements(stm):
    from leanmultisig_b200.whir import SparseStatement

    return [SparseStatement(s.total_num_variables, W._pts(s.point), [(sel, W.tm(v)) for sel, v in s.values], s.is_next)
            for s in stm]


@pytest.fixture(scope="module")
def ctx():
    from leanmultisig_b200.whir import Context

    c = Context(0, 24)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [1, 7, 12, 16, 18])
def test_gpu_pow_grind_smallest_witness(ctx, bits):
    from leanmultisig_b200.fiat_shamir import ProverState

    rng = np.random.default_rng(bits)
    a, b = ProverState(ctx), W.ProverState()
    for _ in range(3):
        v = O.random_field(rng, 11)
        a.add_base_scalars(v), b.add_base_scalars(v)
        a.pow_grinding(bits), b.pow_grinding(bits)
        assert a.transcript == b.transcript
        assert np.array_equal(a.challenger.state, b.challenger.state)


def gpu_prove(ctx, cfg, poly, stm, actual_len=None, native=False):
    """native: the C++ transcript of the library (lm_fs) — the sumcheck phases then run in the spine without returning to
    Python between rounds (lm_whir_sumcheck_rounds)"""
    from leanmultisig_b200.fiat_shamir import NativeProverState, ProverState
    from leanmultisig_b200.whir import WhirProver

    ps = NativeProverState(ctx) if native else ProverState(ctx)
    prover = WhirProver(ctx, cfg)
    wit = prover.commit(ps, poly, actual_len)
    point = prover.prove(ps, to_product_statements(stm), wit)
    wit.free()
    return ps, point


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["small", "small_next", "small_padded", "run_whir", "small_native", "small_next_native",
                                  "run_whir_native"])
def test_gpu_prover_transcript_equals_oracle_and_verifies(ctx, case):
    rng = np.random.default_rng(21)
    native = case.endswith("_native")
    case = case[:-7] if native else case
    if case == "run_whir":  # the parameters of crates/whir/tests/run_whir.rs:34-56
        nv, kw = 18, dict(security_level=124, pow_bits=18, first_folding=7, subsequent_folding=4,
                          rs_domain_initial_reduction_factor=5, max_num_variables_to_send_coeffs=9, starting_log_inv_rate=2)
    else:
        nv, kw = 13, SMALL
    cfg_o, cfg_p = W.WhirConfig(nv, **kw), WC.WhirConfig(nv, **kw)
    poly = O.random_field(rng, 1 << nv)
    live = None
    if case == "small_padded":
        live = (1 << nv) * 3 // 8
        poly[live:] = 0
    stm = make_statements(rng, poly, nv, n_sparse=7 if case == "run_whir" else 3, with_next=case == "small_next")
    ps_g, point_g = gpu_prove(ctx, cfg_p, poly, stm, live, native=native)
    ps_o, point_o = oracle_prove(cfg_o, poly, stm, live)
    assert ps_g.transcript == ps_o.transcript
    assert point_g == point_o
    assert len(ps_g.merkle_paths) == len(ps_o.merkle_paths)
    for ga, oa in zip(ps_g.merkle_paths, ps_o.merkle_paths):
        for (gl, gp, gi), (ol, op, oi) in zip(ga, oa):
            assert gi == oi and np.array_equal(gl, ol) and np.array_equal(gp, op)
    assert oracle_verify(cfg_o, ps_g.transcript, ps_g.merkle_paths, stm) == point_g


@pytest.mark.gpu
def test_gpu_prover_full_size_accepted_by_verifier(ctx):
    """Size-independent property at a size the oracle prover does not reach in seconds: a 2^22-variable opening with
    the production parameters (lean_prover defaults) proved on the GPU must be accepted by the oracle verifier."""
    rng = np.random.default_rng(22)
    nv = 22
    cfg = WC.WhirConfig(nv)
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv, n_sparse=4)
    ps, point = gpu_prove(ctx, cfg, poly, stm)
    assert oracle_verify(W.WhirConfig(nv), ps.transcript, ps.merkle_paths, stm) == point
    bad = list(ps.transcript)
    bad[len(bad) // 3] ^= 2
    with pytest.raises(W.ProofError):
        oracle_verify(W.WhirConfig(nv), bad, ps.merkle_paths, stm)

"""The product's verifier side (leanmultisig_b200/verify.py: VerifierState with restored paths, WhirVerifier) against the
oracle's restatement of verify.rs.

CPU tier: the host logic — level-synchronous restore of pruned Merkle paths, transcript replay, sumcheck / STIR / final
checks — on proofs of the oracle CPU prover, with the two device services (batched Poseidon1 hashes, lm_verify_openings)
replaced by doubles built on the oracle.  GPU tier: the same with the real Context: GPU prover -> bytes -> GPU-hashed
verifier accepts, returns the prover's point, rejects tampering."""
import numpy as np
import pytest

import oracle as O
from oracle import merkle_pruning as OMP
from oracle import whir as W

from leanmultisig_b200 import verify as V
from leanmultisig_b200 import whir_config as WC
from leanmultisig_b200 import wire

from test_whir_protocol import SMALL, make_statements, oracle_prove, oracle_verify


class OracleHasher:
    calls = 0

    def hash_leaves(self, rows):
        OracleHasher.calls += 1
        return np.stack([O.hash_slice(np.ascontiguousarray(r)) for r in rows])

    def compress_pairs(self, left, right):
        OracleHasher.calls += 1
        return O.poseidon1_compress(np.concatenate([left, right], axis=1).astype(np.uint32))[:, :8].copy()


class OracleOpenings:
    def verify_openings(self, root, log_height, indices, rows, paths, elem_dim=1, fold_point=None):
        ok = np.array([O.merkle_verify(root, log_height, int(i), rows[q], paths[q]) for q, i in enumerate(indices)])
        ev = np.stack([O.mle_eval(r if elem_dim == 1 else r.reshape(-1, 5), fold_point) for r in rows])
        return ok, ev


def _statements(stm):
    return [V.Statement(s.total_num_variables, s.point, s.values, s.is_next) for s in stm]


def _verify(proof, cfg, stm, hasher, openings):
    vs = V.VerifierState(proof, hasher)
    ver = V.WhirVerifier(openings, cfg)
    return ver.verify(vs, ver.parse_commitment(vs), _statements(stm)), vs


@pytest.mark.parametrize("nv,with_next", [(10, False), (13, True), (12, False)])
def test_product_verifier_accepts_oracle_proofs_and_matches_the_oracle_verifier(nv, with_next):
    rng = np.random.default_rng(60 + nv)
    cfg_o, cfg_p = W.WhirConfig(nv, **SMALL), WC.WhirConfig(nv, **SMALL)
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv, with_next=with_next)
    ps, point = oracle_prove(cfg_o, poly, stm)
    proof = wire.Proof.decompress(wire.Proof.from_prover_state(ps).compress())
    got, vs = _verify(proof, cfg_p, stm, OracleHasher(), OracleOpenings())
    assert got == point == oracle_verify(cfg_o, ps.transcript, ps.merkle_paths, stm)
    assert vs.off == len(ps.transcript) and vs.open_idx == len(vs.openings)
    # the level-synchronous restore returns exactly the prover's hints, like the reference-order restore of the oracle
    flat = [o for batch in ps.merkle_paths for o in batch]
    assert len(flat) == len(vs.openings)
    for (i, row, sibs), (orow, opath, oi) in zip(vs.openings, flat):
        assert i == oi and np.array_equal(row, orow) and np.array_equal(sibs, opath)
    for pruned in proof.merkle_paths:
        a, b = V.restore(pruned, OracleHasher()), OMP.restore(pruned)
        assert len(a) == len(b)
        for (i1, r1, s1), (i2, r2, s2) in zip(a, b):
            assert i1 == i2 and np.array_equal(r1, r2) and np.array_equal(s1, s2)


def test_restore_batches_its_hashes():
    """one leaf-hash batch + one compression batch per tree level, whatever the number of queries"""
    rng = np.random.default_rng(3)
    nv = 12
    cfg = W.WhirConfig(nv, **SMALL)
    poly = O.random_field(rng, 1 << nv)
    ps, _ = oracle_prove(cfg, poly, make_statements(rng, poly, nv))
    pruned = wire.Proof.from_prover_state(ps).merkle_paths[0]
    OracleHasher.calls = 0
    assert V.restore(pruned, OracleHasher()) is not None
    assert OracleHasher.calls == 1 + pruned.merkle_height and len(pruned.paths) > 8


def test_product_verifier_rejects_tampering():
    rng = np.random.default_rng(77)
    nv = 11
    cfg_o, cfg_p = W.WhirConfig(nv, **SMALL), WC.WhirConfig(nv, **SMALL)
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv)
    ps, point = oracle_prove(cfg_o, poly, stm)
    good = wire.Proof.from_prover_state(ps)
    assert _verify(good, cfg_p, stm, OracleHasher(), OracleOpenings())[0] == point
    n = good.transcript.size
    for pos in (3, n // 3, n // 2, n - 2):                      # root / sumcheck / round data / final coefficients
        bad = wire.Proof.from_postcard(good.to_postcard())
        bad.transcript = bad.transcript.copy()
        bad.transcript[pos] = (int(bad.transcript[pos]) + 1) % O.P
        with pytest.raises(V.ProofError):
            _verify(bad, cfg_p, stm, OracleHasher(), OracleOpenings())
    bad = wire.Proof.from_postcard(good.to_postcard())           # a leaf word
    bad.merkle_paths[0].leaf_data[1] = bad.merkle_paths[0].leaf_data[1].copy()
    bad.merkle_paths[0].leaf_data[1][0] ^= 1
    with pytest.raises(V.ProofError):
        _verify(bad, cfg_p, stm, OracleHasher(), OracleOpenings())
    bad = wire.Proof.from_postcard(good.to_postcard())           # a sibling digest
    k = next(i for i, (_, s) in enumerate(bad.merkle_paths[-1].paths) if len(s))
    bad.merkle_paths[-1].paths[k][1][0] = bad.merkle_paths[-1].paths[k][1][0].copy()
    bad.merkle_paths[-1].paths[k][1][0][5] ^= 8
    with pytest.raises(V.ProofError):
        _verify(bad, cfg_p, stm, OracleHasher(), OracleOpenings())
    bad = wire.Proof.from_postcard(good.to_postcard())           # a dropped sibling: the paths do not restore
    bad.merkle_paths[0].paths[k][1].pop()
    with pytest.raises(V.ProofError):
        _verify(bad, cfg_p, stm, OracleHasher(), OracleOpenings())
    wrong = make_statements(np.random.default_rng(5), poly, nv)  # a claim the proof was not made for
    wrong[0].values[0] = (wrong[0].values[0][0], W.add(wrong[0].values[0][1], W.ONE))
    with pytest.raises(V.ProofError):
        _verify(good, cfg_p, wrong, OracleHasher(), OracleOpenings())


# ------------------------------------------------------------------------------------------------ GPU tier
@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as L

    c = L.Context(0, 22)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["small", "small_next", "run_whir"])
def test_gpu_prove_bytes_gpu_verify(ctx, case):
    from test_whir_protocol import gpu_prove

    rng = np.random.default_rng(31)
    if case == "run_whir":  # the parameters of crates/whir/tests/run_whir.rs:34-56
        nv, kw = 18, dict(security_level=124, pow_bits=18, first_folding=7, subsequent_folding=4,
                          rs_domain_initial_reduction_factor=5, max_num_variables_to_send_coeffs=9, starting_log_inv_rate=2)
    else:
        nv, kw = 13, SMALL
    cfg = WC.WhirConfig(nv, **kw)
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv, n_sparse=7 if case == "run_whir" else 3, with_next=case == "small_next")
    ps, point = gpu_prove(ctx, cfg, poly, stm)
    blob = ps.into_proof().compress()
    proof = wire.Proof.decompress(blob)
    got, vs = _verify(proof, cfg, stm, V.DeviceHasher(ctx), ctx)
    assert got == point == oracle_verify(W.WhirConfig(nv, **kw), ps.transcript, ps.merkle_paths, stm)
    flat = [o for batch in ps.merkle_paths for o in batch]
    for (i, row, sibs), (orow, opath, oi) in zip(vs.openings, flat):
        assert i == oi and np.array_equal(row, orow) and np.array_equal(sibs, opath)
    bad = wire.Proof.decompress(blob)
    bad.merkle_paths[0].leaf_data[0] = bad.merkle_paths[0].leaf_data[0].copy()
    bad.merkle_paths[0].leaf_data[0][3] ^= 1
    with pytest.raises(V.ProofError):
        _verify(bad, cfg, stm, V.DeviceHasher(ctx), ctx)
    bad = wire.Proof.decompress(blob)
    bad.transcript = bad.transcript.copy()
    bad.transcript[bad.transcript.size // 2] ^= 2
    with pytest.raises(V.ProofError):
        _verify(bad, cfg, stm, V.DeviceHasher(ctx), ctx)

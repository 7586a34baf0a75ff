"""GPU tier: verifier-side batch check of Merkle openings (lm_verify_openings) against the oracle's merkle_verify
(crates/backend/symetric/src/merkle.rs:92-122) on openings the ORACLE produced, then on the product's own openings."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as L

    c = L.Context(0, 22)
    yield c
    c.close()


def _oracle_openings(rng, log_h, stored, full, eff, n):
    m = O.random_field(rng, (1 << log_h, stored))
    m[:, eff:] = 0
    layers = O.merkle_tree(m, full, eff)
    idx = rng.integers(0, 1 << log_h, n, dtype=np.uint64)
    idx[0], idx[-1] = 0, (1 << log_h) - 1
    rows = np.empty((n, full), dtype=np.uint32)
    paths = np.empty((n, log_h, 8), dtype=np.uint32)
    for q, i in enumerate(idx):
        rows[q], paths[q] = O.merkle_open(m, full, layers, int(i))
    return layers[-1], idx, rows, paths


@pytest.mark.parametrize("log_h,stored,full,eff", [
    (0, 16, 16, 16), (1, 16, 16, 16), (3, 64, 128, 64), (9, 64, 128, 57), (10, 160, 160, 160), (7, 90, 160, 85),
    (4, 16, 64, 1), (12, 64, 128, 64), (8, 24, 24, 24), (13, 128, 128, 128),
])
def test_verify_matches_oracle_with_tampering(ctx, rng, log_h, stored, full, eff):
    n = 37
    root, idx, rows, paths = _oracle_openings(rng, log_h, stored, full, eff, n)
    assert ctx.verify_openings(root, log_h, idx, rows, paths).all()
    # tamper: a row word, a path word, the index, a non-canonical word; one opening left alone between them
    rows[3, int(rng.integers(0, full))] ^= 1
    rows[5, full - 1] = (int(rows[5, full - 1]) + 1) % O.P
    rows[7, 0] = (int(rows[7, 0]) + O.P - 1) % O.P
    if log_h:
        paths[9, int(rng.integers(0, log_h)), int(rng.integers(0, 8))] ^= 4
        paths[11, log_h - 1, 7] = (int(paths[11, log_h - 1, 7]) + 5) % O.P
        idx[13] ^= 1
        idx[15] ^= np.uint64(1 << (log_h - 1))
    idx[17] += np.uint64(1 << log_h)  # not a leaf of this tree (the oracle would only look at the low bits)
    rows[19, 2] = np.uint32(int(rows[19, 2]) + O.P)  # same residue, not canonical
    got = ctx.verify_openings(root, log_h, idx, rows, paths)
    want = np.array([O.merkle_verify(root, log_h, int(i), rows[q], paths[q]) for q, i in enumerate(idx)])
    want[17] = False
    want[19] = False
    assert np.array_equal(got, want)
    bad = {3, 5, 7, 17, 19} | ({9, 11, 13, 15} if log_h else set())
    assert set(np.nonzero(~got)[0].tolist()) == bad


def test_wrong_root_and_empty(ctx, rng):
    root, idx, rows, paths = _oracle_openings(rng, 6, 64, 128, 64, 9)
    other = root.copy()
    other[7] = (int(other[7]) + 1) % O.P
    assert not ctx.verify_openings(other, 6, idx, rows, paths).any()
    assert ctx.verify_openings(root, 6, idx[:0], rows[:0], paths[:0]).size == 0


@pytest.mark.parametrize("n_vars,folding,dim", [(14, 7, 1), (12, 4, 1), (13, 5, 5)])
def test_product_openings_and_fold(ctx, rng, n_vars, folding, dim):
    """Tree.open_fold -> verify_openings: the prover's rows, paths and STIR answers are accepted and re-derived."""
    import leanmultisig_b200 as L  # noqa: F401

    ev = O.random_field(rng, 1 << n_vars) if dim == 1 else O.random_field(rng, (1 << n_vars, 5))
    tree = ctx.commit(ev, n_vars, folding, 1)
    idx = rng.integers(0, tree.height, 50, dtype=np.uint64)
    pt = O.random_field(rng, (folding, 5))
    rows, paths, evals = tree.open_fold(idx, pt)
    ok, ev2 = ctx.verify_openings(tree.root, tree.log_height, idx, rows, paths, elem_dim=dim, fold_point=pt)
    assert ok.all()
    assert np.array_equal(ev2, evals)
    for q in (0, 17, 49):
        leaf = rows[q] if dim == 1 else rows[q].reshape(-1, 5)
        assert np.array_equal(ev2[q], O.mle_eval(leaf, pt))
        assert O.merkle_verify(tree.root, tree.log_height, int(idx[q]), rows[q], paths[q])
    tree.free()


def test_argument_errors(ctx, rng):
    import leanmultisig_b200 as L

    root, idx, rows, paths = _oracle_openings(rng, 3, 16, 16, 16, 4)
    with pytest.raises(L.LmError):
        ctx.verify_openings(root, 3, idx, rows[:, :12], paths)  # width not >= 16
    with pytest.raises(L.LmError):
        ctx.verify_openings(root, 3, idx, rows, paths, elem_dim=1, fold_point=O.random_field(rng, (3, 5)))  # 2^3 != 16

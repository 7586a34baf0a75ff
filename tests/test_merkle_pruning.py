"""Merkle-path pruning: the product's prover-side `prune` against the oracle's verifier-side
`restore` — the reference's own test scenarios (merkle_pruning.rs tests: basic [5, 1, 3], duplicates, adjacent leaves,
single path, all leaves) on real Poseidon1 trees, plus the GPU openings of a committed tree."""
import numpy as np
import pytest

import oracle as O
from oracle import merkle_pruning as OMP
from leanmultisig_b200.merkle_pruning import lca_level, prune


def open_all(mat, full_w, layers, indices):
    rows, paths = zip(*[O.merkle_open(mat, full_w, layers, i) for i in indices])
    return np.stack(rows), np.stack(paths)


class _OracleHasher:
    """the two batched hash services of the product's level-synchronous restore (leanmultisig_b200/verify.py), on the oracle"""

    def hash_leaves(self, rows):
        return np.stack([O.hash_slice(np.ascontiguousarray(r)) for r in rows])

    def compress_pairs(self, left, right):
        return O.poseidon1_compress(np.concatenate([left, right], axis=1).astype(np.uint32))[:, :8].copy()


def check_roundtrip(indices, rows, paths, root, log_h):
    from leanmultisig_b200.verify import restore as product_restore

    pruned = prune(indices, rows, paths)
    restored = OMP.restore(pruned)
    assert restored is not None and len(restored) == len(indices)
    again = product_restore(pruned, _OracleHasher())  # same result level by level as path by path
    assert again is not None and len(again) == len(restored)
    for (a, b, c), (d, e, f) in zip(again, restored):
        assert a == d and np.array_equal(b, e) and np.array_equal(c, f)
    for q, (li, row, sibs) in enumerate(restored):
        assert li == indices[q] and np.array_equal(row, rows[q]) and np.array_equal(sibs, paths[q])
        assert O.merkle_verify(root, log_h, li, row, sibs)
    return pruned


@pytest.mark.parametrize("log_h,indices", [(3, [5, 1, 3]), (3, [5, 1, 5, 3, 1]), (3, [2, 3]), (3, [6]), (3, list(range(8))),
                                           (6, [63, 0, 32, 31, 33, 1]), (1, [1, 0])])
def test_prune_restore_roundtrip(rng, log_h, indices):
    h, stored, full, eff = 1 << log_h, 24, 48, 24
    mat = O.random_field(rng, (h, stored))
    layers = O.merkle_tree(mat, full, eff)
    rows, paths = open_all(mat, full, layers, indices)
    pruned = check_roundtrip(indices, rows, paths, layers[-1], log_h)
    assert pruned.n_trailing_zeros == full - stored and pruned.merkle_height == log_h
    # fewer digests than the unpruned hint whenever two distinct leaves are opened
    if len(set(indices)) > 1:
        assert pruned.n_digests() < len(set(indices)) * log_h


def test_lca_level():
    assert lca_level(5, 4) == 1 and lca_level(0, 7) == 3 and lca_level(2, 3) == 1 and lca_level(1, 5) == 3


def test_restore_rejects_malformed_hints(rng):
    h, log_h = 8, 3
    mat = O.random_field(rng, (h, 16))
    layers = O.merkle_tree(mat, 16, 16)
    rows, paths = open_all(mat, 16, layers, [1, 6])
    pruned = prune([1, 6], rows, paths)
    from leanmultisig_b200.verify import restore as product_restore

    pruned.paths[0] = (pruned.paths[0][0], pruned.paths[0][1][:-1])   # a digest is missing
    assert OMP.restore(pruned) is None and product_restore(pruned, _OracleHasher()) is None
    pruned = prune([1, 6], rows, paths)
    pruned.paths[1] = (9, pruned.paths[1][1])                         # leaf index outside the tree
    assert OMP.restore(pruned) is None and product_restore(pruned, _OracleHasher()) is None
    pruned = prune([1, 6], rows, paths)
    pruned.original_order = [0, 2]                                    # a query that points at no kept leaf
    assert OMP.restore(pruned) is None and product_restore(pruned, _OracleHasher()) is None
    pruned = prune([1, 6], rows, paths)
    pruned.leaf_data[0] = pruned.leaf_data[0].copy()
    pruned.leaf_data[0][0] ^= 1                                       # tampered leaf: restores, but does not verify
    li, row, sibs = OMP.restore(pruned)[0]
    assert not O.merkle_verify(layers[-1], log_h, li, row, sibs)


@pytest.mark.gpu
def test_gpu_openings_prune_and_restore(rng):
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    n_vars, k, r = 17, 7, 1
    live = 1 << 16
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[:live] = O.random_field(rng, live)
    tree = ctx.commit(ev, n_vars, k, r, actual_len=live)
    indices = [int(x) for x in rng.integers(0, tree.height, 40)] + [7, 7]
    rows, paths = tree.open(indices)
    pruned = check_roundtrip(indices, rows, paths, tree.root, tree.log_height)
    assert pruned.n_trailing_zeros == 64 and pruned.n_digests() < len(set(indices)) * tree.log_height
    tree.free()
    ctx.close()


@pytest.mark.parametrize("log_h,indices,sorted_kept,kept_lens", [
    (3, [5, 1, 3], [1, 3, 5], [2, 2, 3]),                                  # test_prune_and_restore_basic (merkle_pruning.rs:253-298)
    (2, [1, 2], [1, 2], [1, 2]),                                           # two leaves under different parents
    (2, [0, 1], [0, 1], [1, 1]),                                           # test_prune_adjacent_leaves (:301-325)
    (3, list(range(8)), list(range(8)), [2, 1, 1, 1, 2, 1, 1, 1]),         # test_prune_all_leaves (:328-367), 10 digests in all
    (2, [2], [2], [2]),                                                    # test_single_path (:370-382)
    (3, [5, 1, 3, 1], [1, 3, 5], [2, 2, 3]),                               # test_duplicated_paths_preserved (:385-422)
])
def test_prune_keeps_exactly_the_digests_the_reference_counts(rng, log_h, indices, sorted_kept, kept_lens):
    """the sibling counts the reference's tests assert path by path (the numbers are theirs), on Poseidon1 trees"""
    h = 1 << log_h
    mat = O.random_field(rng, (h, 16))
    layers = O.merkle_tree(mat, 16, 16)
    rows, paths = open_all(mat, 16, layers, indices)
    pruned = check_roundtrip(indices, rows, paths, layers[-1], log_h)
    assert [i for i, _ in pruned.paths] == sorted_kept
    assert [len(s) for _, s in pruned.paths] == kept_lens
    if indices == list(range(8)):
        assert pruned.n_digests() == 10


def test_trailing_zeros_stripped(rng):
    """test_trailing_zeros_stripped (merkle_pruning.rs:425-449): leaves that all end in the same run of zeros are sent without it"""
    mat = O.random_field(rng, (8, 16))
    mat[:, 13:] = 0
    mat[:, 12] = np.maximum(mat[:, 12], 1)
    layers = O.merkle_tree(mat, 16, 16)
    rows, paths = open_all(mat, 16, layers, [2, 5, 7])
    pruned = check_roundtrip([2, 5, 7], rows, paths, layers[-1], 3)
    assert pruned.n_trailing_zeros == 3 and all(len(d) == 13 for d in pruned.leaf_data)


def test_prune_restore_random_query_sets(rng):
    """120 random (height, query multiset) cases incl. duplicates, neighbours and full coverage: prune -> oracle restore and
    prune -> product (level-synchronous) restore both return the original openings in the original order"""
    for case in range(120):
        log_h = int(rng.integers(1, 9))
        h = 1 << log_h
        n_q = int(rng.integers(1, min(3 * h, 40) + 1))
        indices = [int(i) for i in rng.integers(0, h, n_q)]
        if case % 7 == 0:
            indices = list(range(h))[: max(1, n_q)]
        mat = O.random_field(rng, (h, 16))
        layers = O.merkle_tree(mat, 16, 16)
        rows, paths = open_all(mat, 16, layers, indices)
        check_roundtrip(indices, rows, paths, layers[-1], log_h)

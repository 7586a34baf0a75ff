"""Structural pins for the parts the reference holds no stored vector for."""
import numpy as np
import pytest

import oracle as O


@pytest.mark.parametrize("full,eff,stored", [(16, 16, 16), (32, 32, 32), (128, 64, 64), (128, 128, 128),
                                             (128, 57, 64), (128, 112, 112), (128, 120, 128), (160, 160, 160),
                                             (160, 85, 90), (64, 8, 16), (64, 1, 16)])
def test_leaf_sponge_prover_form_equals_verifier_form(rng, full, eff, stored):
    h = 8
    mat = O.random_field(rng, (h, stored))
    mat[:, eff:] = 0
    dig = O.first_digest_layer(mat, full, eff)
    for r in range(h):
        row = np.zeros(full, dtype=np.uint32)
        row[:stored] = mat[r]
        assert np.array_equal(dig[r], O.hash_slice(row))


def test_zero_suffix_state_matches_hash_of_zeros():
    for k in (2, 3, 8):
        st = O.zero_suffix_state(k)
        assert np.array_equal(st[:8], O.hash_slice(np.zeros(8 * k, dtype=np.uint32)))


@pytest.mark.parametrize("log_h,full,eff", [(1, 16, 16), (3, 128, 64), (6, 160, 160)])
def test_tree_open_verify(rng, log_h, full, eff):
    h = 1 << log_h
    mat = O.random_field(rng, (h, eff))
    layers = O.merkle_tree(mat, full, eff)
    root = layers[-1]
    # level-by-level definition
    off = 0
    n = h
    while n > 1:
        prev = layers[off:off + n]
        nxt = layers[off + n:off + n + n // 2]
        st = np.concatenate([prev[0::2], prev[1::2]], axis=1)
        assert np.array_equal(O.poseidon1_compress(st)[:, :8], nxt)
        off += n
        n //= 2
    for idx in range(h):
        row, path = O.merkle_open(mat, full, layers, idx)
        assert O.merkle_verify(root, log_h, idx, row, path)
        bad = row.copy()
        bad[0] ^= 1
        assert not O.merkle_verify(root, log_h, idx, bad, path)


@pytest.mark.parametrize("n_vars", [1, 2, 3, 5, 8, 11])
def test_dft_equals_mle_evaluation(rng, n_vars):
    """The reference's own property test (whir/src/dft.rs:582-604): out[i] = P(w^i, w^2i, w^4i, ...)."""
    evals = O.random_field(rng, (1 << n_vars, 5))
    out = O.dft_batch_by_evals(evals)  # EF matrix of width 1 == base matrix of width 5
    g = O.two_adic_generator(n_vars)
    one = int(O.to_monty(1))
    for i in rng.integers(0, 1 << n_vars, size=10).tolist() + [0, (1 << n_vars) - 1]:
        y = one
        for _ in range(i):
            y = O.kb_mul(y, g)
        pt = O.expand_from_univariate(np.array([y, 0, 0, 0, 0], dtype=np.uint32), n_vars)
        assert np.array_equal(out[i], O.mle_eval(evals, pt))


def test_prepare_evals_layout(rng):
    # M[i][j] = evals[((j << log_block) + i) >> log_inv_rate]   (whir/src/utils.rs:128-150)
    n, k, r, cols = 8, 3, 1, 6
    ev = O.random_field(rng, 1 << n)
    m = O.prepare_evals(ev, n, 1, k, r, cols)
    log_block = n + r - k
    assert m.shape == (1 << log_block, cols)
    for i in range(m.shape[0]):
        for j in range(cols):
            assert m[i, j] == ev[((j << log_block) + i) >> r]


def test_rs_first_layers_are_identity_on_repeated_input(rng):
    # with log_inv_rate = r each value is repeated 2^r times, so codeword rows are the
    # evaluations of the *unrepeated* polynomial over the larger domain.
    n, k, r = 7, 2, 2
    ev = O.random_field(rng, 1 << n)
    cw = O.reorder_and_dft(ev, n, 1, k, r, 1 << k)
    h = 1 << (n + r - k)
    g = O.two_adic_generator(n + r - k)
    one = int(O.to_monty(1))
    col = 1
    chunk = ev[col << (n - k):(col + 1) << (n - k)]
    for i in (0, 1, 5, h - 1):
        y = one
        for _ in range(i):
            y = O.kb_mul(y, g)
        pt_full = O.expand_from_univariate(np.array([y, 0, 0, 0, 0], dtype=np.uint32), n + r - k)
        # variables of the repeated (low) bits do not matter: evaluate the (n-k)-variate chunk on the first n-k coords
        val = O.mle_eval(chunk, pt_full[: n - k])
        assert val[0] == cw[i, col] and not val[1:].any()


def test_mle_eval_small_cases_and_fold(rng):
    n = 6
    ev = O.random_field(rng, 1 << n)
    pt = O.random_field(rng, (n, 5))
    full = O.mle_eval(ev, pt)
    folded = O.fold_msb(ev, pt[0])
    assert np.array_equal(O.mle_eval(folded, pt[1:]), full)
    # boolean point picks an entry (x0 = MSB)
    one = int(O.to_monty(1))
    idx = 0b101100
    bpt = np.zeros((n, 5), dtype=np.uint32)
    for b in range(n):
        bpt[b, 0] = one if (idx >> (n - 1 - b)) & 1 else 0
    assert O.mle_eval(ev, bpt)[0] == ev[idx]
    eq = O.eq_table(pt)
    acc = np.zeros(5, dtype=np.uint64)
    # sum_b eq(b) * ev[b] == eval
    tot = np.zeros(5, dtype=np.uint32)
    for b in range(1 << n):
        term = O.ef_mul(eq[b], np.array([ev[b], 0, 0, 0, 0], dtype=np.uint32))
        tot = ((tot.astype(np.uint64) + term) % O.P).astype(np.uint32)
    assert np.array_equal(tot, full)


def test_next_mle_on_the_boolean_cube():
    """The reference's own test (crates/backend/poly/src/next_mle.rs:57-82, test_matrix_down_folded): on boolean inputs
    next_mle(x, y) = [y = x + 1] in big-endian binary, with next_mle(2^n - 1, 2^n - 1) = 1 — for the oracle's folded table
    (what the prover adds to the weights), the oracle's verifier formula and the product's verifier formula."""
    from oracle import whir as W

    from leanmultisig_b200 import verify as V

    n = 5
    one = int(O.to_monty(1))

    def bools(v):  # to_big_endian_in_field
        return [(1, 0, 0, 0, 0) if (v >> (n - 1 - k)) & 1 else (0, 0, 0, 0, 0) for k in range(n)]

    for x in range(1 << n):
        table = O.next_mle_folded(W._pts(bools(x)))
        for y in range(1 << n):
            expected = 1 if (x + 1 == y or (x == y == (1 << n) - 1)) else 0
            assert table[y].tolist() == [one * expected, 0, 0, 0, 0], (x, y)
            assert W._next_mle(bools(x), bools(y)) == (expected, 0, 0, 0, 0)
            assert V._next_mle(bools(x), bools(y)) == (expected, 0, 0, 0, 0)


def test_sparse_eq_statement_equals_dense_eq_with_boolean_prefix():
    """The reference's own test (crates/backend/poly/src/eq_mle.rs:1117-1136, test_compute_sparse_eval) with its values: the
    eq table of the point (0, 1, 1, 0, 96, 85, 1, 854, 2) scaled by 789 equals the SPARSE form — selector 0b0110 = 6 over
    the four boolean coordinates, eq over the remaining five — which is how statements with selectors enter the weights."""
    pt = np.zeros((9, 5), dtype=np.uint32)
    pt[:, 0] = O.to_monty(np.array([0, 1, 1, 0, 96, 85, 1, 854, 2], dtype=np.uint64))
    scalar = np.zeros(5, dtype=np.uint32)
    scalar[0] = int(O.to_monty(789))
    structured = np.zeros((1 << 9, 5), dtype=np.uint32)
    unstructured = np.zeros((1 << 9, 5), dtype=np.uint32)
    O.weights_add_eq(structured, 6, pt[4:], scalar)
    O.weights_add_eq(unstructured, 0, pt, scalar)
    assert np.array_equal(structured, unstructured)
    assert np.count_nonzero(structured[:, 0]) <= 32 and np.array_equal(O.eq_table(pt, scalar), unstructured)

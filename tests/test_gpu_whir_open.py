"""GPU tier: WHIR-open kernels (statement weights, product sumcheck rounds, STIR updates) against the oracle."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as L

    c = L.Context(0, 24)
    yield c
    c.close()


def ef_poly_eval(c0, c1, c2, r):
    return O.ef_add(O.ef_add(c0, O.ef_mul(c1, r)), O.ef_mul(c2, O.ef_mul(r, r)))


@pytest.mark.parametrize("n,dim,live_frac", [(1, 1, 1.0), (2, 1, 1.0), (3, 5, 1.0), (9, 1, 1.0), (11, 1, 0.5), (12, 5, 1.0),
                                             (14, 1, 0.37), (16, 1, 0.5)])
def test_rounds_and_folds_match_oracle(ctx, rng, n, dim, live_frac):
    """run_product_sumcheck semantics: round, fold+round (fused), plain fold; c0/c2 and tables every round."""
    N = 1 << n
    live = max(1, int(N * live_frac))
    p = O.random_field(rng, (N, 5) if dim == 5 else (N,))
    p[live:] = 0
    w0 = O.random_field(rng, (N, 5))
    sc = ctx.sumcheck(p, n, live_len=live)
    # load the weights through the public surface: one dense eq statement would not give arbitrary tables, so use
    # n unit statements?  Arbitrary tables are instead injected as 2^n single-index statements when small,
    # else as an eq table of a random point.
    if n <= 9:
        one = np.array([int(O.to_monty(1)), 0, 0, 0, 0], dtype=np.uint32)
        for i in range(N):
            sc.add_eq(i, np.zeros((0, 5), dtype=np.uint32), w0[i])
        w = w0.copy()
    else:
        pt = O.random_field(rng, (n, 5))
        scal = O.random_field(rng, 5)
        sc.add_eq(0, pt, scal)
        w = O.eq_table(pt, scal)
    cur_p, cur_w = p, w
    got_p, got_w = sc.read()
    assert np.array_equal(got_w, cur_w) and np.array_equal(got_p, cur_p)
    c0, c2 = sc.round()
    e0, e2 = O.prod_round(cur_p, cur_w)
    assert np.array_equal(c0, e0) and np.array_equal(c2, e2)
    for rnd in range(n - 1):
        r = O.random_field(rng, 5)
        cur_p, cur_w = O.fold_msb(cur_p, r), O.fold_msb(cur_w, r)
        if rnd % 2 == 0:
            c0, c2 = sc.fold_round(r)
        else:
            sc.fold(r)
            c0, c2 = sc.round()
        e0, e2 = O.prod_round(cur_p, cur_w)
        assert np.array_equal(c0, e0) and np.array_equal(c2, e2), rnd
        if rnd in (0, 1, n - 2):
            got_p, got_w = sc.read()
            assert np.array_equal(got_p, cur_p) and np.array_equal(got_w, cur_w)
    r = O.random_field(rng, 5)
    sc.fold(r)
    got_p, got_w = sc.read()
    assert np.array_equal(got_p, O.fold_msb(cur_p, r)) and np.array_equal(got_w, O.fold_msb(cur_w, r))
    sc.free()


def test_combine_statement_terms(ctx, rng):
    """combine_statement (open.rs:518-584): dense eq, sparse selectors, single-index and `next` statements."""
    n = 13
    N = 1 << n
    p = O.random_field(rng, N)
    sc = ctx.sumcheck(p, n)
    w = np.zeros((N, 5), dtype=np.uint32)
    gamma = O.random_field(rng, 5)
    g = np.array([int(O.to_monty(1)), 0, 0, 0, 0], dtype=np.uint32)
    total = np.zeros(5, dtype=np.uint32)
    terms = [("eq", 0, n), ("eq", 0, n), ("eq", 5, 10), ("eq", 2, 10), ("eq", 77, 0), ("next", 3, 11), ("next", 0, n),
             ("eq", 1, 12), ("eq", 1023, 3), ("next", 100, 6)]
    for kind, sel, m in terms:
        pt = O.random_field(rng, (m, 5))
        if kind == "eq":
            sc.add_eq(sel, pt, g)
            O.weights_add_eq(w, sel, pt, g)
        else:
            sc.add_next(sel, pt, g)
            O.weights_add_next(w, sel, pt, g)
        g = O.ef_mul(g, gamma)
    _, got = sc.read()
    assert np.array_equal(got, w)
    # the claimed sum of a statement is the MLE evaluation: check one dense eq statement end to end
    sc2 = ctx.sumcheck(p, n)
    pt = O.random_field(rng, (n, 5))
    one = np.array([int(O.to_monty(1)), 0, 0, 0, 0], dtype=np.uint32)
    sc2.add_eq(0, pt, one)
    c0, c2 = sc2.round()
    val = O.mle_eval(p, pt)
    # h(0) + h(1) = sum  with h(0) = c0, h(1) = c0 + c1 + c2; verify with the folded tables instead
    r = O.random_field(rng, 5)
    c1 = O.ef_sub(O.ef_sub(val, O.ef_add(c0, c0)), c2)
    sc2.fold(r)
    pf, wf = sc2.read()
    tot = np.zeros(5, dtype=np.uint32)
    d0, d2 = O.prod_round(pf, wf)
    # sum of the folded product = h(r)
    s_fold = O.ef_add(O.ef_add(d0, d0), O.ef_zero() if hasattr(O, "ef_zero") else np.zeros(5, dtype=np.uint32))
    hr = ef_poly_eval(c0, c1, c2, r)
    acc = np.zeros(5, dtype=np.uint32)
    for i in range(pf.shape[0]):
        acc = O.ef_add(acc, O.ef_mul(pf[i], wf[i]))
    assert np.array_equal(acc, hr)
    sc.free(), sc2.free()


@pytest.mark.parametrize("m,n_q", [(4, 3), (10, 7), (12, 33), (15, 75)])
def test_add_base_equality_batched(ctx, rng, m, n_q):
    """add_new_base_equality (open.rs:360-382) with STIR points (g^i, g^2i, ...)."""
    N = 1 << m
    p = O.random_field(rng, (N, 5))
    sc = ctx.sumcheck(p, m)
    g = O.two_adic_generator(m + 3)
    pts = np.empty((n_q, m), dtype=np.uint32)
    one = int(O.to_monty(1))
    for q in range(n_q):
        y = one
        for _ in range(int(rng.integers(0, 1 << (m + 3)))):
            y = O.kb_mul(y, g)
        for i in range(m):
            pts[q, i] = y
            y = O.kb_mul(y, y)
    scal = O.random_field(rng, (n_q, 5))
    w = np.zeros((N, 5), dtype=np.uint32)
    O.weights_add_base_eq(w, pts, scal)
    sc.add_base_eq(pts, scal)
    _, got = sc.read()
    assert np.array_equal(got, w)
    sc.free()


def test_whir_round_commit_and_ood_on_folded_polynomial(ctx, rng):
    """One WHIR round after the initial sumcheck: fold 7 variables, commit the folded EF polynomial
    (reorder_and_dft + Merkle over 32 EF columns), OOD-evaluate it (open.rs:81-99)."""
    n, k0, k1 = 14, 7, 5
    p = O.random_field(rng, 1 << n)
    tree0 = ctx.commit(p, n, k0, 1)
    sc = ctx.sumcheck_from_tree(tree0)
    pt = O.random_field(rng, (n, 5))
    sc.add_eq(0, pt, np.array([int(O.to_monty(1)), 0, 0, 0, 0], dtype=np.uint32))
    cur = p
    c0, c2 = sc.round()
    for j in range(k0):
        r = O.random_field(rng, 5)
        cur = O.fold_msb(cur, r)
        if j < k0 - 1:
            sc.fold_round(r)
        else:
            sc.fold(r)
    assert sc.n_vars == n - k0 and sc.poly_dim == 5
    # round commit: rate 1/2 -> log_inv_rate for the folded polynomial = 1 + (k0 - rs_reduction) ... use 3
    tree1 = sc.commit_poly(k1, 3)
    cw = O.reorder_and_dft(cur, n - k0, 5, k1, 3, 1 << k1)
    assert np.array_equal(tree1.codeword(), cw)
    layers = O.merkle_tree(cw, 5 << k1, 5 << k1)
    assert np.array_equal(tree1.root, layers[-1])
    z = O.random_field(rng, 5)
    ood = O.expand_from_univariate(z, n - k0)
    assert np.array_equal(sc.eval_poly(ood), O.mle_eval(cur, ood))
    rows, paths = tree1.open([3, 17])
    for q, i in enumerate([3, 17]):
        assert O.merkle_verify(tree1.root, tree1.log_height, i, rows[q], paths[q])
    tree0.free(), tree1.free(), sc.free()


@pytest.mark.gpu
@pytest.mark.parametrize("n,sel_bits,k", [(6, 0, 2), (12, 0, 10), (14, 3, 5), (16, 0, 19), (11, 11, 3),
                                          # m >= 17: the tensor-core GEMM path (csrc/sumcheck.cu weights_gemm_kernel), 12 statements per pass
                                          (17, 0, 1), (18, 1, 8), (19, 0, 3), (18, 0, 16), (19, 2, 21)])
def test_add_eq_batch_equals_separate_statements(rng, n, sel_bits, k):
    """lm_sc_add_eq_batch (all statements of one selector and length in ONE pass over the weights, delayed reduction over the
    statements) produces the weight table of adding them one by one, i.e. the oracle's"""
    import leanmultisig_b200 as lm

    if n - sel_bits < 1:
        pytest.skip("needs at least one inner variable")
    ctx = lm.Context(0, 16)
    m = n - sel_bits
    p = O.random_field(rng, 1 << n)
    w = np.zeros((1 << n, 5), dtype=np.uint32)
    sel = int(rng.integers(0, 1 << sel_bits)) if sel_bits else 0
    pts, scs = O.random_field(rng, (k, m, 5)), O.random_field(rng, (k, 5))
    for i in range(k):
        O.weights_add_eq(w, sel, pts[i], scs[i])
    sc = ctx.sumcheck(p, n)
    sc.add_eq_batch(sel, pts, scs)
    _, gw = sc.read()
    assert np.array_equal(gw, w)
    if m >= 17:  # accumulation on top of existing weights (read-modify-write of the table)
        for i in range(min(k, 2)):
            O.weights_add_eq(w, sel, pts[i], scs[i])
        sc.add_eq_batch(sel, pts[:2], scs[:2])
        _, gw = sc.read()
        assert np.array_equal(gw, w)
    sc.free()
    ctx.close()

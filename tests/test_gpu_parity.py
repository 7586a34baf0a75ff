"""GPU tier: every kernel behind the C ABI against the oracle, bit for bit, on the same seeded inputs."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

KAT_OUT = [610090613, 935319874, 1893335292, 796792199, 356405232, 552237741, 55134556, 1215104204,
           1823723405, 1133298033, 1780633798, 1453946561, 710069176, 1128629550, 1917333254, 1175481618]


@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as L

    c = L.Context(0, 24)
    yield c
    c.close()


def test_poseidon1_kat_and_random(ctx, rng):
    x = O.to_monty(np.arange(16))
    assert O.from_monty(ctx.poseidon1(x)).tolist() == KAT_OUT
    s = O.random_field(rng, (5000, 16))
    s[0] = O.P - 1
    s[1] = 0
    assert np.array_equal(ctx.poseidon1(s), O.poseidon1_permute(s))
    assert np.array_equal(ctx.poseidon1(s, compress=True), O.poseidon1_compress(s))


def test_scalar_poseidon_kernels_subprocess():
    """LM_P1_SCALAR=1 selects the one-state-per-thread kernels (poseidon1.cuh) instead of the tensor-core formulation: same
    permutation, sponge and tree, bit for bit (the switch is read once per process, hence the subprocess)."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy as np, oracle as O, leanmultisig_b200 as L\n"
        "rng = np.random.default_rng(7)\n"
        "c = L.Context(0, 20)\n"
        "s = O.random_field(rng, (3000, 16))\n"
        "assert np.array_equal(c.poseidon1(s), O.poseidon1_permute(s))\n"
        "assert np.array_equal(c.poseidon1(s, compress=True), O.poseidon1_compress(s))\n"
        "for log_h, stored, full, eff in [(14, 64, 128, 64), (9, 64, 128, 57), (15, 24, 24, 24)]:\n"
        "    m = O.random_field(rng, (1 << log_h, stored)); m[:, eff:] = 0\n"
        "    assert np.array_equal(c.merkle_tree(m, full, eff), O.merkle_tree(m, full, eff))\n"
        "c.close()\n"
        "print('scalar path ok')\n"
    )
    env = dict(os.environ, LM_P1_SCALAR="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "scalar path ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("log_h,stored,full,eff", [
    (0, 16, 16, 16), (1, 16, 16, 16), (3, 64, 128, 64), (5, 128, 128, 128), (9, 64, 128, 57), (10, 160, 160, 160),
    (7, 90, 160, 85), (4, 16, 64, 8), (4, 16, 64, 1), (12, 64, 128, 64), (6, 20, 160, 20), (8, 6, 16, 6), (8, 24, 24, 24),
    (15, 16, 16, 16), (14, 90, 160, 85), (16, 8, 64, 8),
])
def test_merkle_tree_layers(ctx, rng, log_h, stored, full, eff):
    h = 1 << log_h
    mat = O.random_field(rng, (h, stored))
    mat[:, eff:] = 0
    got = ctx.merkle_tree(mat, full, eff)
    exp = O.merkle_tree(mat, full, eff)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("log_h,w", [(1, 8), (2, 4), (3, 8), (5, 16), (8, 64), (11, 8), (12, 24), (13, 12), (14, 160), (10, 5), (6, 3), (17, 8)])
def test_dft_batch_by_evals(ctx, rng, log_h, w):
    mat = O.random_field(rng, (1 << log_h, w))
    assert np.array_equal(ctx.dft_batch_by_evals(mat), O.dft_batch_by_evals(mat))


@pytest.mark.parametrize("n_vars,dim,k,r,cols", [
    (10, 1, 3, 1, 8), (12, 1, 4, 1, 8), (14, 1, 7, 1, 64), (14, 1, 7, 2, 128), (16, 1, 7, 1, 60), (9, 5, 5, 1, 32),
    (12, 5, 5, 2, 32), (13, 5, 5, 3, 12), (8, 1, 2, 0, 4), (15, 1, 7, 4, 128), (18, 1, 7, 1, 64),
])
def test_reorder_and_dft(ctx, rng, n_vars, dim, k, r, cols):
    shape = (1 << n_vars, 5) if dim == 5 else (1 << n_vars,)
    ev = O.random_field(rng, shape)
    got = ctx.reorder_and_dft(ev, n_vars, k, r, cols)
    exp = O.reorder_and_dft(ev, n_vars, dim, k, r, cols)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("n_vars,dim", [(0, 1), (1, 1), (3, 5), (9, 1), (10, 5), (11, 1), (14, 1), (16, 5), (20, 1)])
def test_mle_eval_and_fold(ctx, rng, n_vars, dim):
    shape = (1 << n_vars, 5) if dim == 5 else (1 << n_vars,)
    ev = O.random_field(rng, shape)
    pt = O.random_field(rng, (n_vars, 5))
    assert np.array_equal(ctx.mle_eval(ev, pt), O.mle_eval(ev, pt))
    if n_vars >= 1:
        assert np.array_equal(ctx.fold_msb(ev, pt[0]), O.fold_msb(ev, pt[0]))
        assert np.array_equal(ctx.eq_table(pt[: min(n_vars, 12)]), O.eq_table(pt[: min(n_vars, 12)]))


def test_mle_eval_live_prefix(ctx, rng):
    n = 15
    ev = O.random_field(rng, 1 << n)
    live = 12345
    ev[live:] = 0
    pt = O.random_field(rng, (n, 5))
    assert np.array_equal(ctx.mle_eval(ev, pt, live_len=live), O.mle_eval(ev, pt))


@pytest.mark.parametrize("n_vars,dim,k,r,live_frac", [
    (14, 1, 7, 1, 0.5), (14, 1, 7, 1, 1.0), (16, 1, 7, 2, 0.37), (12, 5, 5, 1, 1.0), (13, 5, 5, 2, 1.0), (10, 1, 4, 3, 0.9),
    (18, 1, 7, 1, 0.5),
])
def test_commit_open_eval(ctx, rng, n_vars, dim, k, r, live_frac):
    """WhirConfig::commit seam: codeword, every digest layer, root, openings, OOD evaluation."""
    n = 1 << n_vars
    live = max(1, int(n * live_frac))
    shape = (n, 5) if dim == 5 else (n,)
    ev = O.random_field(rng, shape)
    ev[live:] = 0
    tree = ctx.commit(ev, n_vars, k, r, actual_len=live)
    n_blocks = 1 << k
    eff_cols = -(-live // (n >> k))
    assert tree.height == 1 << (n_vars + r - k) and tree.full_width == n_blocks * dim
    stored_cols = tree.stored_width // dim
    assert stored_cols >= eff_cols
    cw_exp = O.reorder_and_dft(ev, n_vars, dim, k, r, stored_cols)
    assert np.array_equal(tree.codeword(), cw_exp)
    layers_exp = O.merkle_tree(cw_exp, n_blocks * dim, eff_cols * dim)
    assert np.array_equal(tree.layers(), layers_exp)
    assert np.array_equal(tree.root, layers_exp[-1])
    # the stored width does not change the commitment (commit.rs:70-74): hash the full-width matrix too
    if stored_cols < n_blocks and n_vars <= 14:
        cw_full = O.reorder_and_dft(ev, n_vars, dim, k, r, n_blocks)
        assert np.array_equal(O.merkle_tree(cw_full, n_blocks * dim, n_blocks * dim)[-1], tree.root)
    idx = rng.integers(0, tree.height, size=17).tolist() + [0, tree.height - 1]
    rows, paths = tree.open(idx)
    for q, i in enumerate(idx):
        er, ep = O.merkle_open(cw_exp, n_blocks * dim, layers_exp, i)
        assert np.array_equal(rows[q], er) and np.array_equal(paths[q], ep)
        assert O.merkle_verify(tree.root, tree.log_height, i, rows[q], paths[q])
    # OOD sample at (z, z^2, z^4, ...) (commit.rs:89-92)
    z = O.random_field(rng, 5)
    pt = O.expand_from_univariate(z, n_vars)
    assert np.array_equal(tree.evaluate(pt), O.mle_eval(ev, pt))
    tree.free()


def test_commit_dev_matches_commit(ctx, rng):
    n_vars, k, r = 15, 7, 1
    ev = O.random_field(rng, 1 << n_vars)
    t1 = ctx.commit(ev, n_vars, k, r)
    d = ctx.to_device(ev)
    t2 = ctx.commit_dev(d, n_vars, 1, k, r, 1 << n_vars)
    assert np.array_equal(t1.root, t2.root)
    t1.free(), t2.free(), d.free()


def test_error_behaviour(ctx):
    import leanmultisig_b200 as L

    ev = np.zeros(1 << 8, dtype=np.uint32)
    with pytest.raises(L.LmError):
        ctx.commit(ev, 8, 9, 1)  # folding > n_vars
    with pytest.raises(L.LmError):
        ctx.commit(ev, 30, 2, 1, actual_len=16)  # domain beyond the twiddle table
    t = ctx.commit(ev, 8, 4, 1)
    with pytest.raises(L.LmError):
        t.open([t.height])
    t.free()


def test_rs_linearity_at_full_size(ctx, rng):
    """Size-independent property at the BASELINE shape (2^22 x 64): the encoder is F-linear per column,
    so commit(a) + commit(b) == commit(a + b) on the codeword; roots are checked through a checksum path:
    every opened row verifies against the root."""
    n_vars, k, r = 28, 7, 1
    live = 1 << 27
    rs = np.random.default_rng(7)
    a = rs.integers(0, O.P, size=live, dtype=np.uint32)
    tree = ctx.commit(a, n_vars, k, r, actual_len=live)
    assert tree.height == 1 << 22 and tree.stored_width == 64 and tree.full_width == 128
    idx = rs.integers(0, tree.height, size=64)
    rows, paths = tree.open(idx)
    for q, i in enumerate(idx.tolist()):
        assert O.merkle_verify(tree.root, 22, i, rows[q], paths[q])
        # row i, column j is the (n-k)-variate chunk j evaluated at (w^i, w^2i, ...): check 2 columns
    g = O.two_adic_generator(22)
    one = int(O.to_monty(1))
    i = int(idx[0])
    y = one
    base, e = g, i
    while e:
        if e & 1:
            y = O.kb_mul(y, base)
        base = O.kb_mul(base, base)
        e >>= 1
    pt = O.expand_from_univariate(np.array([y, 0, 0, 0, 0], dtype=np.uint32), 22)[:21]
    for col in (0, 63):
        chunk = a[col << 21:(col + 1) << 21]
        val = O.mle_eval(chunk, pt)
        assert val[0] == rows[0][col] and not val[1:].any()
    tree.free()


@pytest.mark.parametrize("log_h,g,w", [(10, 1, 8), (11, 2, 16), (12, 3, 64), (9, 3, 4), (12, 4, 8)])
def test_dft_layers_mapped(ctx, rng, log_h, g, w):
    """last g butterfly layers on the rows one rank holds after the exchange of the row-sharded commit
    (lm_dev_dft_layers_mapped: fused radix-2^g kernel for g <= 3, one launch per layer otherwise) against the oracle,
    for every rank's row set"""
    from leanmultisig_b200._lib import check, lib

    G = 1 << g
    h = 1 << log_h
    block, run = h // G, h // (G * G)
    for rank in (0, G - 1, G // 2):
        mat = O.random_field(rng, (G * run, w))
        exp = O.dft_layers_mapped(mat.copy(), log_h, log_h - g, G, run, block, rank * run)
        d = ctx.to_device(mat)
        check(lib().lm_dev_dft_layers_mapped(ctx.handle, d.ptr, w, log_h, log_h - g, G, run, block, rank * run))
        got = d.download(mat.shape)
        d.free()
        assert np.array_equal(got, exp), (log_h, g, w, rank)
        # out-of-place form (what the sharded commit uses to leave the exchange buffer): source untouched, result in `out`;
        # a column range leaves the other columns of `out` alone
        d, out = ctx.to_device(mat), ctx.to_device(np.zeros_like(mat))
        check(lib().lm_dev_dft_layers_mapped_out(ctx.handle, d.ptr, out.ptr, w, log_h, log_h - g, G, run, block, rank * run, 0, 0))
        assert np.array_equal(out.download(mat.shape), exp) and np.array_equal(d.download(mat.shape), mat)
        if g <= 3 and w >= 8:
            out2 = ctx.to_device(np.zeros_like(mat))
            check(lib().lm_dev_dft_layers_mapped_out(ctx.handle, d.ptr, out2.ptr, w, log_h, log_h - g, G, run, block, rank * run, 4, w - 4))
            part = out2.download(mat.shape)
            assert np.array_equal(part[:, 4:], exp[:, 4:]) and not part[:, :4].any()
            out2.free()
        d.free()
        out.free()


@pytest.mark.parametrize("n_vars,k,r,cols,G", [(13, 4, 1, 16, 2), (15, 5, 1, 32, 4), (19, 7, 1, 64, 2), (20, 7, 1, 64, 8)])
def test_reorder_and_dft_scatter_single_device(ctx, rng, n_vars, k, r, cols, G):
    """lm_dev_reorder_and_dft_scatter with all "peer" matrices on this device: the transform of every rank's shard stores
    row i into matrix i >> log_run at local row (rank << log_run) | (i mod run) — what the NCCL all-to-all of the
    row-sharded commit would deliver (one and two passes)."""
    import ctypes as C

    from leanmultisig_b200._lib import check, lib
    from leanmultisig_b200.sharded import shard_of

    g = G.bit_length() - 1
    chunk = 1 << (n_vars - k)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: cols * chunk] = O.random_field(rng, cols * chunk)
    block = 1 << (n_vars - g + r - k)
    run = block // G
    mats = [ctx.alloc(block * cols * 4) for _ in range(G)]
    work = ctx.alloc(block * cols * 4)
    table = np.array([int(m.ptr.value) for m in mats], dtype=np.uint64)
    expected = [np.zeros((block, cols), dtype=np.uint32) for _ in range(G)]
    for rank in range(G):
        shard = shard_of(ev, n_vars, k, rank, G).reshape(1 << k, -1)[:cols].reshape(-1)
        t_local = O.reorder_and_dft(np.concatenate([shard, np.zeros((1 << (n_vars - g)) - shard.size, dtype=np.uint32)]),
                                    n_vars - g, 1, k, r, cols)
        for q in range(G):
            expected[q][rank * run:(rank + 1) * run] = t_local[q * run:(q + 1) * run]
        d = ctx.to_device(shard)
        check(lib().lm_dev_reorder_and_dft_scatter(ctx.handle, d.ptr, n_vars - g, k, r, cols, work.ptr,
                                                   table.ctypes.data_as(C.POINTER(C.c_uint64)), G, rank))
        ctx.sync()
        d.free()
    for q in range(G):
        assert np.array_equal(mats[q].download((block, cols)), expected[q]), f"matrix {q}"
    for m in mats + [work]:
        m.free()


@pytest.mark.parametrize("G,cols", [(2, 64), (4, 32), (8, 64)])
def test_sharded_pipeline_pieces_by_column_groups(ctx, rng, G, cols):
    """The column-range forms used by the pipelined host-input sharded commit, on one device: transform + scatter, last
    layers and leaf absorb applied group by group (1, 1, 2, 4 .. chunks, right to left) give the same matrices and leaf
    digests as the whole-width calls / the oracle."""
    import ctypes as C

    from leanmultisig_b200._lib import check, lib
    from leanmultisig_b200.sharded import shard_of

    n_vars, k, r = 17, 7, 1
    g = G.bit_length() - 1
    chunk = 1 << (n_vars - k)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: cols * chunk] = O.random_field(rng, cols * chunk)
    log_h = n_vars + r - k
    block, run = (1 << log_h) // G, (1 << log_h) // (G * G)
    mats = [ctx.alloc(block * cols * 4) for _ in range(G)]
    work = ctx.alloc(block * cols * 4)
    table = np.array([int(m.ptr.value) for m in mats], dtype=np.uint64)
    tptr = table.ctypes.data_as(C.POINTER(C.c_uint64))
    groups, chunk_end, take, first = [], cols // 8, 1, True
    while chunk_end > 0:
        take = min(take, chunk_end)
        groups.append((chunk_end, take))
        chunk_end -= take
        take, first = (take if first else take * 2), False
    shards = [ctx.to_device(shard_of(ev, n_vars, k, q, G).reshape(1 << k, -1)[:cols].reshape(-1)) for q in range(G)]
    digests = [ctx.alloc(block * 8 * 4) for _ in range(G)]
    for chunk_end, take in groups:
        cb, cnt = (chunk_end - take) * 8, take * 8
        for q in range(G):   # every "rank" scatters this group, then every rank finishes it
            check(lib().lm_dev_reorder_and_dft_scatter_cols(ctx.handle, shards[q].ptr, n_vars - g, k, r, cols, work.ptr, tptr, G, q, cb, cnt))
        for q in range(G):
            check(lib().lm_dev_dft_layers_mapped_cols(ctx.handle, mats[q].ptr, cols, log_h, log_h - g, G, run, block, q * run, cb, cnt))
            check(lib().lm_dev_merkle_absorb(ctx.handle, mats[q].ptr, block, cols, 128, cols, chunk_end - 1, take, digests[q].ptr))
    ctx.sync()
    cw = O.reorder_and_dft(ev, n_vars, 1, k, r, cols)
    for q in range(G):
        exp = np.concatenate([cw[m * block + q * run: m * block + (q + 1) * run] for m in range(G)])
        assert np.array_equal(mats[q].download((block, cols)), exp), f"matrix {q}"
        assert np.array_equal(digests[q].download((block, 8)), O.first_digest_layer(exp, 128, cols)), f"digests {q}"
    for b in mats + digests + shards + [work]:
        b.free()

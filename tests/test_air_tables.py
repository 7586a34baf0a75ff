"""AIR bodies of the extension_op and poseidon16 tables and the Poseidon16 trace generator.

CPU tier pins the oracle: on traces built by the tables' own execution semantics (extension_op/exec.rs:96-190,
poseidon_16/trace_gen.rs) every constraint vanishes, and the Poseidon16 outputs are those of the KAT-pinned
permutation.  GPU tier compares the device sessions with the oracle round by round, and replays the reference's
sub_protocols/tests/prove_poseidon_16.rs flow (commit -> AIR sumcheck -> WHIR open -> verify) with the GPU as prover.
"""
import numpy as np
import pytest

import oracle as O
from oracle import whir as W
from leanmultisig_b200 import field as F
P = 0x7F000001

ONE_M = int(O.to_monty(1))


def m(x):
    return int(O.to_monty(x))


# ------------------------------------------------------------------------------------------------ trace builders
def make_poseidon16_trace(rng, log_n, modes=True):
    n = 1 << log_n
    cols = np.zeros((109, n), dtype=np.uint32)
    cols[9:25] = O.random_field(rng, (16, n))
    cols[0] = ONE_M
    for r in range(n):
        mode = r % 4 if modes else 0
        if mode == 1:
            cols[3, r] = ONE_M  # half output
        if mode == 2:
            cols[8, r] = ONE_M  # permute
        if mode == 3:  # hardcoded left: offset = effective_index_left_first
            cols[4, r], cols[5, r], cols[6, r], cols[7, r] = ONE_M, m(24), m(24), m(100 + r)
        else:
            cols[6, r], cols[7, r] = m(8 * r), m(8 * r + 4)
        cols[1, r], cols[2, r] = m(3 * r + 1), m(5 * r + 2)
    return cols


def make_ext_op_trace(rng, log_n, n_active):
    """rows as pushed by exec_multi_row (extension_op/exec.rs:96-190) followed by padding rows (mod.rs:123-132)"""
    n = 1 << log_n
    cols = np.zeros((29, n), dtype=np.uint32)
    row = 0
    rnd_ef = lambda: tuple(int(x) for x in rng.integers(0, F.P, 5))
    while row < n_active:
        size = min(int(rng.integers(1, 5)), n_active - row)
        is_be = int(rng.integers(0, 2))
        op = int(rng.integers(0, 3))  # 0 add, 1 mul, 2 poly_eq
        ptr_a, ptr_b, ptr_res = (int(rng.integers(1, 1 << 20)) for _ in range(3))
        vas, vbs, elems = [], [], []
        for i in range(size):
            mem_a = rnd_ef()  # five consecutive memory words at idx_a
            va = (mem_a[0], 0, 0, 0, 0) if is_be else mem_a
            vb = rnd_ef()
            if op == 0:
                e = F.add(va, vb)
            elif op == 1:
                e = F.mul(va, vb)
            else:
                ab = F.mul(va, vb)
                e = F.add(F.sub(F.sub(F.add(ab, ab), va), vb), F.ONE)
            vas.append(mem_a), vbs.append(vb), elems.append(e)
        comps = [None] * size
        comps[-1] = elems[-1]
        for i in range(size - 2, -1, -1):
            comps[i] = F.mul(elems[i], comps[i + 1]) if op == 2 else F.add(elems[i], comps[i + 1])
        for i in range(size):
            r = row + i
            cols[0, r] = m(is_be)
            cols[1, r] = m(1 if i == 0 else 0)
            cols[2, r] = m(size - i)
            cols[3 + op, r] = ONE_M
            cols[6, r] = m(ptr_a + i * (1 if is_be else 5))
            cols[7, r] = m(ptr_b + 5 * i)
            cols[8:13, r] = F.to_monty(comps[i])
            cols[13, r] = m(ptr_res)
            cols[14:19, r] = F.to_monty(vas[i])
            cols[19:24, r] = F.to_monty(vbs[i])
            cols[24:29, r] = F.to_monty(comps[0])
        row += size
    cols[1, row:] = ONE_M  # padding: start = 1, len = 1, everything else 0 (indexes: zero_vec_ptr = 0 here)
    cols[2, row:] = ONE_M
    return cols


def with_shifts(table, cols):
    _, n_shift, _ = O.air_shape(table)
    if n_shift == 0:
        return cols
    return np.concatenate([cols, np.stack([O.shift_column(cols[c]) for c in range(n_shift)])])


def extras(rng, n_alpha=101):
    alpha = O.random_field(rng, 5)
    ap = [np.array([ONE_M, 0, 0, 0, 0], dtype=np.uint32)]
    for _ in range(n_alpha - 1):
        ap.append(O.ef_mul(ap[-1], alpha))
    return np.stack(ap), O.random_field(rng, (6, 5)), O.random_field(rng, 5)


def bus_value(la, beta, flag, data):
    s = F.ZERO
    for i, d in enumerate(data):
        s = F.add(s, F.scal(F.from_monty(la[i]), d))
    s = F.add(s, F.from_monty(la[-1]))
    return F.add(F.mul(s, F.from_monty(beta)), (flag, 0, 0, 0, 0))


# ------------------------------------------------------------------------------------------------ CPU tier
def test_poseidon16_trace_outputs_are_the_permutation(rng):
    tr = O.poseidon16_fill_trace(make_poseidon16_trace(rng, 6))
    st = tr[9:25].T.copy()
    perm, comp = O.poseidon1_permute(st), O.poseidon1_compress(st)
    for r in range(64):
        if r % 4 == 2:
            assert np.array_equal(tr[93:101, r], perm[r, :8]) and np.array_equal(tr[101:109, r], perm[r, 8:])
        else:
            assert np.array_equal(tr[93:101, r], comp[r, :8]) and not tr[101:109, r].any()


def test_poseidon16_air_vanishes_on_valid_rows(rng):
    tr = O.poseidon16_fill_trace(make_poseidon16_trace(rng, 5))
    ap, la, beta = extras(rng)
    for r in range(32):
        pt = O.embed(tr[:, r])
        assert not O.air_eval(O.AIR_POSEIDON16 | O.AIR_NO_BUS, pt, ap, la, beta).any(), r
        # BUS = true: only the bus term survives, at alpha^0
        flag_hard, off = (1, 24) if r % 4 == 3 else (0, 0)
        pdata = 1 + 4 * (r % 4 == 1) + 8 * flag_hard + 16 * flag_hard * off + 2 * (r % 4 == 2)
        index_a = (100 + r) if r % 4 == 3 else 8 * r
        exp = F.mul(F.from_monty(ap[0]), bus_value(la, beta, 1, [pdata, index_a, 3 * r + 1, 5 * r + 2]))
        assert F.from_monty(O.air_eval(O.AIR_POSEIDON16, pt, ap, la, beta)) == exp
    for col in (0, 3, 12, 30, 60, 80, 95, 105):  # every region of the row is constrained
        pt = O.embed(tr[:, 2])
        pt[col, 0] = (int(pt[col, 0]) + 1) % O.P
        assert O.air_eval(O.AIR_POSEIDON16 | O.AIR_NO_BUS, pt, ap, la, beta).any(), col


def test_ext_op_air_vanishes_on_valid_rows(rng):
    cols = make_ext_op_trace(rng, 6, 50)
    full = with_shifts(O.AIR_EXT_OP, cols)
    assert full.shape[0] == 42
    ap, la, beta = extras(rng)
    seen = set()
    for r in range(64):
        pt = O.embed(full[:, r])
        assert not O.air_eval(O.AIR_EXT_OP | O.AIR_NO_BUS, pt, ap, la, beta).any(), r
        c = O.from_monty(cols[:, r])
        active = int(c[1]) * int(c[3] + c[4] + c[5])
        aux = 4 * int(c[0]) + 8 * int(c[3]) + 16 * int(c[4]) + 32 * int(c[5]) + 64 * int(c[2])
        exp = F.mul(F.from_monty(ap[0]), bus_value(la, beta, active, [aux, int(c[6]), int(c[7]), int(c[13])]))
        assert F.from_monty(O.air_eval(O.AIR_EXT_OP, pt, ap, la, beta)) == exp
        seen.add((int(c[0]), int(c[3]), int(c[4]), int(c[5])))
    assert len(seen) >= 6  # add / mul / poly_eq x base / extension first operand (+ padding)
    canon = O.from_monty(cols)
    # a first row of a multi-row extension-field operation: every column of it (and of its shift) is constrained
    r0 = next(r for r in range(63) if canon[0, r] == 0 and canon[1, r] == 1 and canon[2, r] > 1)
    for col in (0, 1, 2, 6, 9, 15, 20, 26, 31, 38):
        pt = O.embed(full[:, r0])
        pt[col, 0] = (int(pt[col, 0]) + 1) % O.P
        assert O.air_eval(O.AIR_EXT_OP | O.AIR_NO_BUS, pt, ap, la, beta).any(), col


@pytest.mark.parametrize("table", [O.AIR_EXT_OP, O.AIR_POSEIDON16 | O.AIR_NO_BUS])
def test_oracle_round_is_consistent_with_pointwise_evaluation(rng, table):
    """sum_j eq(j) C(row pair j at z) from lm_or_air_round equals the same sum built from lm_or_air_eval"""
    L = 3
    base = make_ext_op_trace(rng, L, 5) if (table & 0xFF) == 1 else O.poseidon16_fill_trace(make_poseidon16_trace(rng, L))
    cols = with_shifts(table, base)
    cols = O.random_field(rng, cols.shape)  # arbitrary values: the round polynomial is defined for any table
    ap, la, beta = extras(rng)
    eqp = O.random_field(rng, (L - 1, 5))
    got = O.air_round(table, cols, eqp, ap, la, beta)
    eq = O.eq_table(eqp)
    deg = O.air_shape(table)[2]
    for zi in range(deg):
        z = 0 if zi == 0 else zi + 1
        tot = F.ZERO
        for j in range(1 << (L - 1)):
            lo, hi = O.from_monty(cols[:, 2 * j]).astype(np.int64), O.from_monty(cols[:, 2 * j + 1]).astype(np.int64)
            pt = O.embed(O.to_monty((lo + z * (hi - lo)) % O.P))
            tot = F.add(tot, F.mul(F.from_monty(eq[j]), F.from_monty(O.air_eval(table, pt, ap, la, beta))))
        assert F.from_monty(got[zi]) == tot


# ------------------------------------------------------------------------------------------------ GPU tier
@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as lm

    c = lm.Context(0, 22)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [0, 3, 7, 12])
def test_gpu_fill_trace_poseidon16(ctx, rng, log_n):
    import leanmultisig_b200 as lm

    cols = make_poseidon16_trace(rng, log_n)
    exp = O.poseidon16_fill_trace(cols)
    trace = [cols[c].copy() for c in range(109)]
    lm.fill_trace_poseidon_16(ctx, trace)
    assert np.array_equal(np.stack(trace), exp)


def oracle_rounds(table, cols, eq_factor, ap, la, beta, challenges):
    cur, raws = cols, []
    L = eq_factor.shape[0]
    for r in range(L):
        raws.append(O.air_round(table, cur, eq_factor[: L - r - 1], ap, la, beta))
        cur = np.stack([O.fold_lsb(cur[c], challenges[r]) for c in range(cur.shape[0])])
    return raws, cur[:, 0, :]


@pytest.mark.gpu
@pytest.mark.parametrize("table,L,valid", [(1, 1, True), (1, 5, True), (1, 9, False), (0x101, 6, False),
                                           (2, 1, True), (2, 4, True), (2, 8, False), (0x102, 7, True)])
def test_gpu_air_session_matches_oracle(ctx, rng, table, L, valid):
    import leanmultisig_b200 as lm

    if (table & 0xFF) == 1:
        base = make_ext_op_trace(rng, L, max(1, (1 << L) - 3))
    else:
        base = O.poseidon16_fill_trace(make_poseidon16_trace(rng, L))
    if not valid:
        base = O.random_field(rng, base.shape)
    cols = with_shifts(table, base)
    n_cols, n_shift, deg = O.air_shape(table)
    eq_factor = O.random_field(rng, (L, 5))
    ap, la, beta = extras(rng)
    challenges = O.random_field(rng, (L, 5))
    sum0 = np.zeros(5, dtype=np.uint32)
    sess = lm.AirSumcheckSession(ctx, table, list(base), eq_factor, sum0, ap, la, beta)
    assert sess.initial_n_vars() == L and sess.bare_degree() == deg
    raws, finals = oracle_rounds(table, cols, eq_factor, ap, la, beta, challenges)
    s, mmf = F.ZERO, F.ONE
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
    assert np.array_equal(sess.final_column_evals(), finals)
    sess.free()


@pytest.mark.gpu
@pytest.mark.parametrize("log_n_rows", [10, 11])
def test_gpu_prove_poseidon_16_end_to_end(ctx, rng, log_n_rows):
    """sub_protocols/tests/prove_poseidon_16.rs:27-141 with the GPU as prover and the oracle as verifier"""
    import leanmultisig_b200 as lm
    from leanmultisig_b200.whir import _points_to_monty

    n_rows, n_cols = 1 << log_n_rows, 109
    table = O.AIR_POSEIDON16 | O.AIR_NO_BUS
    trace_in = np.zeros((n_cols, n_rows), dtype=np.uint32)
    trace_in[9:25] = O.random_field(rng, (16, n_rows))
    trace_in[0] = ONE_M
    trace_in[7] = m(4)
    trace = [trace_in[c].copy() for c in range(n_cols)]
    lm.fill_trace_poseidon_16(ctx, trace)
    kw = dict(security_level=124, pow_bits=16, first_folding=7, subsequent_folding=4, rs_domain_initial_reduction_factor=5,
              max_num_variables_to_send_coeffs=9, starting_log_inv_rate=1)
    packed_n_vars = (n_cols * n_rows - 1).bit_length()
    cfg = lm.WhirConfig(packed_n_vars, **kw)

    # ---- prover (GPU) ----
    ps = lm.ProverState(ctx)
    poly = np.zeros(1 << packed_n_vars, dtype=np.uint32)
    poly[: n_cols * n_rows] = np.concatenate(trace)
    prover = lm.WhirProver(ctx, cfg)
    witness = prover.commit(ps, poly, n_cols * n_rows)
    alpha = ps.sample()
    ap = [np.array([ONE_M, 0, 0, 0, 0], dtype=np.uint32)]
    for _ in range(99):
        ap.append(F.to_monty(F.mul(F.from_monty(ap[-1]), F.from_monty(alpha))))
    ap = np.stack(ap)
    ps.duplex()
    eq_factor = np.stack(ps.sample_vec(log_n_rows))
    zero = np.zeros(5, dtype=np.uint32)
    sess = lm.AirSumcheckSession(ctx, table, trace, eq_factor, zero, ap, np.zeros((0, 5), dtype=np.uint32), zero)

    def absorb_and_sample(coeffs):
        ps.add_sumcheck_polynomial(coeffs)
        return ps.sample()

    point = lm.prove_batched_air_sumcheck([sess], np.array([ONE_M, 0, 0, 0, 0], dtype=np.uint32), absorb_and_sample)
    col_evals = sess.final_column_evals()
    sess.free()
    ps.add_extension_scalars(col_evals.reshape(-1))
    natural = [F.from_monty(x) for x in point[::-1]]
    log_cols = (n_cols - 1).bit_length()
    betas = [F.from_monty(x) for x in ps.sample_vec(log_cols)]
    padded = [F.from_monty(v) for v in col_evals] + [F.ZERO] * ((1 << log_cols) - n_cols)
    packed_eval = W.mle_eval_small(padded, betas)
    stm = lm.SparseStatement.dense(_points_to_monty(betas + natural), F.to_monty(packed_eval))
    prover.prove(ps, [stm], witness)
    witness.free()

    # ---- verifier (oracle) ----
    cfg_o = W.WhirConfig(packed_n_vars, **kw)
    vs = W.VerifierState(ps.transcript, ps.merkle_paths)
    pc = W.parse_commitment(cfg_o, vs)
    alpha_v = vs.sample()
    assert np.array_equal(alpha_v, alpha)
    vs.duplex()
    eq_v = [W.fm(x) for x in vs.sample_vec(log_n_rows)]
    chals, claimed = W.sumcheck_verify(vs, log_n_rows, 10 + 1, W.ZERO)
    col_evals_v = vs.next_extension_scalars_vec(n_cols)
    constraint_eval = W.fm(O.air_eval(table, col_evals_v, ap, np.zeros((1, 5), dtype=np.uint32), zero))
    natural_v = chals[::-1]
    assert W.mul(W.eq_outside(eq_v, natural_v), constraint_eval) == claimed
    betas_v = [W.fm(x) for x in vs.sample_vec(log_cols)]
    padded_v = [W.fm(v) for v in col_evals_v] + [W.ZERO] * ((1 << log_cols) - n_cols)
    stm_v = W.SparseStatement.dense(betas_v + natural_v, W.mle_eval_small(padded_v, betas_v))
    W.verify(cfg_o, vs, pc, [stm_v])


@pytest.mark.gpu
@pytest.mark.parametrize("table,L,G", [(0, 6, 2), (1, 5, 4), (2, 4, 2)])
def test_gpu_air_shard_sessions_add_up_to_the_whole_table(ctx, rng, table, L, G):
    """lm_air_new_shard / lm_air_new_folded on ONE device: G row-range shard sessions (halo row + prefix eq scale), their round
    sums added mod p, equal the oracle's rounds over the whole table; the gathered column values continue in a folded
    session (what leanmultisig_b200/sharded.py does across ranks)."""
    import leanmultisig_b200 as lm

    n_cols, n_shift, deg = O.air_shape(table)
    g = G.bit_length() - 1
    base = O.random_field(rng, (n_cols, 1 << L))
    cols = with_shifts(table, base)
    eq_factor = O.random_field(rng, (L, 5))
    ap, la, beta = extras(rng)
    challenges = O.random_field(rng, (L, 5))
    raws, finals = oracle_rounds(table, cols, eq_factor, ap, la, beta, challenges)
    per = (1 << L) // G
    zero = np.zeros(5, dtype=np.uint32)
    shards = []
    for q in range(G):
        scale = F.ONE
        for k in range(g):
            e = F.from_monty(eq_factor[k])
            scale = F.mul(scale, e if (q >> (g - 1 - k)) & 1 else F.sub(F.ONE, e))
        halo = base[:n_shift, (q + 1) * per] if q + 1 < G and n_shift else None
        shards.append(lm.AirSumcheckSession(ctx, table, [base[c, q * per:(q + 1) * per] for c in range(n_cols)], eq_factor[g:],
                                            zero, ap, la, beta, halo_next_row=halo, eq_scale=F.to_monty(scale)))
    for r in range(L - g):
        total = sum(s._raw_round().astype(np.int64) for s in shards) % P
        assert np.array_equal(total.astype(np.uint32), raws[r]), f"round {r}"
        for s in shards:
            s._fold(challenges[r])
    table_g = np.stack([s.final_column_evals() for s in shards]).transpose(1, 0, 2)
    tail = lm.AirSumcheckSession(ctx, table, None, eq_factor[:g], zero, ap, la, beta, folded_columns=table_g)
    for r in range(L - g, L):
        assert np.array_equal(tail._raw_round(), raws[r]), f"round {r}"
        tail._fold(challenges[r])
    assert np.array_equal(tail.final_column_evals(), finals)
    for s in shards + [tail]:
        s.free()


@pytest.mark.gpu
def test_gpu_batched_air_sumcheck_native_spine_matches_python_spine(ctx, rng):
    """lm_air_prove_batched (C++ spine) against prove_batched_air_sumcheck driven from Python: three sessions of different
    heights and degrees (execution 2^9, extension_op 2^7, poseidon16 2^6) joined back-loaded; same transcript, same
    challenges, same final column evaluations."""
    import leanmultisig_b200 as lm

    shapes = [(0, 9), (1, 7), (2, 6)]
    ap, la, beta = extras(rng)
    data = []
    for table, L in shapes:
        n_cols, _, _ = O.air_shape(table)
        data.append((table, O.random_field(rng, (n_cols, 1 << L)), O.random_field(rng, (L, 5)), O.random_field(rng, 5)))
    eta = O.random_field(rng, 5)

    def sessions():
        return [lm.AirSumcheckSession(ctx, t, list(cols), eqf, s, ap, la, beta) for t, cols, eqf, s in data]

    ps_py = lm.ProverState(ctx)

    def absorb_and_sample(coeffs):
        ps_py.add_sumcheck_polynomial(coeffs)
        return ps_py.sample()

    s_py = sessions()
    ch_py = lm.prove_batched_air_sumcheck(s_py, eta, absorb_and_sample)
    finals_py = [s.final_column_evals() for s in s_py]
    ps_n = lm.NativeProverState(ctx)
    s_n = sessions()
    ch_n = lm.prove_batched_air_sumcheck_native(s_n, eta, ps_n)
    finals_n = [s.final_column_evals() for s in s_n]
    assert ps_n.transcript == ps_py.transcript
    assert np.array_equal(np.stack(ch_n), np.stack(ch_py))
    for a, b in zip(finals_n, finals_py):
        assert np.array_equal(a, b)
    for s in s_py + s_n:
        s.free()
    ps_n.free()

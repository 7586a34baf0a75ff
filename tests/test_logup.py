"""Logup table assembly + quotient GKR over the three lean_vm tables.

The traces are synthetic but CONSISTENT: memory lookups hit the memory, the execution table's instruction columns are
rows of the bytecode, and every precompile row is pushed once by the execution table's bus, so the logup sum is zero
exactly as prove_generic_logup asserts (logup.rs:221).  CPU tier: the oracle assembles, proves and verifies
(verify_generic_logup).  GPU tier: the device-assembled table equals the oracle's, and the GPU prover's transcript is
accepted by the oracle verifier and equals the oracle prover's.
"""
import numpy as np
import pytest

import oracle as O
from oracle import logup as OL
from oracle import whir as W

ONE_M = int(O.to_monty(1))


def m(x):
    return O.to_monty(np.asarray(x, dtype=np.uint64))


def make_instance(rng, log_memory=10, log_bytecode=5, log_cycles=7, log_ext=6, log_pos=5):
    """canonical-integer model of a consistent set of tables"""
    mem = rng.integers(0, O.P, 1 << log_memory).astype(np.uint64)
    mem_acc = np.zeros(1 << log_memory, dtype=np.uint64)
    bytecode = np.zeros((1 << log_bytecode, 16), dtype=np.uint64)
    bytecode[:, :12] = rng.integers(0, O.P, (1 << log_bytecode, 12))
    bytecode_acc = np.zeros(1 << log_bytecode, dtype=np.uint64)

    def look(addr):  # one lookup of mem[addr]
        mem_acc[addr] += 1
        return mem[addr]

    # precompile tables first: every active row must be pushed once by an execution row whose bus data is
    # (precompile_data, nu_a, nu_b, nu_c).  precompile_data is instruction column 19, i.e. part of a bytecode row, so
    # the few distinct precompile_data values get dedicated bytecode rows.
    aux_values = [64 + 8, 2 * 64 + 16 + 4, 3 * 64 + 32, 64 + 16]
    pd_row = {1: 0}
    for i, v in enumerate(aux_values):
        pd_row[v] = i + 1
    for v, row in pd_row.items():
        bytecode[row, 11] = v
    n_ext, n_pos = 1 << log_ext, 1 << log_pos
    ext = np.zeros((31, n_ext), dtype=np.uint64)
    pos = np.zeros((111, n_pos), dtype=np.uint64)
    pushes = []  # (precompile_data, a, b, c)
    for r in range(n_ext):
        active = r < n_ext - 5
        ia, ib, ir = (int(rng.integers(0, (1 << log_memory) - 8)) for _ in range(3))
        ext[6, r], ext[7, r], ext[13, r] = ia, ib, ir
        for k in range(5):
            ext[14 + k, r], ext[19 + k, r], ext[24 + k, r] = look(ia + k), look(ib + k), look(ir + k)
        ext[29, r] = 1 if active else 0                       # activation flag (virtual column, bus selector)
        ext[30, r] = aux_values[r % 4]                        # aux (virtual column)
        if active:
            pushes.append((int(ext[30, r]), ia, ib, ir))
    for r in range(n_pos):
        active = r < n_pos - 3
        il, ib, ir = (int(rng.integers(0, (1 << log_memory) - 20)) for _ in range(3))
        pos[0, r] = 1 if active else 0
        pos[1, r], pos[2, r], pos[6, r], pos[7, r] = ib, ir, il, il + 4
        for k in range(4):
            pos[9 + k, r], pos[13 + k, r] = look(il + k), look(il + 4 + k)
        for k in range(8):
            pos[17 + k, r] = look(ib + k)
        for k in range(16):
            pos[93 + k, r] = look(ir + k)
        pos[109, r], pos[110, r] = il, 1                     # index_input_left, precompile_data (virtual columns)
        if active:
            pushes.append((1, il, ib, ir))
    n_cyc = 1 << log_cycles
    assert len(pushes) <= n_cyc
    ex = np.zeros((24, n_cyc), dtype=np.uint64)
    for r in range(n_cyc):
        if r < len(pushes):
            pd, a, b, c = pushes[r]
            pc = pd_row[pd]
            ex[20, r], ex[21, r], ex[22, r], ex[23, r] = 1, a, b, c   # is_precompile, nu_a, nu_b, nu_c
        else:
            pc = int(rng.integers(0, 1 << log_bytecode))
        bytecode_acc[pc] += 1
        ex[0, r] = pc
        ex[8:20, r] = bytecode[pc, :12]
        for k in range(3):
            addr = int(rng.integers(0, 1 << log_memory))
            ex[2 + k, r], ex[5 + k, r] = addr, look(addr)
    return dict(memory=mem, memory_acc=mem_acc, bytecode=bytecode, bytecode_acc=bytecode_acc, ex=ex, ext=ext, pos=pos,
                logs=(log_memory, log_bytecode, log_cycles, log_ext, log_pos))


def to_arrays(inst):
    cols = lambda a: [np.ascontiguousarray(m(a[c])) for c in range(a.shape[0])]
    lm_, lb, lc, le, lp = inst["logs"]
    traces = {"execution": (cols(inst["ex"]), lc), "extension_op": (cols(inst["ext"]), le), "poseidon16": (cols(inst["pos"]), lp)}
    return (m(inst["memory"]), m(inst["memory_acc"] % O.P), m(inst["bytecode"].reshape(-1)), m(inst["bytecode_acc"] % O.P), traces)


def challenges(rng):
    alphas = O.random_field(rng, (5, 5))           # log2(32) alphas -> eq poly of 32 entries (> 12 + 1 data columns)
    return O.random_field(rng, 5), alphas, O.eq_table(alphas)


def oracle_prove(c, al_eq, memory, memory_acc, bytecode, bytecode_acc, traces):
    nums, dens = OL.build_table(c, al_eq, memory, memory_acc, bytecode, bytecode_acc, traces)
    ps = W.ProverState()
    quotient, point, _, _ = OL.prove_gkr_quotient_cpu(ps, nums, dens)
    assert quotient == W.ZERO
    pt = lambda k: W._pts(point[len(point) - k:])
    add = lambda v: ps.add_extension_scalars(v)
    lm_ = memory.size.bit_length() - 1
    lb = bytecode_acc.size.bit_length() - 1
    add(O.mle_eval(memory_acc, pt(lm_))), add(O.mle_eval(memory, pt(lm_))), add(O.mle_eval(bytecode_acc, pt(lb)))
    cc = W.fm(c)
    al = [W.fm(a) for a in al_eq]
    for name, h in OL.sort_tables_by_height({k: v[1] for k, v in traces.items()}):
        cols = traces[name][0]
        _, pull, selector, bus_data, lookups = OL.TABLES[name]
        ev = lambda col: O.mle_eval(col, pt(h))
        if name == "execution":
            add(ev(cols[0]))
            add(np.concatenate([ev(cols[8 + k]) for k in range(12)]))
        sel = W.fm(ev(cols[selector]))
        add(W.tm(W.sub(W.ZERO, sel) if pull else sel))
        add(W.tm(W.add(cc, OL.finger_print(1, [W.fm(ev(cols[k])) for k in bus_data], al))))
        for index, values in lookups:
            add(ev(cols[index]))
            for v in values:
                add(ev(cols[v]))
    return ps, nums, dens


@pytest.fixture(scope="module")
def instance():
    rng = np.random.default_rng(77)
    inst = make_instance(rng)
    return to_arrays(inst), challenges(rng)


def test_synthetic_instance_balances(instance):
    (memory, memory_acc, bytecode, bytecode_acc, traces), (c, alphas, al_eq) = instance
    nums, dens = OL.build_table(c, al_eq, memory, memory_acc, bytecode, bytecode_acc, traces)
    tot = W.ZERO
    for a, b in zip(O.from_monty(nums), dens):
        if a:
            tot = W.add(tot, W.scal(W.inv(W.fm(b)), int(a)))
    assert tot == W.ZERO


def test_oracle_logup_prove_verify(instance):
    (memory, memory_acc, bytecode, bytecode_acc, traces), (c, alphas, al_eq) = instance
    ps, _, _ = oracle_prove(c, al_eq, memory, memory_acc, bytecode, bytecode_acc, traces)
    heights = {k: v[1] for k, v in traces.items()}
    vs = W.VerifierState(ps.transcript, [])
    st = OL.verify_generic_logup(vs, c, alphas, al_eq, memory.size.bit_length() - 1, bytecode, heights)
    assert vs.off == len(ps.transcript)
    assert st["columns_values"]["poseidon16"][93] == W.fm(O.mle_eval(traces["poseidon16"][0][93], W._pts(st["gkr_point"][-5:])))
    # a wrong column evaluation is caught
    bad = list(ps.transcript)
    bad[-7] = (bad[-7] + 1) % O.P
    with pytest.raises(W.ProofError):
        OL.verify_generic_logup(W.VerifierState(bad, []), c, alphas, al_eq, memory.size.bit_length() - 1, bytecode, heights)


def test_oracle_logup_rejects_unbalanced_memory(instance):
    (memory, memory_acc, bytecode, bytecode_acc, traces), (c, alphas, al_eq) = instance
    acc = memory_acc.copy()
    acc[3] = int(O.to_monty((int(O.from_monty(acc[3])) + 1) % O.P))
    nums, dens = OL.build_table(c, al_eq, memory, acc, bytecode, bytecode_acc, traces)
    ps = W.ProverState()
    quotient, *_ = OL.prove_gkr_quotient_cpu(ps, nums, dens)
    assert quotient != W.ZERO


# ------------------------------------------------------------------------------------------------ GPU tier
def product_traces(traces):
    from leanmultisig_b200 import tables as T

    by_name = {t.name: t for t in T.ALL_TABLES}
    return {by_name[k]: T.TableTrace(v[0], v[1]) for k, v in traces.items()}


@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as lm

    c = lm.Context(0, 20)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_logup_table_matches_oracle(ctx, instance):
    from leanmultisig_b200.logup import build_logup_table

    (memory, memory_acc, bytecode, bytecode_acc, traces), (c, alphas, al_eq) = instance
    exp_n, exp_d = OL.build_table(c, al_eq, memory, memory_acc, bytecode, bytecode_acc, traces)
    b = build_logup_table(ctx, c, al_eq, memory, memory_acc, bytecode, bytecode_acc, product_traces(traces))
    got_n, got_d = b.read(exp_n.size)
    assert np.array_equal(got_n, exp_n) and np.array_equal(got_d, exp_d)
    col = traces["poseidon16"][0][17]
    pt = O.random_field(np.random.default_rng(1), (5, 5))
    assert np.array_equal(b.col_eval(col, 5, pt), O.mle_eval(col, pt))
    b.free()


@pytest.mark.gpu
def test_gpu_prove_generic_logup_verifies_and_matches_oracle_transcript(ctx, instance):
    import leanmultisig_b200 as lm
    from leanmultisig_b200.logup import prove_generic_logup

    (memory, memory_acc, bytecode, bytecode_acc, traces), (c, alphas, al_eq) = instance
    ps = lm.ProverState(ctx)
    st = prove_generic_logup(ctx, ps, c, al_eq, memory, memory_acc, bytecode, bytecode_acc, product_traces(traces))
    ps_o, _, _ = oracle_prove(c, al_eq, memory, memory_acc, bytecode, bytecode_acc, traces)
    assert ps.transcript == ps_o.transcript
    heights = {k: v[1] for k, v in traces.items()}
    vs = W.VerifierState(ps.transcript, [])
    sv = OL.verify_generic_logup(vs, c, alphas, al_eq, memory.size.bit_length() - 1, bytecode, heights)
    assert vs.off == len(ps.transcript)
    assert [W.fm(x) for x in st["gkr_point"]] == sv["gkr_point"]
    assert W.fm(st["value_memory"]) == sv["value_memory"]


@pytest.mark.gpu
def test_gpu_logup_larger_instance_verifies(ctx):
    """2^16 memory, 2^13 cycles: too slow for the pure-Python oracle prover, the verifier still closes the loop"""
    import leanmultisig_b200 as lm
    from leanmultisig_b200.logup import prove_generic_logup

    rng = np.random.default_rng(78)
    inst = make_instance(rng, log_memory=16, log_bytecode=8, log_cycles=13, log_ext=10, log_pos=9)
    memory, memory_acc, bytecode, bytecode_acc, traces = to_arrays(inst)
    c, alphas, al_eq = challenges(rng)
    ps = lm.ProverState(ctx)
    prove_generic_logup(ctx, ps, c, al_eq, memory, memory_acc, bytecode, bytecode_acc, product_traces(traces))
    vs = W.VerifierState(ps.transcript, [])
    OL.verify_generic_logup(vs, c, alphas, al_eq, 16, bytecode, {k: v[1] for k, v in traces.items()})
    assert vs.off == len(ps.transcript)


def test_mle_of_zeros_then_ones_is_the_evaluation_of_that_slice():
    """The reference's own test (crates/backend/poly/src/mle/mle_custom.rs:32-44): the closed form used by the Logup verifier
    for padded tables equals the multilinear evaluation of [0; n_zeros] ++ [1; 2^n - n_zeros], for every split."""
    from oracle import logup as OL
    from oracle import whir as W

    rng = np.random.default_rng(0)
    for n_vars in range(0, 7):
        for n_zeros in range(0, (1 << n_vars) + 1):
            point = [tuple(int(x) for x in rng.integers(0, W.P, 5)) for _ in range(n_vars)]
            values = [W.ZERO] * n_zeros + [W.ONE] * ((1 << n_vars) - n_zeros)
            assert OL.mle_of_zeros_then_ones(n_zeros, point) == W.mle_eval_small(values, point), (n_vars, n_zeros)


def test_mle_of_the_index_column_is_its_evaluation():
    """mle_of_01234567_etc (crates/backend/poly/src/mle/mle_custom.rs): the closed form of the multilinear extension of
    (0, 1, 2, ..., 2^n - 1), the implicit index column of the Logup memory / bytecode tables, against the plain evaluation."""
    from oracle import logup as OL
    from oracle import whir as W

    rng = np.random.default_rng(1)
    for n_vars in range(0, 8):
        point = [tuple(int(x) for x in rng.integers(0, W.P, 5)) for _ in range(n_vars)]
        values = [(i, 0, 0, 0, 0) for i in range(1 << n_vars)]
        assert OL.mle_of_01234567_etc(point) == W.mle_eval_small(values, point), n_vars

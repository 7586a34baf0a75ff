"""Witness-side steps in front of the first commitment: access counts and the stacked commit.
CPU tier: the oracle restatements against direct loops / the layout arithmetic.  GPU tier: lm_access_counts and
lm_commit_stacked against the oracle (counts, root, codeword, OOD evaluation of the stacked polynomial)."""
import numpy as np
import pytest

import oracle as O
from leanmultisig_b200 import tables as T
from leanmultisig_b200.stacked_pcs import compute_stacked_n_vars, stacked_layout


def make_traces(rng, log_memory, log_cycles, log_ext, log_pos):
    """random tables whose lookup index columns stay inside the memory (values need not be consistent here)"""
    traces = {}
    for table, log_n in ((T.EXECUTION, log_cycles), (T.EXTENSION_OP, log_ext), (T.POSEIDON16, log_pos)):
        cols = [np.ascontiguousarray(O.random_field(rng, 1 << log_n)) for _ in range(table.n_columns_total)]
        for lk in table.lookups:
            cols[lk.index] = O.to_monty(rng.integers(0, (1 << log_memory) - len(lk.values), 1 << log_n).astype(np.uint64))
        traces[table] = T.TableTrace(cols, log_n)
    return traces


def test_oracle_access_counts_is_the_reference_loop(rng):
    log_memory = 7
    traces = make_traces(rng, log_memory, 6, 4, 3)
    cols, nv = [], []
    for table, tr in traces.items():
        for lk in table.lookups:
            cols.append(tr.columns[lk.index]), nv.append(len(lk.values))
    got = O.from_monty(O.access_counts(cols, nv, 1 << log_memory))
    exp = np.zeros(1 << log_memory, dtype=np.uint64)
    for col, n in zip(cols, nv):          # prove_execution.rs:94-101 verbatim
        for i in O.from_monty(col):
            for j in range(n):
                exp[int(i) + j] += 1
    assert np.array_equal(got, exp) and exp.sum() == sum(c.size * n for c, n in zip(cols, nv))


def test_stacked_layout_matches_oracle_stacking(rng):
    log_memory, log_bytecode = 8, 5
    traces = make_traces(rng, log_memory, 7, 7, 4)  # a tie in heights: the Table-enum order decides (table_trait.rs:66-70)
    memory, memory_acc = O.random_field(rng, 1 << log_memory), O.random_field(rng, 1 << log_memory)
    bytecode_acc = O.random_field(rng, 1 << log_bytecode)
    order = T.sort_tables_by_height({t: tr.log_n_rows for t, tr in traces.items()})
    g, n_vars, end = O.stack_polynomials(memory, memory_acc, bytecode_acc,
                                         [(traces[t].columns[: t.n_columns], h) for t, h in order])
    assert n_vars == compute_stacked_n_vars(log_memory, log_bytecode, {t: tr.log_n_rows for t, tr in traces.items()})
    layout, end2 = stacked_layout(1 << log_memory, 1 << log_bytecode, traces)
    assert end2 == end and [name for name, *_ in layout[:3]] == ["memory", "memory_acc", "bytecode_acc"]
    by_name = {t.name: tr for t, tr in traces.items()}
    for name, c, off, length in layout[3:]:
        assert np.array_equal(g[off:off + length], by_name[name].columns[c])
    assert layout[3][0] == "execution" and not g[end:].any()


@pytest.fixture(scope="module")
def ctx():
    import leanmultisig_b200 as lm

    c = lm.Context(0, 22)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("log_memory,log_cycles,log_ext,log_pos", [(9, 8, 6, 5), (14, 13, 11, 10)])
def test_gpu_access_counts(ctx, rng, log_memory, log_cycles, log_ext, log_pos):
    import leanmultisig_b200 as lm
    from leanmultisig_b200.stacked_pcs import build_bytecode_acc, build_memory_acc

    traces = make_traces(rng, log_memory, log_cycles, log_ext, log_pos)
    cols, nv = [], []
    for table, tr in traces.items():
        for lk in table.lookups:
            cols.append(tr.columns[lk.index]), nv.append(len(lk.values))
    assert np.array_equal(build_memory_acc(ctx, 1 << log_memory, traces), O.access_counts(cols, nv, 1 << log_memory))
    ex = traces[T.EXECUTION]
    ex.columns[T.COL_PC] = O.to_monty(rng.integers(0, 32, 1 << log_cycles).astype(np.uint64))
    assert np.array_equal(build_bytecode_acc(ctx, 32, ex), O.access_counts([ex.columns[T.COL_PC]], [1], 32))
    # an address outside the table is an error, as the reference's slice index panics
    bad = O.to_monty(np.array([3, (1 << log_memory) - 1], dtype=np.uint64))
    with pytest.raises(lm.LmError):
        lm.stacked_pcs.access_counts(ctx, [bad], [2], 1 << log_memory)


@pytest.mark.gpu
@pytest.mark.parametrize("log_memory,log_bytecode,log_cycles,log_ext,log_pos", [(10, 5, 9, 7, 6), (13, 14, 12, 12, 8)])
def test_gpu_stacked_commit(ctx, rng, log_memory, log_bytecode, log_cycles, log_ext, log_pos):
    from leanmultisig_b200.stacked_pcs import stack_polynomials_and_commit

    traces = make_traces(rng, log_memory, log_cycles, log_ext, log_pos)
    memory, memory_acc = O.random_field(rng, 1 << log_memory), O.random_field(rng, 1 << log_memory)
    bytecode_acc = O.random_field(rng, 1 << log_bytecode)
    order = T.sort_tables_by_height({t: tr.log_n_rows for t, tr in traces.items()})
    g, n_vars, end = O.stack_polynomials(memory, memory_acc, bytecode_acc,
                                         [(traces[t].columns[: t.n_columns], h) for t, h in order])
    k, r = 7, 1
    tree, nv, actual = stack_polynomials_and_commit(ctx, k, r, memory, memory_acc, bytecode_acc, traces)
    assert (nv, actual) == (n_vars, end)
    cw = O.reorder_and_dft(g, n_vars, 1, k, r, tree.stored_width)
    eff = -(-end // (1 << (n_vars - k)))
    layers = O.merkle_tree(cw, 1 << k, eff)
    assert np.array_equal(tree.codeword(), cw) and np.array_equal(tree.root, layers[-1])
    pt = O.expand_from_univariate(O.random_field(rng, 5), n_vars)
    assert np.array_equal(tree.evaluate(pt), O.mle_eval(g, pt))
    tree.free()

"""Witness-side steps in front of the first commitment, on top of the C ABI.

Reference names kept: compute_stacked_n_vars / stack_polynomials_and_commit (crates/sub_protocols/src/stacked_pcs.rs:
100-157, 183-196) and the memory_acc / bytecode_acc loops of prove_execution (crates/lean_prover/src/prove_execution.rs:
91-110).  The reference assembles a host-side global_polynomial (a copy of the whole witness) and counts accesses in a
sequential loop; here every column goes straight to its offset of the device-resident polynomial (lm_commit_stacked) and
the counts are device atomics (lm_access_counts).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, u32p
from .tables import COL_PC, sort_tables_by_height
from .whir import Tree


class _Segment(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint64), ("offset", C.c_uint64)]


def compute_stacked_n_vars(log_memory: int, log_bytecode: int, tables_log_heights: dict) -> int:
    """stacked_pcs.rs:183-196"""
    max_table = max(tables_log_heights.values())
    total = (2 << log_memory) + (1 << max(log_bytecode, max_table)) + sum(t.n_columns << h for t, h in tables_log_heights.items())
    return (total - 1).bit_length()


def stacked_layout(memory_len: int, bytecode_acc_len: int, traces: dict):
    """[(name, column index or None, offset, length)] in the order of stack_polynomials_and_commit (stacked_pcs.rs:118-135):
    memory, memory_acc, bytecode_acc (padded to the tallest table), then the tables tallest first, column after column.
    traces: {Table: TableTrace}."""
    heights = {t: tr.log_n_rows for t, tr in traces.items()}
    order = sort_tables_by_height(heights)
    out = [("memory", None, 0, memory_len), ("memory_acc", None, memory_len, memory_len),
           ("bytecode_acc", None, 2 * memory_len, bytecode_acc_len)]
    offset = 2 * memory_len + max(1 << order[0][1], bytecode_acc_len)
    for table, log_n in order:
        for c in range(table.n_columns):
            out.append((table.name, c, offset, 1 << log_n))
            offset += 1 << log_n
    return out, offset


def access_counts(ctx, index_columns, n_values, table_len: int) -> np.ndarray:
    """acc[addr + j] += 1 for every entry `addr` of index_columns[k] and j < n_values[k]; Montgomery in, Montgomery out"""
    cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in index_columns]
    ptrs = (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
    rows = np.array([c.size for c in cols] or [0], dtype=np.uint64)
    nv = np.array(list(n_values) or [0], dtype=np.uint32)
    out = np.empty(table_len, dtype=np.uint32)
    check(lib().lm_access_counts(ctx.handle, ptrs, rows.ctypes.data_as(C.POINTER(C.c_uint64)), nv.ctypes.data_as(u32p), len(cols),
                                 table_len, out.ctypes.data_as(u32p)))
    return out


def build_memory_acc(ctx, memory_len: int, traces: dict) -> np.ndarray:
    """prove_execution.rs:91-103 over all tables' lookups"""
    cols, nv = [], []
    for table, trace in traces.items():
        for lookup in table.lookups:
            cols.append(trace.columns[lookup.index])
            nv.append(len(lookup.values))
    return access_counts(ctx, cols, nv, memory_len)


def build_bytecode_acc(ctx, bytecode_padded_len: int, execution_trace) -> np.ndarray:
    """prove_execution.rs:105-110"""
    return access_counts(ctx, [execution_trace.columns[COL_PC]], [1], bytecode_padded_len)


def stack_polynomials_and_commit(ctx, folding_factor: int, log_inv_rate: int, memory, memory_acc, bytecode_acc, traces: dict):
    """-> (Tree, stacked_n_vars, actual_len).  The caller adds the root to its transcript and samples the OOD points as
    WhirConfig::commit does (leanmultisig_b200.whir.WhirProver.commit does both for a host polynomial)."""
    memory, memory_acc, bytecode_acc = (np.ascontiguousarray(a, dtype=np.uint32) for a in (memory, memory_acc, bytecode_acc))
    assert memory.size == memory_acc.size
    heights = {t: tr.log_n_rows for t, tr in traces.items()}
    execution = next(t for t in traces if t.is_execution)
    assert memory.size.bit_length() - 1 >= heights[execution] and heights[execution] >= max(heights.values())
    layout, end = stacked_layout(memory.size, bytecode_acc.size, traces)
    n_vars = compute_stacked_n_vars(memory.size.bit_length() - 1, bytecode_acc.size.bit_length() - 1, heights)
    assert (end - 1).bit_length() == n_vars
    by_name = {t.name: tr for t, tr in traces.items()}
    keep, segs = [], (_Segment * len(layout))()
    for i, (name, c, off, length) in enumerate(layout):
        a = {"memory": memory, "memory_acc": memory_acc, "bytecode_acc": bytecode_acc}.get(name)
        if a is None:
            a = np.ascontiguousarray(by_name[name].columns[c], dtype=np.uint32)[:length]
        keep.append(a)
        segs[i] = _Segment(a.ctypes.data, a.size, off)
    root = np.empty(8, dtype=np.uint32)
    t = C.c_void_p()
    check(lib().lm_commit_stacked(ctx.handle, C.byref(segs), len(layout), n_vars, folding_factor, log_inv_rate, C.byref(t),
                                  root.ctypes.data_as(u32p)))
    return Tree(ctx, t, root), n_vars, end

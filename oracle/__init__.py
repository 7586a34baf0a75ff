"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so``, the CPU restatement of the reference's
(leanEthereum/leanMultisig) hot-path algorithms.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``leanmultisig_b200``) never does.

Parity pin status is recorded in ``oracle/oracle.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

P = 0x7F000001
u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.lm_or_kb_mul.restype = C.c_uint32
        _lib.lm_or_kb_from_u32.restype = C.c_uint32
        _lib.lm_or_kb_to_u32.restype = C.c_uint32
        _lib.lm_or_kb_inv.restype = C.c_uint32
        _lib.lm_or_kb_two_adic_generator.restype = C.c_uint32
        _lib.lm_or_merkle_verify.restype = C.c_int
        _lib.lm_or_poseidon1_init()
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


# ---------------------------------------------------------------- field helpers
def to_monty(x) -> np.ndarray:
    """canonical integers -> Montgomery-form u32 (x * 2^32 mod p)."""
    x = np.asarray(x, dtype=np.uint64) % P
    return ((x << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


_RINV = pow(1 << 32, -1, P)


def from_monty(x) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64)
    return ((x * np.uint64(_RINV)) % np.uint64(P)).astype(np.uint32)


def random_field(rng: np.random.Generator, shape) -> np.ndarray:
    """Uniform Montgomery-form residues in [0, p) (the reference samples the Monty value directly,
    monty_31.rs:139-149)."""
    return rng.integers(0, P, size=shape, dtype=np.uint32)


def kb_mul(a: int, b: int) -> int:
    return lib().lm_or_kb_mul(C.c_uint32(a), C.c_uint32(b))


def kb_inv(a: int) -> int:
    return lib().lm_or_kb_inv(C.c_uint32(a))


def two_adic_generator(bits: int) -> int:
    return lib().lm_or_kb_two_adic_generator(C.c_uint32(bits))


def ef_mul(a, b) -> np.ndarray:
    a, b = _u32(a), _u32(b)
    out = np.empty(5, dtype=np.uint32)
    lib().lm_or_ef_mul(_p(a), _p(b), _p(out))
    return out


def ef_inv(a) -> np.ndarray:
    a = _u32(a)
    out = np.empty(5, dtype=np.uint32)
    lib().lm_or_ef_inv(_p(a), _p(out))
    return out


# ---------------------------------------------------------------- Poseidon1
def poseidon1_permute(states, dense: bool = False) -> np.ndarray:
    s = _u32(states).copy().reshape(-1, 16)
    lib().lm_or_poseidon1_permute_batch(_p(s), C.c_uint64(s.shape[0]), C.c_int(1 if dense else 0))
    return s.reshape(np.shape(states))


def poseidon1_compress(states) -> np.ndarray:
    s = _u32(states).copy().reshape(-1, 16)
    lib().lm_or_poseidon1_compress_batch(_p(s), C.c_uint64(s.shape[0]))
    return s.reshape(np.shape(states))


# ---------------------------------------------------------------- Merkle
def hash_slice(data) -> np.ndarray:
    d = _u32(data)
    out = np.empty(8, dtype=np.uint32)
    lib().lm_or_hash_slice(_p(d), C.c_uint64(d.size), _p(out))
    return out


def zero_suffix_state(n_zero_chunks: int) -> np.ndarray:
    out = np.empty(16, dtype=np.uint32)
    lib().lm_or_zero_suffix_state(C.c_uint32(n_zero_chunks), _p(out))
    return out


def first_digest_layer(mat, full_width: int, effective_width: int) -> np.ndarray:
    m = _u32(mat)
    h, w = m.shape
    out = np.empty((h, 8), dtype=np.uint32)
    lib().lm_or_first_digest_layer(_p(m), C.c_uint64(h), C.c_uint32(w), C.c_uint32(full_width),
                                   C.c_uint32(effective_width), _p(out))
    return out


def merkle_tree(mat, full_width: int, effective_width: int) -> np.ndarray:
    """All digest layers back to back: (2h-1) x 8; root = last row."""
    m = _u32(mat)
    h, w = m.shape
    out = np.empty((2 * h - 1, 8), dtype=np.uint32)
    lib().lm_or_merkle_tree(_p(m), C.c_uint64(h), C.c_uint32(w), C.c_uint32(full_width),
                            C.c_uint32(effective_width), _p(out))
    return out


def merkle_open(mat, full_width: int, layers, index: int):
    m, l = _u32(mat), _u32(layers)
    h, w = m.shape
    log_h = h.bit_length() - 1
    row = np.empty(full_width, dtype=np.uint32)
    path = np.empty((log_h, 8), dtype=np.uint32)
    lib().lm_or_merkle_open(_p(m), C.c_uint64(h), C.c_uint32(w), C.c_uint32(full_width), _p(l),
                            C.c_uint64(index), _p(row), _p(path))
    return row, path


def merkle_verify(root, log_h: int, index: int, row, path) -> bool:
    root, row, path = _u32(root), _u32(row), _u32(path)
    return bool(lib().lm_or_merkle_verify(_p(root), C.c_uint32(log_h), C.c_uint64(index), _p(row),
                                          C.c_uint32(row.size), _p(path)))


# ---------------------------------------------------------------- RS encode
def prepare_evals(evals, n_vars: int, dim: int, folding: int, log_inv_rate: int, dft_n_cols: int) -> np.ndarray:
    e = _u32(evals)
    h = 1 << (n_vars + log_inv_rate - folding)
    out = np.empty((h, dft_n_cols * dim), dtype=np.uint32)
    lib().lm_or_prepare_evals(_p(e), C.c_uint32(n_vars), C.c_uint32(dim), C.c_uint32(folding),
                              C.c_uint32(log_inv_rate), C.c_uint32(dft_n_cols), _p(out))
    return out


def dft_batch_by_evals(mat) -> np.ndarray:
    m = _u32(mat).copy()
    h, w = m.shape
    lib().lm_or_dft_batch_by_evals(_p(m), C.c_uint64(h), C.c_uint64(w))
    return m


def reorder_and_dft(evals, n_vars: int, dim: int, folding: int, log_inv_rate: int, dft_n_cols: int) -> np.ndarray:
    e = _u32(evals)
    h = 1 << (n_vars + log_inv_rate - folding)
    out = np.empty((h, dft_n_cols * dim), dtype=np.uint32)
    lib().lm_or_reorder_and_dft(_p(e), C.c_uint32(n_vars), C.c_uint32(dim), C.c_uint32(folding),
                                C.c_uint32(log_inv_rate), C.c_uint32(dft_n_cols), _p(out))
    return out


def dft_layers_mapped(mat, log_h: int, l_first: int, n_blocks: int, run: int, block: int, offset: int) -> np.ndarray:
    m = _u32(mat).copy()
    lib().lm_or_dft_layers_mapped(_p(m), C.c_uint64(m.shape[1]), C.c_uint32(log_h), C.c_uint32(l_first),
                                  C.c_uint64(n_blocks), C.c_uint64(run), C.c_uint64(block), C.c_uint64(offset))
    return m


# ---------------------------------------------------------------- multilinear
def eq_table(point, scalar=None) -> np.ndarray:
    pt = _u32(point).reshape(-1, 5)
    k = pt.shape[0]
    sc = _u32(scalar) if scalar is not None else np.array([to_monty(1), 0, 0, 0, 0], dtype=np.uint32)
    out = np.empty((1 << k, 5), dtype=np.uint32)
    lib().lm_or_eq_table(_p(pt), C.c_uint32(k), _p(sc), _p(out))
    return out


def expand_from_univariate(y, n: int) -> np.ndarray:
    y = _u32(y)
    out = np.empty((n, 5), dtype=np.uint32)
    lib().lm_or_expand_from_univariate(_p(y), C.c_uint32(n), _p(out))
    return out


def mle_eval(evals, point) -> np.ndarray:
    e = _u32(evals)
    pt = _u32(point).reshape(-1, 5)
    n = pt.shape[0]
    dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
    assert e.size == (1 << n) * dim
    out = np.empty(5, dtype=np.uint32)
    lib().lm_or_mle_eval(_p(e), C.c_uint32(n), C.c_uint32(dim), _p(pt), _p(out))
    return out


def fold_msb(evals, r) -> np.ndarray:
    e, r = _u32(evals), _u32(r)
    dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
    n = e.size // dim
    out = np.empty((n // 2, 5), dtype=np.uint32)
    lib().lm_or_fold_msb(_p(e), C.c_uint64(n), C.c_uint32(dim), _p(r), _p(out))
    return out


# ---------------------------------------------------------------- WHIR open: weights + product sumcheck
def weights_add_eq(weights, selector: int, point, scalar) -> None:
    """in place: weights[(selector << m) + x] += scalar * eq(point, x)"""
    pt, sc = _u32(point).reshape(-1, 5), _u32(scalar)
    lib().lm_or_weights_add_eq(_p(weights), C.c_uint64(selector), _p(pt), C.c_uint32(pt.shape[0]), _p(sc))


def next_mle_folded(point) -> np.ndarray:
    pt = _u32(point).reshape(-1, 5)
    out = np.empty((1 << pt.shape[0], 5), dtype=np.uint32)
    lib().lm_or_next_mle_folded(_p(pt), C.c_uint32(pt.shape[0]), _p(out))
    return out


def weights_add_next(weights, selector: int, point, scalar) -> None:
    pt, sc = _u32(point).reshape(-1, 5), _u32(scalar)
    lib().lm_or_weights_add_next(_p(weights), C.c_uint64(selector), _p(pt), C.c_uint32(pt.shape[0]), _p(sc))


def weights_add_base_eq(weights, points, scalars) -> None:
    pts, sc = _u32(points), _u32(scalars).reshape(-1, 5)
    lib().lm_or_weights_add_base_eq(_p(weights), C.c_uint32(pts.shape[1]), _p(pts), C.c_uint32(pts.shape[0]), _p(sc))


def prod_round(p, w):
    p, w = _u32(p), _u32(w)
    dim = 5 if (p.ndim == 2 and p.shape[1] == 5) else 1
    n = w.shape[0]
    c0, c2 = np.empty(5, dtype=np.uint32), np.empty(5, dtype=np.uint32)
    lib().lm_or_prod_round(_p(p), C.c_uint32(dim), _p(w), C.c_uint64(n), _p(c0), _p(c2))
    return c0, c2


def evals_to_coeffs(data) -> np.ndarray:
    d = _u32(data).copy().reshape(-1, 5)
    lib().lm_or_evals_to_coeffs(_p(d), C.c_uint64(d.shape[0]))
    return d


def ef_add(a, b) -> np.ndarray:
    return ((_u32(a).astype(np.uint64) + _u32(b)) % P).astype(np.uint32)


def ef_sub(a, b) -> np.ndarray:
    return ((_u32(a).astype(np.uint64) + P - _u32(b)) % P).astype(np.uint32)


# ---------------------------------------------------------------- AIR sumcheck (execution table)
EXEC_N_COLS, EXEC_N_SHIFT, EXEC_N_CONSTRAINTS, EXEC_DEGREE = 20, 2, 13, 5


def air_exec_eval(point, alpha_powers, la, beta) -> np.ndarray:
    pt, ap, la, beta = _u32(point).reshape(22, 5), _u32(alpha_powers).reshape(-1, 5), _u32(la).reshape(-1, 5), _u32(beta)
    out = np.empty(5, dtype=np.uint32)
    lib().lm_or_air_exec_eval(_p(pt), _p(ap), _p(la), C.c_uint32(la.shape[0]), _p(beta), _p(out))
    return out


def shift_column(col) -> np.ndarray:
    c = _u32(col)
    out = np.empty_like(c)
    lib().lm_or_shift_column(_p(c), C.c_uint64(c.size), _p(out))
    return out


def air_exec_round(cols, eq_point, alpha_powers, la, beta) -> np.ndarray:
    """cols: (22, n) base or (22, n, 5) extension; returns evaluations at z = 0, 2, 3, 4, 5 (5 x 5)."""
    c = _u32(cols)
    dim = 5 if c.ndim == 3 else 1
    n = c.shape[1]
    eqp, ap, la, beta = _u32(eq_point).reshape(-1, 5), _u32(alpha_powers).reshape(-1, 5), _u32(la).reshape(-1, 5), _u32(beta)
    out = np.empty((5, 5), dtype=np.uint32)
    lib().lm_or_air_exec_round(_p(c), C.c_uint64(n), C.c_uint32(dim), _p(eqp), _p(ap), _p(la), C.c_uint32(la.shape[0]),
                               _p(beta), _p(out))
    return out


AIR_EXEC, AIR_EXT_OP, AIR_POSEIDON16, AIR_NO_BUS = 0, 1, 2, 0x100


def air_shape(table: int):
    """(n_cols, n_shift, degree) of a table's AIR"""
    out = (C.c_uint32 * 3)()
    assert lib().lm_or_air_shape(C.c_uint32(table), out) == 0
    return int(out[0]), int(out[1]), int(out[2])


def air_eval(table: int, point, alpha_powers, la, beta) -> np.ndarray:
    nc, ns, _ = air_shape(table)
    pt, ap, la, beta = _u32(point).reshape(nc + ns, 5), _u32(alpha_powers).reshape(-1, 5), _u32(la).reshape(-1, 5), _u32(beta)
    out = np.empty(5, dtype=np.uint32)
    lib().lm_or_air_eval(C.c_uint32(table), _p(pt), _p(ap), _p(la), C.c_uint32(la.shape[0]), _p(beta), _p(out))
    return out


def air_round(table: int, cols, eq_point, alpha_powers, la, beta) -> np.ndarray:
    """cols: (n_cols + n_shift, n) base or (.., n, 5) extension; returns evaluations at z = 0, 2, .., degree"""
    nc, ns, deg = air_shape(table)
    c = _u32(cols)
    assert c.shape[0] == nc + ns
    dim = 5 if c.ndim == 3 else 1
    n = c.shape[1]
    eqp, ap, la, beta = _u32(eq_point).reshape(-1, 5), _u32(alpha_powers).reshape(-1, 5), _u32(la).reshape(-1, 5), _u32(beta)
    out = np.empty((deg, 5), dtype=np.uint32)
    lib().lm_or_air_round(C.c_uint32(table), _p(c), C.c_uint64(n), C.c_uint32(dim), _p(eqp), _p(ap), _p(la),
                          C.c_uint32(la.shape[0]), _p(beta), _p(out))
    return out


def poseidon16_fill_trace(cols) -> np.ndarray:
    """cols: (109, n) base field with the 9 control columns and the 16 inputs set; returns the completed trace"""
    c = _u32(cols).copy()
    assert c.shape[0] == 109
    lib().lm_or_poseidon16_fill_trace(_p(c), C.c_uint64(c.shape[1]))
    return c


def fold_lsb(col, r) -> np.ndarray:
    c, r = _u32(col), _u32(r)
    dim = 5 if (c.ndim == 2 and c.shape[1] == 5) else 1
    n = c.size // dim
    out = np.empty((n // 2, 5), dtype=np.uint32)
    lib().lm_or_fold_lsb(_p(c), C.c_uint64(n), C.c_uint32(dim), _p(r), _p(out))
    return out


# ---------------------------------------------------------------- Logup / quotient GKR
def gkr_layer_up(nums, dens):
    n_, d_ = _u32(nums), _u32(dens).reshape(-1, 5)
    num_dim = 5 if (n_.ndim == 2 and n_.shape[1] == 5) else 1
    n = d_.shape[0]
    on, od = np.empty((n // 2, 5), dtype=np.uint32), np.empty((n // 2, 5), dtype=np.uint32)
    lib().lm_or_gkr_layer_up(_p(n_), C.c_uint32(num_dim), _p(d_), C.c_uint64(n), _p(on), _p(od))
    return on, od


def gkr_round(nl, nr, dl, dr, eq_point, alpha):
    nl, nr, dl, dr = (_u32(x).reshape(-1, 5) for x in (nl, nr, dl, dr))
    eqp, alpha = _u32(eq_point).reshape(-1, 5), _u32(alpha)
    c0, c2 = np.empty(5, dtype=np.uint32), np.empty(5, dtype=np.uint32)
    lib().lm_or_gkr_round(_p(nl), _p(nr), _p(dl), _p(dr), C.c_uint64(nl.shape[0]), _p(eqp), _p(alpha), _p(c0), _p(c2))
    return c0, c2


def finger_print(data, alphas, c) -> np.ndarray:
    d, al, c = _u32(data), _u32(alphas).reshape(-1, 5), _u32(c)
    out = np.empty((d.shape[0], 5), dtype=np.uint32)
    lib().lm_or_finger_print(_p(d), C.c_uint64(d.shape[0]), C.c_uint32(d.shape[1]), _p(al), _p(c), _p(out))
    return out


def embed(base) -> np.ndarray:
    """base-field vector -> EF vector"""
    b = _u32(base).reshape(-1)
    out = np.zeros((b.size, 5), dtype=np.uint32)
    out[:, 0] = b
    return out


# ---------------------------------------------------------------- witness-side steps
def access_counts(index_columns, n_values, table_len: int) -> np.ndarray:
    """memory_acc / bytecode_acc (crates/lean_prover/src/prove_execution.rs:91-110): acc[addr + j] += 1 for every entry
    of every index column and j < n_values[k]; Montgomery-form addresses in, Montgomery-form counts out"""
    acc = np.zeros(table_len, dtype=np.uint64)
    for col, nv in zip(index_columns, n_values):
        addr = from_monty(_u32(col)).astype(np.int64)
        for j in range(nv):
            np.add.at(acc, addr + j, 1)
    return to_monty(acc % P)


def stack_polynomials(memory, memory_acc, bytecode_acc, tables_sorted):
    """global_polynomial of stack_polynomials_and_commit (crates/sub_protocols/src/stacked_pcs.rs:100-135).
    tables_sorted: [(columns (list of arrays), log_n_rows)] tallest first.  -> (evals padded to 2^n_vars, n_vars, offset)"""
    memory, memory_acc, bytecode_acc = _u32(memory), _u32(memory_acc), _u32(bytecode_acc)
    largest = 1 << tables_sorted[0][1]
    total = 2 * memory.size + max(largest, bytecode_acc.size) + sum(len(cols) << h for cols, h in tables_sorted)
    n_vars = (total - 1).bit_length()
    g = np.zeros(1 << n_vars, dtype=np.uint32)
    g[: memory.size] = memory
    off = memory.size
    g[off:off + memory_acc.size] = memory_acc
    off += memory_acc.size
    g[off:off + bytecode_acc.size] = bytecode_acc
    off += max(largest, bytecode_acc.size)
    for cols, h in tables_sorted:
        for c in cols:
            g[off:off + (1 << h)] = _u32(c)[: 1 << h]
            off += 1 << h
    assert off == total
    return g, n_vars, off

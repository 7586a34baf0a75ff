"""Host-side mirror of the Logup quotient-GKR prover on top of the C ABI.

Reference names kept: prove_gkr_quotient / prove_gkr_layer (crates/sub_protocols/src/quotient_gkr/mod.rs:31-141),
build_bare_from_coeffs (quotient_gkr/sumcheck_utils.rs:491-503), finger_print (crates/utils/src/multilinear.rs:76).
The transcript is abstracted as three callables so the tests can drive it with any challenge source:
  add_scalars(list of EF)           prover_state.add_extension_scalars
  add_sumcheck_poly(coeffs, alpha)  prover_state.add_sumcheck_polynomial(&bare.coeffs, Some(eq_alpha))
  sample() -> EF (5 Montgomery words)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import field as F
from ._lib import check, lib, u32p

N_VARS_TO_SEND_GKR_COEFFS = 5


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a):
    return a.ctypes.data_as(u32p)


def finger_print(ctx, data, alphas, c) -> np.ndarray:
    d, al, cc = _u32(data), _u32(alphas).reshape(-1, 5), _u32(c)
    out = np.empty((d.shape[0], 5), dtype=np.uint32)
    check(lib().lm_finger_print(ctx.handle, _p(d), d.shape[0], d.shape[1], _p(al), _p(cc), _p(out)))
    return out


def _mle_eval_small(values, point):
    """values: list of EF tuples (2^k), point: list of EF tuples (x_0 = MSB)"""
    cur = list(values)
    for x in point:
        half = len(cur) // 2
        cur = [F.add(cur[i], F.mul(x, F.sub(cur[i + half], cur[i]))) for i in range(half)]
    return cur[0]


def build_bare_from_coeffs(c0_raw, c2_raw, eq_alpha, sum_, mmf):
    c0 = F.mul(c0_raw, mmf)
    c2 = F.mul(c2_raw, mmf)
    h1 = F.mul(F.sub(sum_, F.mul(F.sub(F.ONE, eq_alpha), c0)), F.inv(eq_alpha))
    return [c0, F.sub(F.sub(h1, c0), c2), c2]


class GkrQuotientProver:
    def __init__(self, ctx, nums, dens):
        n_, d_ = _u32(nums).reshape(-1), _u32(dens).reshape(-1, 5)
        assert n_.size == d_.shape[0]
        h = C.c_void_p()
        check(lib().lm_gkr_new(ctx.handle, _p(n_), _p(d_), n_.size, C.byref(h)))
        self.handle = h
        nv = C.c_uint32()
        check(lib().lm_gkr_num_vars(h, C.byref(nv)))
        self.n_vars = nv.value

    def top(self):
        tn, td = np.empty((32, 5), dtype=np.uint32), np.empty((32, 5), dtype=np.uint32)
        check(lib().lm_gkr_top(self.handle, _p(tn), _p(td)))
        return tn, td

    # ---- device steps of one layer sumcheck; the sharded prover (leanmultisig_b200/sharded.py) overrides these five ----
    def _layer_begin(self, k: int, point_m: np.ndarray, alpha_m: np.ndarray) -> None:
        check(lib().lm_gkr_layer_begin(self.handle, k, _p(point_m), _p(alpha_m)))

    def _round(self):
        c0, c2 = np.empty(5, dtype=np.uint32), np.empty(5, dtype=np.uint32)
        check(lib().lm_gkr_round(self.handle, _p(c0), _p(c2)))
        return c0, c2

    def _fold(self, r_m: np.ndarray) -> None:
        check(lib().lm_gkr_fold(self.handle, _p(r_m)))

    def _layer_end(self) -> np.ndarray:
        inner = np.empty((4, 5), dtype=np.uint32)
        check(lib().lm_gkr_layer_end(self.handle, _p(inner)))
        return inner

    @classmethod
    def from_device(cls, ctx, d_nums: int, d_dens: int, active_len: int):
        """lm_gkr_new_dev: numerators (active_len F) and denominators (active_len x 5 words) already on the device"""
        h = C.c_void_p()
        check(lib().lm_gkr_new_dev(ctx.handle, C.c_void_p(d_nums), C.c_void_p(d_dens), active_len, C.byref(h)))
        return cls.from_handle(h)

    @classmethod
    def from_handle(cls, handle):
        self = cls.__new__(cls)
        self.handle = handle
        nv = C.c_uint32()
        check(lib().lm_gkr_num_vars(handle, C.byref(nv)))
        self.n_vars = nv.value
        return self

    def prove_with_state(self, prover_state):
        """prove_gkr_quotient driven by an FSProver (mod.rs:31-78): alpha of every layer is sampled after a duplex"""
        def sample_alpha():
            prover_state.duplex()
            return prover_state.sample()

        self._sample_alpha = sample_alpha
        try:
            return self.prove(lambda v: prover_state.add_extension_scalars(_u32(v).reshape(-1)),
                              prover_state.add_sumcheck_polynomial, prover_state.sample,
                              sample_point=lambda n: prover_state.sample_vec(n))
        finally:
            self._sample_alpha = None

    _sample_alpha = None

    def prove_native(self, native_state):
        """prove_gkr_quotient through lm_gkr_prove: top layer on the host, every layer sumcheck on the device INCLUDING the
        challenger (csrc/devfs.cuh), so no round costs a host round trip; same transcript and outputs as prove_with_state."""
        q, cn, cd = (np.empty(5, dtype=np.uint32) for _ in range(3))
        pt = np.empty((self.n_vars, 5), dtype=np.uint32)
        check(lib().lm_gkr_prove(self.handle, native_state.handle, _p(q), _p(pt), _p(cn), _p(cd)))
        return q, pt, cn, cd

    def prove_native_hostloop(self, native_state):
        """lm_gkr_prove_hostloop: the C++ driver with the sponge on the host, one synchronisation per round (cross-check of
        the device-resident challenger that prove_native uses)."""
        q, cn, cd = (np.empty(5, dtype=np.uint32) for _ in range(3))
        pt = np.empty((self.n_vars, 5), dtype=np.uint32)
        check(lib().lm_gkr_prove_hostloop(self.handle, native_state.handle, _p(q), _p(pt), _p(cn), _p(cd)))
        return q, pt, cn, cd

    def prove(self, add_scalars, add_sumcheck_poly, sample, sample_point=None):
        """returns (quotient, point, claim_num, claim_den) as Montgomery-form arrays"""
        tn, td = self.top()
        add_scalars(tn)
        add_scalars(td)
        top_n, top_d = [F.from_monty(v) for v in tn], [F.from_monty(v) for v in td]
        quotient = F.ZERO
        for a, b in zip(top_n, top_d):
            quotient = F.add(quotient, F.mul(a, F.inv(b)))
        if sample_point is not None:
            point = [F.from_monty(x) for x in sample_point(N_VARS_TO_SEND_GKR_COEFFS)]
        else:
            point = [F.from_monty(sample()) for _ in range(N_VARS_TO_SEND_GKR_COEFFS)]
        claim_num, claim_den = _mle_eval_small(top_n, point), _mle_eval_small(top_d, point)
        for k in range(N_VARS_TO_SEND_GKR_COEFFS, self.n_vars):
            point, claim_num, claim_den = self._prove_layer(k, point, claim_num, claim_den, add_scalars, add_sumcheck_poly,
                                                            sample)
        return (F.to_monty(quotient), np.stack([F.to_monty(x) for x in point]), F.to_monty(claim_num),
                F.to_monty(claim_den))

    def _prove_layer(self, k, point, claim_num, claim_den, add_scalars, add_sumcheck_poly, sample):
        alpha_m = _u32(self._sample_alpha() if self._sample_alpha else sample())
        alpha = F.from_monty(alpha_m)
        s = F.add(claim_num, F.mul(alpha, claim_den))
        mmf = F.ONE
        pt = np.stack([F.to_monty(x) for x in point])
        self._layer_begin(k, pt, alpha_m)
        remaining = list(point)
        q = []
        for _ in range(k):
            c0, c2 = self._round()
            eq_alpha = remaining[-1]
            bare = build_bare_from_coeffs(F.from_monty(c0), F.from_monty(c2), eq_alpha, s, mmf)
            add_sumcheck_poly(np.stack([F.to_monty(c) for c in bare]), F.to_monty(eq_alpha))
            r_m = _u32(sample())
            r = F.from_monty(r_m)
            eq_eval = F.add(F.mul(F.sub(F.ONE, eq_alpha), F.sub(F.ONE, r)), F.mul(eq_alpha, r))
            s = F.mul(eq_eval, F.poly_eval(bare, r))
            mmf = F.mul(mmf, eq_eval)
            self._fold(r_m)
            q.append(r)
            remaining.pop()
        q.reverse()
        inner = self._layer_end()
        add_scalars(inner)
        beta = F.from_monty(sample())
        nl, nr, dl, dr = (F.from_monty(v) for v in inner)
        omb = F.sub(F.ONE, beta)
        return q + [beta], F.add(F.mul(omb, nl), F.mul(beta, nr)), F.add(F.mul(omb, dl), F.mul(beta, dr))

    def free(self):
        if self.handle:
            check(lib().lm_gkr_free(self.handle))
            self.handle = None


class GkrShardSession:
    """Device steps of the quotient GKR over ONE row-range shard of the fraction table (lm_gkr_new_shard): the compute
    backend of leanmultisig_b200.sharded.ShardedGkrQuotientProver on a GPU."""

    def __init__(self, ctx, nums, dens, n_vars: int, top_vars: int):
        n_, d_ = _u32(nums).reshape(-1), _u32(dens).reshape(-1, 5)
        assert n_.size == d_.shape[0] <= 1 << n_vars
        h = C.c_void_p()
        check(lib().lm_gkr_new_shard(ctx.handle, _p(n_) if n_.size else None, _p(d_) if n_.size else None, n_.size, n_vars,
                                     top_vars, C.byref(h)))
        self.handle, self.top_vars = h, top_vars

    def top(self):
        m = 1 << self.top_vars
        tn, td = np.empty((m, 5), dtype=np.uint32), np.empty((m, 5), dtype=np.uint32)
        check(lib().lm_gkr_top(self.handle, _p(tn), _p(td)))
        return tn, td

    def layer_begin(self, claim_vars: int, point_m, alpha_m, eq_scale_m) -> None:
        pt = _u32(point_m).reshape(-1, 5)
        check(lib().lm_gkr_layer_begin_shard(self.handle, claim_vars, _p(pt), _p(_u32(alpha_m)), _p(_u32(eq_scale_m))))

    round = GkrQuotientProver._round
    fold = GkrQuotientProver._fold
    layer_end = GkrQuotientProver._layer_end
    free = GkrQuotientProver.free


# ======================================================================================================
# prove_generic_logup (crates/sub_protocols/src/logup.rs:27-320)
# ======================================================================================================
class _Data(C.Structure):
    _fields_ = [("col", C.c_void_p), ("len", C.c_uint64), ("offset", C.c_uint64), ("stride", C.c_uint64),
                ("kind", C.c_uint32), ("value", C.c_uint32)]


NUM_ONE, NUM_COL, NUM_NEG_COL, NUM_ZERO = 0, 1, 2, 3
DATA_COL, DATA_ROW, DATA_CONST = 0, 1, 2


class LogupBuilder:
    """lm_logup_*: numerators / denominators of all sections, assembled on the device in natural row order."""

    def __init__(self, ctx, total_active_len: int, c, alphas_eq_poly):
        al = _u32(alphas_eq_poly).reshape(-1, 5)
        h = C.c_void_p()
        check(lib().lm_logup_new(ctx.handle, total_active_len, _p(_u32(c)), _p(al), al.shape[0], C.byref(h)))
        self.handle, self._keep = h, []

    def section(self, n_rows: int, num_mode: int, num_col, den_sign: int, domainsep: int, data) -> None:
        """data: list of ("col", array, offset, stride, add) | ("row",) | ("const", value)"""
        arr = (_Data * max(len(data), 1))()
        for i, d in enumerate(data):
            if d[0] == "col":
                a = d[1]
                assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
                self._keep.append(a)
                arr[i] = _Data(a.ctypes.data, a.size, d[2], d[3], DATA_COL, d[4])
            elif d[0] == "row":
                arr[i] = _Data(None, 0, 0, 1, DATA_ROW, 0)
            else:
                arr[i] = _Data(None, 0, 0, 1, DATA_CONST, d[1])
        ncol = None
        if num_col is not None:
            assert num_col.dtype == np.uint32 and num_col.flags["C_CONTIGUOUS"]
            self._keep.append(num_col)
            ncol = num_col.ctypes.data
        check(lib().lm_logup_section(self.handle, n_rows, num_mode, ncol, den_sign, domainsep, C.byref(arr), len(data)))

    def col_eval_batch(self, cols, n_vars: int, point) -> np.ndarray:
        """evaluations (len(cols) x 5) of several columns at ONE point: eq tables built once, one read-back"""
        pt = _u32(point).reshape(-1, 5)
        assert pt.shape[0] == n_vars
        out = np.empty((len(cols), 5), dtype=np.uint32)
        for k0 in range(0, len(cols), 256):
            part = cols[k0:k0 + 256]
            for c in part:
                assert c.dtype == np.uint32 and c.flags["C_CONTIGUOUS"]
            ptrs = (C.c_void_p * len(part))(*[c.ctypes.data for c in part])
            lens = np.array([c.size for c in part], dtype=np.uint64)
            check(lib().lm_logup_col_eval_batch(self.handle, ptrs, lens.ctypes.data_as(C.POINTER(C.c_uint64)), len(part), n_vars,
                                                _p(pt), _p(out[k0:k0 + len(part)])))
        return out

    def col_eval(self, col, n_vars: int, point) -> np.ndarray:
        assert col.dtype == np.uint32 and col.flags["C_CONTIGUOUS"]
        out = np.empty(5, dtype=np.uint32)
        pt = _u32(point).reshape(-1, 5)
        assert pt.shape[0] == n_vars
        check(lib().lm_logup_col_eval(self.handle, col.ctypes.data, col.size, n_vars, _p(pt), _p(out)))
        return out

    def read(self, n_rows: int):
        nums, dens = np.empty(n_rows, dtype=np.uint32), np.empty((n_rows, 5), dtype=np.uint32)
        check(lib().lm_logup_read(self.handle, nums.ctypes.data, dens.ctypes.data))
        return nums, dens

    def finish(self) -> GkrQuotientProver:
        h = C.c_void_p()
        check(lib().lm_logup_finish(self.handle, C.byref(h)))
        return GkrQuotientProver.from_handle(h)

    def free(self):
        if self.handle:
            check(lib().lm_logup_free(self.handle))
            self.handle = None


def build_logup_table(ctx, c, alphas_eq_poly, memory, memory_acc, bytecode_multilinear, bytecode_acc, traces) -> LogupBuilder:
    """The section list of logup.rs:96-211.  traces: {tables.Table: tables.TableTrace}"""
    from . import tables as T

    log_memory = memory.size.bit_length() - 1
    assert memory.size == 1 << log_memory and memory_acc.size == memory.size
    stride = 1 << (T.N_INSTRUCTION_COLUMNS - 1).bit_length()
    log_bytecode = (bytecode_multilinear.size // stride).bit_length() - 1
    assert bytecode_acc.size == 1 << log_bytecode
    tables_sorted = T.sort_tables_by_height({t: tr.log_n_rows for t, tr in traces.items()})
    assert memory.size >= 1 << tables_sorted[0][1]
    total = T.compute_total_active_len(log_memory, log_bytecode, tables_sorted)
    b = LogupBuilder(ctx, total, c, alphas_eq_poly)
    max_table_height = 1 << tables_sorted[0][1]
    # memory
    b.section(memory.size, NUM_NEG_COL, memory_acc, -1, T.LOGUP_MEMORY_DOMAINSEP, [("col", memory, 0, 1, 0), ("row",)])
    # bytecode (+ padding up to the tallest table)
    data = [("col", bytecode_multilinear, k, stride, 0) for k in range(T.N_INSTRUCTION_COLUMNS)] + [("row",)]
    b.section(1 << log_bytecode, NUM_NEG_COL, bytecode_acc, -1, T.LOGUP_BYTECODE_DOMAINSEP, data)
    if (1 << log_bytecode) < max_table_height:
        b.section(max_table_height - (1 << log_bytecode), NUM_ZERO, None, 0, 0, [])
    for table, log_n_rows in tables_sorted:
        cols = traces[table].columns
        n = 1 << log_n_rows
        if table.is_execution:
            data = [("col", cols[T.N_RUNTIME_COLUMNS + k], 0, 1, 0) for k in range(T.N_INSTRUCTION_COLUMNS)]
            data.append(("col", cols[T.COL_PC], 0, 1, 0))
            b.section(n, NUM_ONE, None, -1, T.LOGUP_BYTECODE_DOMAINSEP, data)
        bus = table.bus
        b.section(n, NUM_NEG_COL if bus.pull else NUM_COL, cols[bus.selector], +1, T.LOGUP_PRECOMPILE_DOMAINSEP,
                  [("col", cols[k], 0, 1, 0) for k in bus.data])
        for lk in table.lookups:
            for i, vcol in enumerate(lk.values):
                b.section(n, NUM_ONE, None, -1, T.LOGUP_MEMORY_DOMAINSEP,
                          [("col", cols[vcol], 0, 1, 0), ("col", cols[lk.index], 0, 1, i)])
    return b


def prove_generic_logup(ctx, prover_state, c, alphas_eq_poly, memory, memory_acc, bytecode_multilinear, bytecode_acc, traces,
                        timings: dict | None = None):
    """logup.rs:27-320 -> dict mirroring GenericLogupStatements (Montgomery arrays).  `timings`, when given, receives the
    wall-clock seconds of the three phases (table assembly, GKR, column evaluations)."""
    import time

    from . import tables as T

    t_start = time.perf_counter()
    al = _u32(alphas_eq_poly).reshape(-1, 5)
    b = build_logup_table(ctx, c, al, memory, memory_acc, bytecode_multilinear, bytecode_acc, traces)
    gkr = b.finish()
    if timings is not None:
        ctx.sync()
        timings["table assembly + GKR up pass"] = time.perf_counter() - t_start
        t_start = time.perf_counter()
    total_gkr_n_vars = gkr.n_vars
    # a NativeProverState (C++ transcript) runs the whole GKR in the library's spine, anything else round by round here
    quotient, point, _, _ = gkr.prove_native(prover_state) if hasattr(prover_state, "handle") else gkr.prove_with_state(prover_state)
    gkr.free()
    if timings is not None:
        timings["GKR down pass"] = time.perf_counter() - t_start
        t_start = time.perf_counter()
    assert not quotient.any(), "logup sum is not zero"
    log_memory = memory.size.bit_length() - 1
    stride = 1 << (T.N_INSTRUCTION_COLUMNS - 1).bit_length()
    log_bytecode = (bytecode_multilinear.size // stride).bit_length() - 1

    def from_end(k):
        return point[point.shape[0] - k:]

    def add(v):
        prover_state.add_extension_scalars(v)
        return v

    out = dict(gkr_point=point, total_gkr_n_vars=total_gkr_n_vars)
    out["memory_and_acc_point"] = from_end(log_memory)
    mem_evals = b.col_eval_batch([memory_acc, memory], log_memory, from_end(log_memory))
    out["value_memory_acc"] = add(mem_evals[0])
    out["value_memory"] = add(mem_evals[1])
    out["bytecode_and_acc_point"] = from_end(log_bytecode)
    out["value_bytecode_acc"] = add(b.col_eval(bytecode_acc, log_bytecode, from_end(log_bytecode)))
    out["bus_numerators_values"], out["bus_denominators_values"], out["columns_values"] = {}, {}, {}
    c_c = F.from_monty(c)
    al_c = [F.from_monty(a) for a in al]
    for table, log_n_rows in T.sort_tables_by_height({t: tr.log_n_rows for t, tr in traces.items()}):
        cols = traces[table].columns
        inner = from_end(log_n_rows)
        # every column of this table that logup.rs:224-305 evaluates, at the one inner point, in one batch
        needed = [bus_col for bus_col in (table.bus.selector, *table.bus.data)]
        if table.is_execution:
            needed += [T.COL_PC] + [T.N_RUNTIME_COLUMNS + k for k in range(T.N_INSTRUCTION_COLUMNS)]
        for lk in table.lookups:
            needed += [lk.index, *lk.values]
        needed = sorted(set(needed))
        batch = b.col_eval_batch([cols[k] for k in needed], log_n_rows, inner)
        ev = {k: batch[i] for i, k in enumerate(needed)}
        values = {}
        if table.is_execution:
            values[T.COL_PC] = add(ev[T.COL_PC])
            instr = [ev[T.N_RUNTIME_COLUMNS + k] for k in range(T.N_INSTRUCTION_COLUMNS)]
            prover_state.add_extension_scalars(np.concatenate(instr))
            for k, v in enumerate(instr):
                values[T.N_RUNTIME_COLUMNS + k] = v
        bus = table.bus
        sel = F.from_monty(ev[bus.selector])
        if bus.pull:
            sel = F.neg(sel)
        out["bus_numerators_values"][table] = add(F.to_monty(sel))
        data_evals = [F.from_monty(ev[k]) for k in bus.data]
        fp = F.scal(al_c[-1], T.LOGUP_PRECOMPILE_DOMAINSEP)
        for a, d in zip(al_c, data_evals):
            fp = F.add(fp, F.mul(a, d))
        out["bus_denominators_values"][table] = add(F.to_monty(F.add(c_c, fp)))
        for lk in table.lookups:
            values[lk.index] = add(ev[lk.index])
            for vcol in lk.values:
                values[vcol] = add(ev[vcol])
        out["columns_values"][table] = values
    b.free()
    if timings is not None:
        timings["column evaluations"] = time.perf_counter() - t_start
    return out

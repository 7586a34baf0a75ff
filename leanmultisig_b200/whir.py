"""Host-side mirror of the reference's WHIR commit surface on top of the C ABI.

Reference call sites mirrored (names kept):
  WhirConfig::commit      crates/whir/src/commit.rs:64-99   -> Context.commit
  MerkleData::open        crates/whir/src/commit.rs:34-45   -> Tree.open
  MleRef::evaluate        crates/backend/poly/src/evals.rs  -> Tree.evaluate / Context.mle_eval
  reorder_and_dft         crates/whir/src/utils.rs:69       -> Context.reorder_and_dft (device buffers)
  build_merkle_tree_*     crates/whir/src/merkle.rs:59      -> Context.merkle_tree (device buffers)
All field elements are numpy uint32 in Montgomery form, exactly as the reference stores them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import check, lib, u32p, u64p


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(u32p)


class DeviceBuffer:
    """A cudaMalloc'd buffer owned by a Context (inputs resident in HBM for kernel-only timing)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, nbytes
        p = C.c_void_p()
        check(lib().lm_dev_alloc(ctx.handle, nbytes, C.byref(p)))
        self.ptr = p

    def upload(self, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(lib().lm_dev_upload(self.ctx.handle, self.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return self

    def download(self, shape, dtype=np.uint32) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().lm_dev_download(self.ctx.handle, out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            check(lib().lm_dev_free(self.ctx.handle, self.ptr))
            self.ptr = None


class Tree:
    """Prover data of one commitment (reference: MerkleData + Witness, crates/whir/src/commit.rs:11-57)."""

    def __init__(self, ctx: "Context", handle, root: np.ndarray):
        self.ctx, self.handle, self.root = ctx, handle, root
        h, fw, sw, dim = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().lm_tree_shape(handle, C.byref(h), C.byref(fw), C.byref(sw), C.byref(dim)))
        self.height, self.full_width, self.stored_width, self.elem_dim = h.value, fw.value, sw.value, dim.value
        self.log_height = self.height.bit_length() - 1

    def open(self, indices):
        """(rows zero-extended to full width, sibling paths) for the given leaf indices."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        n = idx.size
        rows = np.empty((n, self.full_width), dtype=np.uint32)
        paths = np.empty((n, self.log_height, 8), dtype=np.uint32)
        check(lib().lm_open(self.handle, idx.ctypes.data_as(u64p), n, _p(rows), _p(paths)))
        return rows, paths

    def open_fold(self, indices, fold_point):
        """open() plus the STIR answer of every opened leaf: its multilinear evaluation at `fold_point` (lm_open_fold)"""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        pt = _u32(fold_point).reshape(-1, 5)
        n = idx.size
        rows = np.empty((n, self.full_width), dtype=np.uint32)
        paths = np.empty((n, self.log_height, 8), dtype=np.uint32)
        evals = np.empty((n, 5), dtype=np.uint32)
        check(lib().lm_open_fold(self.handle, idx.ctypes.data_as(u64p), n, _p(pt), pt.shape[0], _p(rows), _p(paths), _p(evals)))
        return rows, paths, evals

    def evaluate(self, point) -> np.ndarray:
        pt = _u32(point).reshape(-1, 5)
        out = np.empty(5, dtype=np.uint32)
        check(lib().lm_tree_eval(self.handle, _p(pt), _p(out)))
        return out

    def codeword(self) -> np.ndarray:
        out = np.empty((self.height, self.stored_width), dtype=np.uint32)
        check(lib().lm_tree_read_codeword(self.handle, _p(out)))
        return out

    def layers(self) -> np.ndarray:
        out = np.empty((2 * self.height - 1, 8), dtype=np.uint32)
        check(lib().lm_tree_read_layers(self.handle, _p(out)))
        return out

    def free(self):
        if self.handle:
            check(lib().lm_tree_free(self.handle))
            self.handle = None


class ProductSumcheck:
    """WHIR-open sumcheck session (reference: SumcheckSingle, crates/whir/src/open.rs:323-446).

    Holds the polynomial table and the weight table on the device; the caller owns the transcript and drives
    one call per round, exactly like run_product_sumcheck / sumcheck_prove_many_rounds do on the CPU."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.handle = ctx, handle

    @property
    def n_vars(self) -> int:
        n, d = C.c_uint32(), C.c_uint32()
        check(lib().lm_sc_num_vars(self.handle, C.byref(n), C.byref(d)))
        return n.value

    @property
    def poly_dim(self) -> int:
        n, d = C.c_uint32(), C.c_uint32()
        check(lib().lm_sc_num_vars(self.handle, C.byref(n), C.byref(d)))
        return d.value

    def add_eq(self, selector: int, point, scalar):
        pt = _u32(point).reshape(-1, 5)
        check(lib().lm_sc_add_eq(self.handle, selector, _p(pt), pt.shape[0], _p(_u32(scalar))))

    def add_eq_batch(self, selector: int, points, scalars):
        """several eq statements with the same selector and point length in one pass over the weights (lm_sc_add_eq_batch)"""
        pts = _u32(points)
        assert pts.ndim == 3 and pts.shape[2] == 5
        sc = _u32(scalars).reshape(-1, 5)
        assert sc.shape[0] == pts.shape[0]
        check(lib().lm_sc_add_eq_batch(self.handle, selector, _p(pts), pts.shape[1], _p(sc), pts.shape[0]))

    def add_next(self, selector: int, point, scalar):
        pt = _u32(point).reshape(-1, 5)
        check(lib().lm_sc_add_next(self.handle, selector, _p(pt), pt.shape[0], _p(_u32(scalar))))

    def add_strided_eq(self, base: int, shift: int, offset: int, point, scalar):
        """w[base + (b << shift) + offset] += scalar * eq(point, b), b < 2^len(point)  (lm_sc_add_strided_eq)"""
        pt = _u32(point).reshape(-1, 5)
        check(lib().lm_sc_add_strided_eq(self.handle, base, shift, offset, _p(pt) if pt.size else None, pt.shape[0], _p(_u32(scalar))))

    def add_base_eq(self, points, scalars):
        pts, sc = _u32(points), _u32(scalars).reshape(-1, 5)
        check(lib().lm_sc_add_base_eq(self.handle, _p(pts), pts.shape[0], _p(sc)))

    def stir_update(self, idx, gen_monty: int, comb, ood_ys, ood_answers, stir_evals, total):
        """the weight update closing a WHIR round (lm_whir_stir_update); returns the new running sum (Montgomery words)"""
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        oy = _u32(ood_ys).reshape(-1, 5)
        oa = _u32(ood_answers).reshape(-1, 5)
        se = _u32(stir_evals).reshape(-1, 5)
        tot = _u32(total).copy()
        check(lib().lm_whir_stir_update(self.handle, idx.ctypes.data_as(u64p), idx.size, int(gen_monty), self.n_vars, _p(_u32(comb)),
                                        _p(oy) if oy.size else None, _p(oa) if oa.size else None, oy.shape[0],
                                        _p(se) if se.size else None, _p(tot)))
        return tot

    def round(self):
        c0, c2 = np.empty(5, dtype=np.uint32), np.empty(5, dtype=np.uint32)
        check(lib().lm_sc_round(self.handle, _p(c0), _p(c2)))
        return c0, c2

    def fold(self, r):
        check(lib().lm_sc_fold(self.handle, _p(_u32(r))))

    def fold_round(self, r):
        c0, c2 = np.empty(5, dtype=np.uint32), np.empty(5, dtype=np.uint32)
        check(lib().lm_sc_fold_round(self.handle, _p(_u32(r)), _p(c0), _p(c2)))
        return c0, c2

    def read(self):
        n, d = self.n_vars, self.poly_dim
        poly = np.empty((1 << n, d) if d == 5 else (1 << n,), dtype=np.uint32)
        w = np.empty((1 << n, 5), dtype=np.uint32)
        check(lib().lm_sc_read(self.handle, _p(poly), _p(w)))
        return poly, w

    def eval_poly(self, point) -> np.ndarray:
        out = np.empty(5, dtype=np.uint32)
        check(lib().lm_sc_eval_poly(self.handle, _p(_u32(point).reshape(-1, 5)), _p(out)))
        return out

    def export_dev(self, d_poly_ptr: int, d_weights_ptr: int) -> None:
        """copy the current EF tables into caller-owned device buffers (row-sharded sumcheck, sharded.py)"""
        check(lib().lm_sc_export_dev(self.handle, C.c_void_p(d_poly_ptr), C.c_void_p(d_weights_ptr)))

    def commit_poly(self, folding_factor: int, log_inv_rate: int) -> Tree:
        root = np.empty(8, dtype=np.uint32)
        t = C.c_void_p()
        check(lib().lm_sc_commit_poly(self.handle, folding_factor, log_inv_rate, C.byref(t), _p(root)))
        return Tree(self.ctx, t, root)

    def free(self):
        if self.handle:
            check(lib().lm_sc_free(self.handle))
            self.handle = None


class Context:
    """One GPU: stream + twiddle table (reference: setup_prover / precompute_dft_twiddles)."""

    def __init__(self, device: int = 0, max_log_domain: int = 24):
        h = C.c_void_p()
        check(lib().lm_init(device, max_log_domain, C.byref(h)))
        self.handle, self.device = h, device

    def close(self):
        if self.handle:
            check(lib().lm_destroy(self.handle))
            self.handle = None

    def set_stream(self, cuda_stream: int | None):
        check(lib().lm_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        check(lib().lm_sync(self.handle))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, a: np.ndarray) -> DeviceBuffer:
        a = np.ascontiguousarray(a)
        return DeviceBuffer(self, max(a.nbytes, 4)).upload(a)

    # ---- host-buffer API (the drop-in boundary) -------------------------------------------------------
    def commit(self, evals, n_vars: int, folding_factor: int, log_inv_rate: int, actual_len: int | None = None) -> Tree:
        e = _u32(evals)
        dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
        n_elems = e.size // dim
        if actual_len is None:
            actual_len = n_elems
        assert n_elems >= actual_len
        root = np.empty(8, dtype=np.uint32)
        t = C.c_void_p()
        check(lib().lm_commit(self.handle, e.ctypes.data_as(C.c_void_p), n_vars, dim, actual_len, folding_factor,
                              log_inv_rate, C.byref(t), _p(root)))
        return Tree(self, t, root)

    def commit_dev(self, d_evals: DeviceBuffer, n_vars: int, dim: int, folding_factor: int, log_inv_rate: int,
                   actual_len: int, retain_evals: bool = False) -> Tree:
        root = np.empty(8, dtype=np.uint32)
        t = C.c_void_p()
        check(lib().lm_commit_dev(self.handle, d_evals.ptr, n_vars, dim, actual_len, folding_factor, log_inv_rate,
                                  1 if retain_evals else 0, C.byref(t), _p(root)))
        return Tree(self, t, root)

    def mle_eval(self, evals, point, live_len: int | None = None) -> np.ndarray:
        e = _u32(evals)
        pt = _u32(point).reshape(-1, 5)
        dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
        n = pt.shape[0]
        out = np.empty(5, dtype=np.uint32)
        live = (e.size // dim) if live_len is None else live_len
        check(lib().lm_mle_eval(self.handle, e.ctypes.data_as(C.c_void_p), n, dim, live, _p(pt), _p(out)))
        return out

    def sumcheck_from_tree(self, tree: Tree) -> ProductSumcheck:
        h = C.c_void_p()
        check(lib().lm_sc_new_from_tree(tree.handle, C.byref(h)))
        return ProductSumcheck(self, h)

    def sumcheck(self, evals, n_vars: int, live_len: int | None = None) -> ProductSumcheck:
        e = _u32(evals)
        dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
        live = (e.size // dim) if live_len is None else live_len
        h = C.c_void_p()
        check(lib().lm_sc_new(self.handle, e.ctypes.data_as(C.c_void_p), n_vars, dim, live, C.byref(h)))
        return ProductSumcheck(self, h)

    def sumcheck_from_dev(self, d_poly_ptr: int, d_weights_ptr: int, n_vars: int) -> ProductSumcheck:
        h = C.c_void_p()
        check(lib().lm_sc_new_from_dev(self.handle, C.c_void_p(d_poly_ptr), C.c_void_p(d_weights_ptr), n_vars, C.byref(h)))
        return ProductSumcheck(self, h)

    # ---- device-buffer API ------------------------------------------------------------------------------
    def poseidon1(self, states, compress: bool = False) -> np.ndarray:
        s = _u32(states).reshape(-1, 16)
        d = self.to_device(s)
        check(lib().lm_dev_poseidon1(self.handle, d.ptr, s.shape[0], 1 if compress else 0))
        out = d.download(s.shape)
        d.free()
        return out.reshape(np.shape(states))

    def reorder_and_dft(self, evals, n_vars: int, folding: int, log_inv_rate: int, dft_n_cols: int) -> np.ndarray:
        e = _u32(evals)
        dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
        h = 1 << (n_vars + log_inv_rate - folding)
        d_in = self.to_device(e)
        d_out = self.alloc(h * dft_n_cols * dim * 4)
        check(lib().lm_dev_reorder_and_dft(self.handle, d_in.ptr, n_vars, dim, folding, log_inv_rate, dft_n_cols, d_out.ptr))
        out = d_out.download((h, dft_n_cols * dim))
        d_in.free(), d_out.free()
        return out

    def dft_batch_by_evals(self, mat) -> np.ndarray:
        m = _u32(mat)
        d = self.to_device(m)
        check(lib().lm_dev_dft(self.handle, d.ptr, m.shape[0], m.shape[1]))
        out = d.download(m.shape)
        d.free()
        return out

    def merkle_tree(self, mat, full_width: int, effective_width: int) -> np.ndarray:
        m = _u32(mat)
        h, w = m.shape
        d = self.to_device(m)
        d_layers = self.alloc((2 * h - 1) * 32)
        check(lib().lm_dev_merkle_tree(self.handle, d.ptr, h, w, full_width, effective_width, d_layers.ptr))
        out = d_layers.download((2 * h - 1, 8))
        d.free(), d_layers.free()
        return out

    def verify_openings(self, root, log_height: int, indices, rows, paths, elem_dim: int = 1, fold_point=None):
        """Verifier side of Tree.open / open_fold (verify.rs:229-345): ok[q] = opening q hashes to `root`; with `fold_point`
        also the fold of every leaf at the round's folding randomness.  One launch for all openings of a round."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        n = idx.size
        r = _u32(rows).reshape(n, -1) if n else _u32(rows).reshape(0, 16)
        pth = _u32(paths).reshape(n, log_height, 8)
        ok = np.zeros(n, dtype=np.uint8)
        pt = evals = None
        if fold_point is not None:
            pt = _u32(fold_point).reshape(-1, 5)
            evals = np.empty((n, 5), dtype=np.uint32)
        check(lib().lm_verify_openings(self.handle, _p(_u32(root)), log_height, idx.ctypes.data_as(u64p), n, _p(r), r.shape[1],
                                       elem_dim, _p(pth), _p(pt) if pt is not None and pt.size else None,
                                       pt.shape[0] if pt is not None else 0, ok.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       _p(evals) if evals is not None else None))
        return (ok.astype(bool), evals) if fold_point is not None else ok.astype(bool)

    def fold_msb(self, evals, r) -> np.ndarray:
        e, r = _u32(evals), _u32(r)
        dim = 5 if (e.ndim == 2 and e.shape[1] == 5) else 1
        n = e.size // dim
        d = self.to_device(e)
        d_out = self.alloc(max(n // 2, 1) * 20)
        check(lib().lm_dev_fold_msb(self.handle, d.ptr, n, dim, _p(r), d_out.ptr))
        out = d_out.download((n // 2, 5))
        d.free(), d_out.free()
        return out

    def eq_table(self, point, scalar=None) -> np.ndarray:
        pt = _u32(point).reshape(-1, 5)
        k = pt.shape[0]
        sc = _u32(scalar) if scalar is not None else np.array([0x01FFFFFE, 0, 0, 0, 0], dtype=np.uint32)
        d_out = self.alloc((1 << k) * 20)
        check(lib().lm_dev_eq_table(self.handle, _p(pt), k, _p(sc), d_out.ptr))
        out = d_out.download((1 << k, 5))
        d_out.free()
        return out


# ======================================================================================================
# WhirConfig::commit / WhirConfig::prove on the device sessions above
# ======================================================================================================
class SparseStatement:
    """crates/whir/src/lib.rs:31-95.  point: m x 5 words over the low (inner) variables; values: [(selector, 5 words)]:
    the claim is  poly(selector bits | point) = value  (or, with is_next, the next_mle-shifted claim)."""

    def __init__(self, total_num_variables: int, point, values, is_next: bool = False):
        self.total_num_variables = total_num_variables
        self.point = _u32(point).reshape(-1, 5)
        self.values = [(int(s), _u32(v).reshape(5)) for s, v in values]
        self.is_next = is_next
        assert self.point.shape[0] <= total_num_variables

    @staticmethod
    def dense(point, value) -> "SparseStatement":
        pt = _u32(point).reshape(-1, 5)
        return SparseStatement(pt.shape[0], pt, [(0, value)])


class Witness:
    """crates/whir/src/commit.rs:49-57: prover data + the out-of-domain samples taken at commit time"""

    def __init__(self, tree: Tree, ood_points, ood_answers):
        self.tree, self.ood_points, self.ood_answers = tree, ood_points, ood_answers

    def free(self):
        self.tree.free()


def _expand_from_univariate(y, n: int):
    """MultilinearPoint::expand_from_univariate (crates/backend/poly/src/point.rs): y, y^2, y^4, ... (canonical)"""
    from . import field as F

    out, cur = [], y
    for _ in range(n):
        out.append(cur)
        cur = F.mul(cur, cur)
    return out


def _points_to_monty(pts) -> np.ndarray:
    from . import field as F

    return np.stack([F.to_monty(x) for x in pts]) if len(pts) else np.zeros((0, 5), dtype=np.uint32)


def _sample_ood(prover_state, n_samples: int, n_vars: int, evaluate):
    """sample_ood_points (crates/whir/src/utils.rs:30-57)"""
    from . import field as F

    pts, answers = [], []
    if n_samples:
        pts = [F.from_monty(x) for x in prover_state.sample_vec(n_samples)]
        answers = [_u32(evaluate(_points_to_monty(_expand_from_univariate(y, n_vars)))) for y in pts]
        prover_state.add_extension_scalars(np.concatenate(answers))
    return pts, answers


def evals_to_coeffs(evals: np.ndarray) -> np.ndarray:
    """crates/backend/poly/src/evals.rs:44-56 on the (at most 2^max_num_variables_to_send_coeffs) final values"""
    from . import field as F

    d = F.np_from_monty(_u32(evals).reshape(-1, 5))
    n = d.shape[0]
    p = np.uint64(F.P)
    half = 1
    while half < n:
        v = d.reshape(n // (2 * half), 2, half, 5)
        v[:, 1] = (v[:, 1] + (p - v[:, 0])) % p
        half *= 2
    log_n = n.bit_length() - 1
    rev = [int(format(i, f"0{log_n}b")[::-1], 2) if log_n else 0 for i in range(n)]
    return F.np_to_monty(d[rev])


class WhirProver:
    """`WhirConfig::commit` (commit.rs:64-99) and `WhirConfig::prove` (open.rs:37-248) with every table on the GPU.

    cfg: a whir_config.WhirConfig (or the reference's, through a binding exposing the same fields);
    prover_state: anything with the FSProver surface of fiat_shamir.ProverState."""

    def __init__(self, ctx: Context, cfg):
        self.ctx, self.cfg = ctx, cfg

    def commit(self, prover_state, polynomial, actual_len: int | None = None) -> Witness:
        cfg = self.cfg
        tree = self.ctx.commit(polynomial, cfg.num_variables, cfg.first_folding, cfg.starting_log_inv_rate, actual_len)
        prover_state.add_base_scalars(tree.root)
        pts, answers = _sample_ood(prover_state, cfg.commitment_ood_samples, cfg.num_variables, tree.evaluate)
        return Witness(tree, pts, answers)

    @staticmethod
    def _open_fold(tree, idx, fold_point):
        """(rows, paths, STIR answers) of the queried leaves; trees without a device fold (the sharded first tree) fold on the host"""
        if hasattr(tree, "open_fold"):
            return tree.open_fold(idx, fold_point)
        from . import field as F

        rows, paths = tree.open(idx)
        ff = fold_point.shape[0]
        leaves = F.np_from_monty(rows)
        if tree.elem_dim == 5:
            leaves = leaves.reshape(len(idx), 1 << ff, 5)
        else:
            z = np.zeros((len(idx), 1 << ff, 5), dtype=np.uint64)
            z[:, :, 0] = leaves
            leaves = z
        ev = F.np_mle_eval_rows(leaves, [F.from_monty(x) for x in fold_point])
        return rows, paths, F.np_to_monty(ev)

    @staticmethod
    def _stir_update(sc, idx, gen, comb_m, ood_points, ood_answers, stir_evals, total, num_variables):
        """open.rs:192-231; sessions without the C++ entry (compute doubles of the CPU test tier) take the host path"""
        from . import field as F

        if getattr(sc, "stir_update", None) is not None and not getattr(sc, "host_stir_update", False):
            oa = np.array(ood_answers, dtype=np.uint32).reshape(-1, 5)
            return F.from_monty(sc.stir_update(idx, gen, comb_m, _points_to_monty(ood_points), oa, stir_evals, F.to_monty(total)))
        comb = F.from_monty(comb_m)
        g = gen * F._RINV % F.P
        p64, r64 = np.uint64(F.P), np.uint64(F._R)
        y = np.array([pow(g, int(i), F.P) for i in idx], dtype=np.uint64)      # canonical, < 2^31: products fit in u64
        stir_pts = np.empty((len(idx), num_variables), dtype=np.uint32)
        for k in range(num_variables):
            stir_pts[:, k] = (y * r64) % p64
            y = (y * y) % p64
        powers = [F.ONE]
        for _ in range(len(ood_points) + len(idx)):
            powers.append(F.mul(powers[-1], comb))
        for k, (yk, ans) in enumerate(zip(ood_points, ood_answers)):
            sc.add_eq(0, _points_to_monty(_expand_from_univariate(yk, num_variables)), F.to_monty(powers[k]))
            total = F.add(total, F.mul(powers[k], F.from_monty(ans)))
        stir_rand = powers[len(ood_points):len(ood_points) + len(idx)]
        if len(idx):
            sc.add_base_eq(stir_pts, _points_to_monty(stir_rand))
        for rnd, ev in zip(stir_rand, F.np_from_monty(_u32(stir_evals).reshape(-1, 5))):
            total = F.add(total, F.mul(rnd, tuple(int(x) for x in ev)))
        return total

    def _session(self, witness: Witness):
        """the product-sumcheck session over the committed polynomial (the sharded prover overrides this)"""
        return self.ctx.sumcheck_from_tree(witness.tree)

    # -- sumcheck_prove_many_rounds with the product-sumcheck kernels (sumcheck/src/prove.rs:86-151) ------
    @staticmethod
    def _rounds(sc: ProductSumcheck, prover_state, n_rounds: int, pow_bits: int, total):
        from . import field as F
        from .fiat_shamir import NativeProverState

        if (type(sc) is ProductSumcheck and isinstance(prover_state, NativeProverState) and n_rounds
                and not os.environ.get("LM_WHIR_PY_ROUNDS")):
            # the whole phase in the library's C++ spine (lm_whir_sumcheck_rounds): same transcript, no per-round return to Python
            tot = np.ascontiguousarray(F.to_monty(total), dtype=np.uint32)
            out = np.empty((n_rounds, 5), dtype=np.uint32)
            check(lib().lm_whir_sumcheck_rounds(sc.handle, prover_state.handle, n_rounds, pow_bits, _p(tot), _p(out)))
            return [F.from_monty(r) for r in out], F.from_monty(tot)
        chals, pending = [], None
        for _ in range(n_rounds):
            c0, c2 = sc.round() if pending is None else sc.fold_round(pending)
            c0c, c2c = F.from_monty(c0), F.from_monty(c2)
            c1c = F.sub(F.sub(total, F.add(c0c, c0c)), c2c)
            prover_state.add_sumcheck_polynomial(np.stack([c0, F.to_monty(c1c), c2]))
            prover_state.pow_grinding(pow_bits)
            pending = prover_state.sample()
            r = F.from_monty(pending)
            total = F.add(c0c, F.mul(r, F.add(c1c, F.mul(r, c2c))))
            chals.append(r)
        if pending is not None:
            sc.fold(pending)
        return chals, total

    def prove(self, prover_state, statements, witness: Witness):
        from . import field as F
        from .whir_config import two_adic_generator

        cfg, ps = self.cfg, prover_state
        nv = cfg.num_variables
        for s in statements:
            assert s.total_num_variables == nv
        # combine_statement (open.rs:518-584): OOD constraints first, then the caller's
        stm = [SparseStatement.dense(_points_to_monty(_expand_from_univariate(y, nv)), a)
               for y, a in zip(witness.ood_points, witness.ood_answers)] + list(statements)
        import os
        import time

        tm = {} if os.environ.get("LM_WHIR_TIMING") else None
        t_last = [time.perf_counter()]

        def lap(name):  # wall-clock phase accounting (LM_WHIR_TIMING=1), printed at the end
            if tm is not None:
                if self.ctx is not None:
                    self.ctx.sync()
                now = time.perf_counter()
                tm[name] = tm.get(name, 0.0) + now - t_last[0]
                t_last[0] = now

        ps.duplex()
        gamma = F.from_monty(ps.sample())
        sc = self._session(witness)
        lap("session (weights alloc)")
        total, gp = F.ZERO, F.ONE
        groups = {}  # eq statements by (selector, point length): added in one pass over the weights each
        for smt in stm:
            for sel, val in smt.values:
                if smt.is_next or smt.point.shape[0] == 0 or not hasattr(sc, "add_eq_batch"):
                    (sc.add_next if smt.is_next else sc.add_eq)(sel, smt.point, F.to_monty(gp))
                else:
                    groups.setdefault((sel, smt.point.shape[0]), []).append((smt.point, F.to_monty(gp)))
                total = F.add(total, F.mul(F.from_monty(val), gp))
                gp = F.mul(gp, gamma)
        for (sel, m), items in groups.items():
            if len(items) == 1:
                sc.add_eq(sel, items[0][0], items[0][1])
            else:
                sc.add_eq_batch(sel, np.stack([p for p, _ in items]), np.stack([s for _, s in items]))
        lap("combine_statement")
        randomness, total = self._rounds(sc, ps, cfg.first_folding, cfg.starting_folding_pow_bits, total)
        lap("sumcheck rounds (incl. PoW)")

        trees = []
        tree = witness.tree
        domain_size = cfg.starting_domain_size()
        gen = two_adic_generator(domain_size.bit_length() - 1 - cfg.first_folding)
        for round_index in range(cfg.n_rounds + 1):
            num_variables = nv - cfg.total_folding(round_index)
            ff = cfg.folding_at(round_index)
            log_folded = (domain_size >> ff).bit_length() - 1
            if round_index == cfg.n_rounds:  # final_round (open.rs:251-321)
                poly, _ = sc.read()
                ps.add_extension_scalars(evals_to_coeffs(poly).reshape(-1))
                ps.pow_grinding(cfg.final_query_pow_bits)
                idx = ps.sample_in_range(log_folded, cfg.final_queries)
                rows, paths = tree.open(idx)
                ps.hint_merkle_paths([(rows[q], paths[q], i) for q, i in enumerate(idx)])
                if cfg.final_sumcheck_rounds:
                    r, total = self._rounds(sc, ps, cfg.final_sumcheck_rounds, 0, total)
                    randomness += r
                lap("final round")
                break
            rp = cfg.round_parameters[round_index]
            ff_next = cfg.folding_at(round_index + 1)
            new_domain_size = domain_size >> cfg.rs_reduction_factor(round_index)
            log_inv_rate = (new_domain_size >> num_variables).bit_length() - 1
            new_tree = sc.commit_poly(ff_next, log_inv_rate)
            trees.append(new_tree)
            ps.add_base_scalars(new_tree.root)
            lap("round commit")
            ood_points, ood_answers = _sample_ood(ps, rp.ood_samples, num_variables, sc.eval_poly)
            ps.pow_grinding(rp.query_pow_bits)
            lap("ood + query PoW")
            idx = ps.sample_in_range(log_folded, rp.num_queries)
            # openings + STIR answers: each opened leaf folded at this round's challenges, on the device (open.rs:161-190)
            fold_pt = _points_to_monty(randomness[len(randomness) - ff:])
            rows, paths, stir_evals = self._open_fold(tree, idx, fold_pt)
            ps.hint_merkle_paths([(rows[q], paths[q], i) for q, i in enumerate(idx)])
            lap("openings")
            ps.duplex()
            comb = ps.sample()
            # combination randomness, OOD + in-domain constraints into the weights, running sum (open.rs:192-231) in C++
            total = self._stir_update(sc, idx, gen, comb, ood_points, ood_answers, stir_evals, total, num_variables)
            lap("STIR answers + weights update")
            r, total = self._rounds(sc, ps, ff_next, rp.folding_pow_bits, total)
            lap("sumcheck rounds (incl. PoW)")
            randomness += r
            domain_size = new_domain_size
            gen = two_adic_generator(new_domain_size.bit_length() - 1 - ff_next)
            tree = new_tree
        sc.free()
        for t in trees:
            t.free()
        lap("free")
        if tm is not None:
            print("    WHIR open phases (ms):", {k: round(v * 1e3, 1) for k, v in tm.items()})
        return randomness

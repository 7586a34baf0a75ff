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

    def add_next(self, selector: int, point, scalar):
        pt = _u32(point).reshape(-1, 5)
        check(lib().lm_sc_add_next(self.handle, selector, _p(pt), pt.shape[0], _p(_u32(scalar))))

    def add_base_eq(self, points, scalars):
        pts, sc = _u32(points), _u32(scalars).reshape(-1, 5)
        check(lib().lm_sc_add_base_eq(self.handle, _p(pts), pts.shape[0], _p(sc)))

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

"""Worker for tests/test_sharded_sumcheck.py (torch.distributed.run; gloo on CPU or nccl on GPUs).

Every rank builds the same stacked polynomial and statements from a seed, keeps its shard, and drives
leanmultisig_b200.sharded.ShardedProductSumcheck; every round's (c0, c2) and the replicated tables after the local folds
must equal the single-process oracle session (oracle weights_add_eq / prod_round / fold_msb) over the whole polynomial."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from leanmultisig_b200.sharded import ShardedProductSumcheck, shard_of  # noqa: E402
from _sharded_air_worker import OracleBackend as _AirOracleBackend  # noqa: E402


class OracleSumcheck:
    """single-process product sumcheck on the oracle (reference AND per-rank compute double)"""

    def __init__(self, poly, weights=None, n_vars=None):
        self.p = np.ascontiguousarray(poly, dtype=np.uint32)
        n = self.p.shape[0]
        self.w = np.zeros((n, 5), dtype=np.uint32) if weights is None else np.ascontiguousarray(weights, dtype=np.uint32)

    @property
    def n_vars(self):
        return self.w.shape[0].bit_length() - 1

    def add_eq(self, selector, point, scalar):
        O.weights_add_eq(self.w, selector, np.ascontiguousarray(point, dtype=np.uint32).reshape(-1, 5), scalar)

    def add_next(self, selector, point, scalar):
        O.weights_add_next(self.w, selector, np.ascontiguousarray(point, dtype=np.uint32).reshape(-1, 5), scalar)

    def add_strided_eq(self, base, shift, offset, point, scalar):
        pt = np.ascontiguousarray(point, dtype=np.uint32).reshape(-1, 5)
        tab = O.eq_table(pt, np.ascontiguousarray(scalar, dtype=np.uint32)) if pt.shape[0] else np.ascontiguousarray(scalar, dtype=np.uint32)[None, :]
        idx = base + (np.arange(tab.shape[0], dtype=np.int64) << shift) + offset
        self.w[idx] = O.ef_add(self.w[idx], tab)

    def round(self):
        return O.prod_round(self.p, self.w)

    def fold(self, r):
        self.p, self.w = O.fold_msb(self.p, r), O.fold_msb(self.w, r)

    def fold_round(self, r):
        self.fold(r)
        return self.round()

    def read(self):
        return self.p, self.w

    def free(self):
        pass


class OracleBackend(_AirOracleBackend):
    def sumcheck(self, evals, n_vars, live_len=None):
        full = np.zeros(1 << n_vars, dtype=np.uint32)
        full[: evals.size] = evals
        return OracleSumcheck(full)

    def sumcheck_gather(self, d, local, n_vars_total):
        both = self.all_gather_words(d, np.stack([local.p, local.w]))       # world x 2 x n_local x 5
        return OracleSumcheck(both[:, 0].reshape(-1, 5), both[:, 1].reshape(-1, 5))


def main():
    mode, n_vars, folding, live_cols = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(77 + n_vars + folding + live_cols)
    chunk = 1 << (n_vars - folding)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: live_cols * chunk] = O.random_field(rng, live_cols * chunk)
    # statements: a full-length evaluation claim, a claim on one column (selector = column index), and claims whose eq
    # part ends inside / below the rank bits
    g = world.bit_length() - 1
    low = n_vars - folding - g
    stmts = [(0, n_vars), (3 % (1 << folding), n_vars - folding), (5 % (1 << (n_vars - low)), low),
             (1, n_vars - 1), (2 % (1 << (n_vars - low - 1)), low + 1), (0, low - 1 if low > 1 else low)]
    stmts = [(sel % (1 << (n_vars - m)), m, O.random_field(rng, (m, 5)), O.random_field(rng, 5)) for sel, m in stmts]
    # next-row statements (stacked_pcs.rs:73-82) of every position relative to the rank bits: point entirely below them, ending
    # inside them, covering them, and covering the column bits too
    nexts = [(0, n_vars), (1, n_vars - 1), (3 % (1 << folding), n_vars - folding), (5 % (1 << (n_vars - low)), low),
             (2 % (1 << (n_vars - low - 1)), low + 1), (0, max(low - 1, 1)), (7 % (1 << (n_vars - low - g)), low + g)]
    nexts = [(sel % (1 << (n_vars - m)), m, O.random_field(rng, (m, 5)), O.random_field(rng, 5)) for sel, m in nexts]
    n_rounds = folding + 3
    challenges = O.random_field(rng, (n_rounds, 5))

    ref = OracleSumcheck(ev)
    for sel, m, pt, sc in stmts:
        ref.add_eq(sel, pt, sc)
    for sel, m, pt, sc in nexts:
        ref.add_next(sel, pt, sc)
    if mode == "gpu":
        import leanmultisig_b200 as lm
        from leanmultisig_b200.sharded import CudaBackend

        ctx = lm.Context(local_rank, 20)
        backend = CudaBackend(ctx)
    else:
        backend = OracleBackend()
    shard = shard_of(ev, n_vars, folding, rank, world).reshape(1 << folding, -1)[:live_cols].reshape(-1)
    sc = ShardedProductSumcheck(backend, dist, shard, n_vars, folding)
    for sel, m, pt, s in stmts:
        sc.add_eq(sel, pt, s)
    for sel, m, pt, s in nexts:
        sc.add_next(sel, pt, s)
    got, exp = sc.round(), ref.round()
    assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]), f"rank {rank}: round 0 differs"
    for k in range(n_rounds):
        if k % 2 == 0:  # exercise both call shapes
            got, exp = sc.fold_round(challenges[k]), ref.fold_round(challenges[k])
        else:
            sc.fold(challenges[k]), ref.fold(challenges[k])
            got, exp = sc.round(), ref.round()
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]), f"rank {rank}: round {k + 1} differs"
        assert sc.n_vars == n_vars - k - 1
    p, w = sc.read()
    assert np.array_equal(p, ref.p) and np.array_equal(w, ref.w), f"rank {rank}: folded tables differ"
    sc.free()
    dist.barrier()
    if rank == 0:
        print("SHARDED_SC_OK", world, mode)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Worker for tests/test_sharded_whir.py (torch.distributed.run; gloo on CPU or nccl on GPUs).

Every rank builds the same polynomial and statements from a seed, keeps its shard, and runs
leanmultisig_b200.sharded.ShardedWhirProver (commit + prove); transcript, Merkle hints and the final point must equal those
of the single-process oracle prover (oracle/whir.py), and the oracle verifier must accept.  CPU mode: oracle compute doubles
under the product's orchestration and host logic; GPU mode: CUDA backend."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from oracle import whir as W  # noqa: E402
from leanmultisig_b200 import whir_config as WC  # noqa: E402
from leanmultisig_b200.sharded import ShardedWhirProver, shard_of  # noqa: E402
from _sharded_sumcheck_worker import OracleSumcheck  # noqa: E402
from _sharded_worker import OracleBackend as _CommitBackend  # noqa: E402
from _sharded_air_worker import OracleBackend as _CollectiveBackend  # noqa: E402
from test_whir_protocol import SMALL, make_statements, oracle_prove, oracle_verify, to_product_statements  # noqa: E402


class OracleTree:
    """a round commitment on the oracle (test double for whir.Tree)"""

    def __init__(self, cw, layers, full_width, dim):
        self.cw, self.layers, self.full_width, self.elem_dim, self.root = cw, layers, full_width, dim, layers[-1]

    def open(self, idx):
        opened = [O.merkle_open(self.cw, self.full_width, self.layers, int(i)) for i in idx]
        return np.stack([r for r, _ in opened]), np.stack([p for _, p in opened])

    def free(self):
        pass


class OracleSession(OracleSumcheck):
    """the replicated single-device session surface (commit_poly / eval_poly / add_base_eq) on the oracle"""

    def add_base_eq(self, points, scalars):
        O.weights_add_base_eq(self.w, np.ascontiguousarray(points, dtype=np.uint32), np.ascontiguousarray(scalars, dtype=np.uint32))

    def eval_poly(self, point):
        return O.mle_eval(self.p, np.ascontiguousarray(point, dtype=np.uint32).reshape(-1, 5))

    def commit_poly(self, ff, log_inv_rate):
        n = self.n_vars
        cw = O.reorder_and_dft(self.p, n, 5, ff, log_inv_rate, 1 << ff)
        layers = O.merkle_tree(cw, 5 << ff, 5 << ff)
        return OracleTree(cw, layers, 5 << ff, 5)


class OracleBackend(_CommitBackend, _CollectiveBackend):
    def sumcheck(self, evals, n_vars, live_len=None):
        full = np.zeros(1 << n_vars, dtype=np.uint32)
        full[: evals.size] = evals
        return OracleSession(full)

    def sumcheck_gather(self, d, local, n_vars_total):
        both = self.all_gather_words(d, np.stack([local.p, local.w]))
        return OracleSession(both[:, 0].reshape(-1, 5), both[:, 1].reshape(-1, 5))

    def mle_eval(self, evals, point_m):
        return O.mle_eval(evals, np.ascontiguousarray(point_m, dtype=np.uint32).reshape(-1, 5))


def main():
    mode, nv, live_frac_16 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(31 + nv + live_frac_16)
    kw = SMALL
    cfg_o, cfg_p = W.WhirConfig(nv, **kw), WC.WhirConfig(nv, **kw)
    k = cfg_p.first_folding
    live_cols = (1 << k) * live_frac_16 // 16
    live = live_cols << (nv - k)
    poly = O.random_field(rng, 1 << nv)
    poly[live:] = 0
    stm = make_statements(rng, poly, nv, with_next=True)  # incl. a next-row statement (stacked_pcs.rs:73-82)
    ps_o, point_o = oracle_prove(cfg_o, poly, stm, live)

    if mode == "gpu":
        import leanmultisig_b200 as lm
        from leanmultisig_b200.sharded import CudaBackend

        ctx = lm.Context(local_rank, 22)
        backend = CudaBackend(ctx)
        ps = lm.ProverState(ctx)
    else:
        backend = OracleBackend()
        ps = W.ProverState()
    shard = shard_of(poly, nv, k, rank, world).reshape(1 << k, -1)[:live_cols].reshape(-1)
    prover = ShardedWhirProver(backend, dist, cfg_p)
    wit = prover.commit(ps, shard, live_cols)
    point = prover.prove(ps, to_product_statements(stm), wit)
    assert ps.transcript == ps_o.transcript, f"rank {rank}: transcript differs from the single-process oracle prover"
    assert point == point_o
    assert len(ps.merkle_paths) == len(ps_o.merkle_paths)
    for ga, oa in zip(ps.merkle_paths, ps_o.merkle_paths):
        for (gl, gp, gi), (ol, op, oi) in zip(ga, oa):
            assert gi == oi and np.array_equal(gl, ol) and np.array_equal(gp, op)
    assert oracle_verify(cfg_o, ps.transcript, ps.merkle_paths, stm) == point
    dist.barrier()
    if rank == 0:
        print("SHARDED_WHIR_OK", world, mode)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

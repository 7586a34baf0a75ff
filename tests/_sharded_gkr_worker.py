"""Worker for tests/test_sharded_gkr.py (torch.distributed.run; gloo on CPU or nccl on GPUs).

Every rank builds the same fraction table from a seed, keeps its row range, and runs
leanmultisig_b200.sharded.ShardedGkrQuotientProver against a transcript; the transcript and the outputs must equal those
of the single-process oracle prover (oracle/logup.py::prove_gkr_quotient_cpu), and the oracle verifier must accept."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle import logup as OL  # noqa: E402
from oracle import whir as W  # noqa: E402
from leanmultisig_b200.sharded import ShardedGkrQuotientProver  # noqa: E402
from _sharded_air_worker import OracleBackend as _AirOracleBackend  # noqa: E402

ONE_M = int(O.to_monty(1))


class OracleGkrShard:
    """one rank's GKR compute on the oracle (test double for GkrShardSession in the gloo tier)"""

    def __init__(self, nums, dens, n_vars, top_vars):
        n = 1 << n_vars
        pn = np.zeros(n, dtype=np.uint32)
        pn[: nums.size] = nums
        pd = np.zeros((n, 5), dtype=np.uint32)
        pd[:, 0] = ONE_M
        pd[: dens.shape[0]] = dens
        self.layers = [(pn, pd)]
        for _ in range(n_vars - top_vars):
            self.layers.append(O.gkr_layer_up(*self.layers[-1]))
        self.n_vars = n_vars

    def top(self):
        return self.layers[-1]

    def layer_begin(self, claim_vars, point_m, alpha_m, eq_scale_m):
        lay_n, lay_d = self.layers[self.n_vars - (claim_vars + 1)]
        emb = O.embed if lay_n.ndim == 1 else (lambda x: x)
        self.cols = [emb(lay_n[0::2]), emb(lay_n[1::2]), np.ascontiguousarray(lay_d[0::2]), np.ascontiguousarray(lay_d[1::2])]
        self.point = np.ascontiguousarray(point_m, dtype=np.uint32).reshape(-1, 5)
        self.alpha, self.scale = alpha_m, eq_scale_m

    def round(self):
        m = self.cols[0].shape[0].bit_length() - 1
        c0, c2 = O.gkr_round(*self.cols, self.point[: m - 1], self.alpha)
        return O.ef_mul(c0, self.scale), O.ef_mul(c2, self.scale)

    def fold(self, r_m):
        self.cols = [O.fold_lsb(c, r_m) for c in self.cols]

    def layer_end(self):
        return np.stack([c[0] for c in self.cols])

    def free(self):
        pass


class OracleBackend(_AirOracleBackend):
    def gkr_session(self, nums, dens, n_vars, top_vars):
        return OracleGkrShard(nums, dens, n_vars, top_vars)


def main():
    mode, n_vars, active = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(1000 + n_vars + active)
    nums, dens = O.random_field(rng, active), O.random_field(rng, (active, 5))
    per = (1 << n_vars) // world
    lo, hi = min(rank * per, active), min((rank + 1) * per, active)
    if mode == "gpu":
        import leanmultisig_b200 as lm
        from leanmultisig_b200.sharded import CudaBackend

        ctx = lm.Context(local_rank, 20)
        backend = CudaBackend(ctx)
    else:
        backend = OracleBackend()
    ps_ref = W.ProverState()
    q_ref, pt_ref, cn_ref, cd_ref = OL.prove_gkr_quotient_cpu(ps_ref, nums, dens)
    ps = W.ProverState()
    prover = ShardedGkrQuotientProver(backend, dist, nums[lo:hi], dens[lo:hi], n_vars)
    q, pt, cn, cd = prover.prove_with_state(ps)
    prover.free()
    assert ps.transcript == ps_ref.transcript, f"rank {rank}: transcript differs from the single-process oracle prover"
    assert np.array_equal(q, W.tm(q_ref)) and np.array_equal(cn, W.tm(cn_ref)) and np.array_equal(cd, W.tm(cd_ref))
    assert np.array_equal(pt, np.stack([W.tm(x) for x in pt_ref]))
    vs = W.VerifierState(ps.transcript, [])
    vq, vpt, vcn, vcd = OL.verify_gkr_quotient(vs, n_vars)
    assert vs.off == len(ps.transcript) and vq == q_ref and vcn == cn_ref and vcd == cd_ref
    dist.barrier()
    if rank == 0:
        print("SHARDED_GKR_OK", world, mode)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

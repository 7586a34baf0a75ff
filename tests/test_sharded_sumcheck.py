"""N > 1 path of the WHIR product sumcheck: `folding` local rounds with one all-reduce of (c0, c2) each, then an all-gather
of the folded tables and replicated rounds.  CPU tier: gloo + oracle compute double; GPU tier
(>= 2 GPUs): CUDA backend over NCCL.  `localize_statement` is also checked exhaustively on a small case."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from leanmultisig_b200 import field as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_sharded_sumcheck_worker.py")


def run_worker(world, mode, n_vars, folding, live_cols, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, mode, str(n_vars), str(folding), str(live_cols)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"SHARDED_SC_OK {world} {mode}" in out.stdout


def test_localize_statement_matches_the_global_weights(rng):
    """for every rank: the shard of the global weight table equals the table of the localized statement"""
    from leanmultisig_b200.sharded import localize_statement, shard_of

    n_vars, folding, g = 7, 2, 2
    for m in range(0, n_vars + 1):
        for sel in {0, (1 << (n_vars - m)) - 1, 5 % (1 << (n_vars - m))}:
            pt, sc = O.random_field(rng, (m, 5)), O.random_field(rng, 5)
            full = np.zeros((1 << n_vars, 5), dtype=np.uint32)
            O.weights_add_eq(full, sel, pt, sc)
            for rank in range(1 << g):
                exp = np.stack([shard_of(np.ascontiguousarray(full[:, c]), n_vars, folding, rank, 1 << g) for c in range(5)], axis=1)
                got = np.zeros((1 << (n_vars - g), 5), dtype=np.uint32)
                loc = localize_statement(n_vars, folding, g, rank, sel, [F.from_monty(x) for x in pt])
                if loc is not None:
                    lsel, lpt, scale = loc
                    lpm = np.stack([F.to_monty(x) for x in lpt]) if lpt else np.zeros((0, 5), dtype=np.uint32)
                    O.weights_add_eq(got, lsel, lpm, F.to_monty(F.mul(scale, F.from_monty(sc))))
                assert np.array_equal(got, exp), (m, sel, rank)


@pytest.mark.parametrize("world,n_vars,folding,live_cols,port", [(2, 10, 3, 8, 29681), (4, 11, 4, 9, 29682), (2, 9, 2, 3, 29683)])
def test_sharded_product_sumcheck_gloo(world, n_vars, folding, live_cols, port):
    run_worker(world, "cpu", n_vars, folding, live_cols, port)


@pytest.mark.gpu
def test_sharded_product_sumcheck_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    run_worker(2, "gpu", 16, 7, 64, 29691)
    run_worker(4 if n >= 4 else 2, "gpu", 15, 4, 13, 29692)

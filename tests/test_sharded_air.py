"""N > 1 path of the AIR sumcheck: row-range shards, one all-reduce of the round sums per round, all-gather of the
per-shard column values for the last log2(G) rounds.

CPU tier: world_size 2 and 4 over gloo with the oracle as compute backend.  GPU tier (>= 2 GPUs): CUDA backend, NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_sharded_air_worker.py")


def run_worker(world, mode, table, log_rows, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, mode, hex(table), str(log_rows)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"SHARDED_AIR_OK {world} {mode}" in out.stdout


@pytest.mark.parametrize("world,table,log_rows,port", [(2, 0, 5, 29641), (4, 0, 6, 29642), (2, 1, 4, 29643),
                                                       (2, 0x102, 3, 29644), (4, 2, 3, 29645)])
def test_sharded_air_sumcheck_gloo(world, table, log_rows, port):
    run_worker(world, "cpu", table, log_rows, port)


@pytest.mark.gpu
def test_sharded_air_sumcheck_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    run_worker(2, "gpu", 0, 12, 29651)
    run_worker(4 if n >= 4 else 2, "gpu", 1, 8, 29652)
    run_worker(2, "gpu", 2, 6, 29653)

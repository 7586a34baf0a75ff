"""N > 1 path of the Logup quotient GKR: row-range shards of the fraction table, local up pass, all-gather of the top
values, one all-reduce of (c0, c2) per local sumcheck round, the last log2(G) rounds of every layer on gathered values
(SURVEY.md section 8e).  CPU tier: gloo + oracle compute double; GPU tier (>= 2 GPUs): CUDA backend over NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_sharded_gkr_worker.py")


def run_worker(world, mode, n_vars, active, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, mode, str(n_vars), str(active)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"SHARDED_GKR_OK {world} {mode}" in out.stdout


# active lengths: full table, ragged inside the last shard, and one that leaves the last shard(s) all padding
@pytest.mark.parametrize("world,n_vars,active,port", [(2, 8, 256, 29661), (2, 9, 300, 29662), (4, 9, 257, 29663)])
def test_sharded_gkr_gloo(world, n_vars, active, port):
    run_worker(world, "cpu", n_vars, active, port)


@pytest.mark.gpu
def test_sharded_gkr_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    run_worker(2, "gpu", 14, (1 << 14) - 77, 29671)
    run_worker(4 if n >= 4 else 2, "gpu", 12, 2100, 29672)

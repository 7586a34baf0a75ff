"""N > 1 path: the row-sharded commit (sharded NTT with one all-to-all, subtree roots all-gather).

CPU tier: world_size 2 and 4 over gloo with the oracle as compute backend — exercises the product's orchestration,
index maps and collectives.  GPU tier (needs >= 2 GPUs): the same worker with the CUDA backend over NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_sharded_worker.py")


def run_worker(world, mode, n_vars, folding, rate, live_cols, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, mode, str(n_vars), str(folding), str(rate), str(live_cols)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"SHARDED_OK {world} {mode}" in out.stdout


def test_geometry():
    from leanmultisig_b200.sharded import ShardGeometry

    g = ShardGeometry(28, 7, 1, 8)
    assert (g.h, g.block, g.run) == (1 << 22, 1 << 19, 1 << 16)
    owned = [0] * 8
    for row in range(0, g.h, g.run):
        owned[g.owner(row)] += 1
    assert owned == [8] * 8
    # local rows are a bijection onto [0, h / G) per rank
    seen = set()
    for row in range(0, g.h, 4099):
        seen.add((g.owner(row), g.local_row(row)))
    assert len(seen) == len(range(0, g.h, 4099))
    assert sorted(g.subtree_index(r, m) for r in range(8) for m in range(8)) == list(range(64))


@pytest.mark.parametrize("world,n_vars,folding,rate,live_cols,port", [(2, 12, 4, 1, 16, 29621), (2, 13, 4, 1, 8, 29622),
                                                                      (4, 14, 4, 2, 16, 29623)])
def test_sharded_commit_gloo(world, n_vars, folding, rate, live_cols, port):
    run_worker(world, "cpu", n_vars, folding, rate, live_cols, port)


@pytest.mark.gpu
def test_sharded_commit_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    run_worker(world, "gpu", 18, 7, 1, 64, 29631)
    run_worker(2, "gpu", 16, 7, 1, 64, 29632)

"""N > 1 path of WHIR commit + open: the row-sharded initial commitment, sharded out-of-domain evaluations, the local
sumcheck rounds with all-reduce, owner-routed STIR openings of the first tree, replicated tail.
The sharded prover's transcript, hints and final point equal the single-process oracle prover's; the oracle verifier accepts.
CPU tier: gloo + oracle compute doubles; GPU tier (>= 2 GPUs): CUDA backend over NCCL."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_sharded_whir_worker.py")


def run_worker(world, mode, nv, live_frac_16, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, mode, str(nv), str(live_frac_16)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"SHARDED_WHIR_OK {world} {mode}" in out.stdout


@pytest.mark.parametrize("world,nv,live_frac_16,port", [(2, 13, 16, 29701), (2, 12, 6, 29702)])
def test_sharded_whir_gloo(world, nv, live_frac_16, port):
    run_worker(world, "cpu", nv, live_frac_16, port)


@pytest.mark.gpu
def test_sharded_whir_nccl():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    run_worker(2, "gpu", 13, 16, 29711)

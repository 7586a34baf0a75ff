import os, sys, traceback, faulthandler
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
try:
    import leanmultisig_b200 as lm
    from leanmultisig_b200.sharded import CudaBackend, ShardedCommit, shard_of
    ctx = lm.Context(lr, 24)
    b = CudaBackend(ctx)
    n_vars, folding, rate, cols = 16, 7, 1, 64
    rng = np.random.default_rng(1)
    ev = rng.integers(0, 0x7F000001, size=1 << n_vars, dtype=np.uint32)
    shard = shard_of(ev, n_vars, folding, dist.get_rank(), dist.get_world_size())
    sc = ShardedCommit(b, dist, n_vars, folding, rate, live_cols=cols)
    print(lr, "commit...", flush=True)
    root = sc.commit(b.to_device(shard))
    print(lr, "root", root[:2], flush=True)
    b.close()
except Exception:
    traceback.print_exc()
dist.destroy_process_group()

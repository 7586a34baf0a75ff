"""mle_eval alone: time of evaluating a 2^n polynomial (live prefix = half) at an extension point through lm_tree_eval.
python tools/prof_mle_eval.py [n_vars=28]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import leanmultisig_b200 as lm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
ctx = lm.Context(0, 24)
live = 1 << (n - 1)
d = torch.randint(0, 0x7F000001, (live,), dtype=torch.int64, device="cuda").to(torch.int32)


class _Buf:
    ptr = d.data_ptr()


tree = ctx.commit_dev(_Buf, n, 1, 7, 1, live, retain_evals=True)
pt = np.random.default_rng(0).integers(0, 0x7F000001, size=(n, 5), dtype=np.uint32)
for _ in range(3):
    ctx.sync(); t0 = time.perf_counter(); v = tree.evaluate(pt); ctx.sync(); dt = time.perf_counter() - t0
    print(f"evaluate 2^{n} ({live} live): {dt * 1e3:.3f} ms  {live * 4 / dt / 1e9:.0f} GB/s")
tree.free(); ctx.close()

import sys, numpy as np
sys.path.insert(0, ".")
import leanmultisig_b200 as lm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(0)
P = 0x7F000001
ctx = lm.Context(0, 24)
p = rng.integers(0, P, size=1 << n, dtype=np.uint32)
sc = ctx.sumcheck(p, n)
pts = rng.integers(0, P, size=(K, n, 5), dtype=np.uint32)
scs = rng.integers(0, P, size=(K, 5), dtype=np.uint32)
import time
for _ in range(3):
    t0 = time.perf_counter(); sc.add_eq_batch(0, pts, scs); print("add_eq_batch ms", (time.perf_counter() - t0) * 1e3)
sc.free(); ctx.close()

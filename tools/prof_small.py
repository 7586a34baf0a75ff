"""Small instances of the config-3 kernels for ncu captures: AIR execution table at 2^log_rows, GKR at 2^log_gkr."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leanmultisig_b200 as lm
P = 0x7F000001
log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
log_gkr = int(sys.argv[2]) if len(sys.argv) > 2 else 14
rng = np.random.default_rng(0)
rf = lambda shape: rng.integers(0, P, size=shape, dtype=np.uint32)
ctx = lm.Context(0, 20)
if log_rows:
    n = 1 << log_rows
    cols = [rf(n) for _ in range(20)]
    sess = lm.AirSumcheckSession(ctx, 0, cols, rf((log_rows, 5)), rf(5), rf((14, 5)), rf((8, 5)), rf(5))
    ps = lm.NativeProverState(ctx)
    lm.prove_batched_air_sumcheck_native([sess], rf(5), ps)
    sess.free(); ps.free()
if log_gkr:
    N = (1 << log_gkr) - 123
    g = lm.GkrQuotientProver(ctx, rf(N), rf((N, 5)))
    ps = lm.NativeProverState(ctx)
    g.prove_native(ps)
    g.free(); ps.free()
ctx.close()

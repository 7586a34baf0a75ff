"""Quotient GKR alone (BASELINE config 3, Logup half): up pass and device-driven down pass at 2^log_n fractions.
python tools/time_gkr.py [log_n=25] [reps=3]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leanmultisig_b200 as lm

P = 0x7F000001
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
ctx = lm.Context(0, 20)
N = (1 << log_n) - 12345
nums = rng.integers(0, P, size=N, dtype=np.uint32)
dens = rng.integers(0, P, size=(N, 5), dtype=np.uint32)
for it in range(reps):
    t0 = time.perf_counter(); g = lm.GkrQuotientProver(ctx, nums, dens); ctx.sync(); t_up = time.perf_counter() - t0
    ps = lm.NativeProverState(ctx)
    t0 = time.perf_counter(); g.prove_native(ps); t_dev = time.perf_counter() - t0
    g.free(); ps.free()
    print(f"GKR 2^{log_n}: new (H2D + transpose + up pass) {t_up*1e3:.1f} ms; down pass, device challenger {t_dev*1e3:.2f} ms")
ctx.close()

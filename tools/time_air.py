"""Execution-table AIR sumcheck alone: per-round wall clock through lm_air_round / lm_air_fold (first rounds) and the
whole session through lm_air_prove_batched.   python tools/time_air.py [log_rows=22] [reps=2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leanmultisig_b200 as lm
P = 0x7F000001
log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(0)
rf = lambda shape: rng.integers(0, P, size=shape, dtype=np.uint32)
ctx = lm.Context(0, 20)
n = 1 << log_rows
cols = [rf(n) for _ in range(20)]
eqf, ap, la, beta = rf((log_rows, 5)), rf((14, 5)), rf((8, 5)), rf(5)
for it in range(reps):
    sess = lm.AirSumcheckSession(ctx, 0, cols, eqf, rf(5), ap, la, beta)
    tr = []
    for r in range(log_rows):
        t0 = time.perf_counter(); bare = sess.compute_bare_round_poly(); t1 = time.perf_counter()
        sess.process_challenge(rf(5), bare)
        tr.append(t1 - t0)
    sess.final_column_evals(); sess.free()
    sess = lm.AirSumcheckSession(ctx, 0, cols, eqf, rf(5), ap, la, beta)
    ps = lm.NativeProverState(ctx)
    t0 = time.perf_counter(); lm.prove_batched_air_sumcheck_native([sess], rf(5), ps); t_nat = time.perf_counter() - t0
    sess.free(); ps.free()
    print(f"AIR exec 2^{log_rows}: rounds ms {[round(x*1e3, 2) for x in tr[:8]]} ... last {round(tr[-1]*1e3, 3)}; sum {sum(tr)*1e3:.1f} ms; "
          f"lm_air_prove_batched {t_nat*1e3:.1f} ms")
ctx.close()

"""Timing of BASELINE config 3 pieces on one GPU: execution-table AIR sumcheck (2^log_rows rows, 24 rounds) and the
quotient GKR over N fractions; wall-clock per phase through the C-ABI sessions (host transcript replaced by a PRNG)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leanmultisig_b200 as lm
from leanmultisig_b200 import field as F

P = 0x7F000001
log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24
log_gkr = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rng = np.random.default_rng(0)
ctx = lm.Context(0, 24)

def rf(shape):
    return rng.integers(0, P, size=shape, dtype=np.uint32)

# ---- AIR sumcheck, execution table
n = 1 << log_rows
cols = [rf(n) for _ in range(20)]
eqf, ap, la, beta = rf((log_rows, 5)), rf((14, 5)), rf((8, 5)), rf(5)
t0 = time.perf_counter()
sess = lm.AirSumcheckSession(ctx, 0, cols, eqf, rf(5), ap, la, beta)
t_setup = time.perf_counter() - t0
t_round, t_fold = [], []
for r in range(log_rows):
    t0 = time.perf_counter(); bare = sess.compute_bare_round_poly(); t1 = time.perf_counter()
    sess.process_challenge(rf(5), bare); t2 = time.perf_counter()
    t_round.append(t1 - t0); t_fold.append(t2 - t1)
finals = sess.final_column_evals()
sess.free()
print(f"AIR exec 2^{log_rows}: setup(H2D 20 cols) {t_setup*1e3:.1f} ms; rounds total {sum(t_round)*1e3:.1f} ms; folds total {sum(t_fold)*1e3:.1f} ms")
print("  first rounds ms:", [round(x*1e3, 2) for x in t_round[:6]], " folds:", [round(x*1e3, 2) for x in t_fold[:6]])
print("  last 8 rounds ms:", [round(x*1e3, 3) for x in t_round[-8:]])
bytes_unfused = 22 * n * 4 + 22 * (n // 2) * 20 + sum(22 * (n >> k) * 20 + 22 * (n >> (k + 1)) * 20 for k in range(1, log_rows))
tot = sum(t_round) + sum(t_fold)
print(f"  algorithmic bytes (unfused) {bytes_unfused/2**30:.2f} GiB -> {bytes_unfused/tot/1e9:.1f} GB/s")

# ---- the same AIR sumcheck with the round loop in the C++ spine (lm_air_prove_batched)
sess = lm.AirSumcheckSession(ctx, 0, cols, eqf, rf(5), ap, la, beta)
ps = lm.NativeProverState(ctx)
t0 = time.perf_counter(); lm.prove_batched_air_sumcheck_native([sess], rf(5), ps); t_nat = time.perf_counter() - t0
sess.free()
print(f"  native spine (C++ round loop + transcript): {t_nat*1e3:.1f} ms for {log_rows} rounds -> {bytes_unfused/t_nat/1e9:.1f} GB/s")

# ---- GKR
if log_gkr < 15:
    ctx.close(); sys.exit(0)
N = (1 << log_gkr) - 12345
nums, dens = rf(N), rf((N, 5))
t0 = time.perf_counter(); g = lm.GkrQuotientProver(ctx, nums, dens); t_up = time.perf_counter() - t0
class T:
    def __init__(s): s.n = 0
    def add_scalars(s, v): pass
    def add_sumcheck_poly(s, c, a): s.n += 1
    def sample(s): return rf(5)
tr = T()
t0 = time.perf_counter(); g.prove(tr.add_scalars, tr.add_sumcheck_poly, tr.sample); t_down = time.perf_counter() - t0
print(f"GKR 2^{log_gkr}: new (H2D + up pass) {t_up*1e3:.1f} ms; down pass {t_down*1e3:.1f} ms over {tr.n} rounds")
g.free()
g = lm.GkrQuotientProver(ctx, nums, dens)
ps = lm.NativeProverState(ctx)
t0 = time.perf_counter(); g.prove_native(ps); t_nat = time.perf_counter() - t0
gkr_bytes = 230 * N
print(f"  native spine: down pass {t_nat*1e3:.1f} ms; up + down algorithmic bytes ~230 N = {gkr_bytes/2**30:.2f} GiB")
g.free(); ctx.close()

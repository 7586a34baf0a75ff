"""Verifier side: all openings of a STIR round against the round's root (lm_verify_openings) and the restore of the pruned
paths on batched device hashes, timed through the C ABI with HOST buffers, next to the CPU oracle's per-opening loop.
python tools/time_verify.py [log_h=21] [width=128] [n=256] [reps=20]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leanmultisig_b200 as lm
from leanmultisig_b200 import verify as V
from leanmultisig_b200.merkle_pruning import prune
import oracle as O

log_h = int(sys.argv[1]) if len(sys.argv) > 1 else 21
width = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
rng = np.random.default_rng(0)
ctx = lm.Context(0, 24)
fold_vars = (width).bit_length() - 1
ev = O.random_field(rng, 1 << (log_h + fold_vars - 1))
tree = ctx.commit(ev, log_h + fold_vars - 1, fold_vars, 1)          # height 2^log_h, leaves of `width` base elements
assert tree.log_height == log_h and tree.full_width == width
idx = rng.integers(0, tree.height, n, dtype=np.uint64)
pt = O.random_field(rng, (fold_vars, 5))
rows, paths, evals = tree.open_fold(idx, pt)
ok, ev2 = ctx.verify_openings(tree.root, log_h, idx, rows, paths, elem_dim=1, fold_point=pt)
assert ok.all() and np.array_equal(ev2, evals)
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    ok, _ = ctx.verify_openings(tree.root, log_h, idx, rows, paths, elem_dim=1, fold_point=pt)
    ts.append(time.perf_counter() - t0)
t_gpu = min(ts)
t0 = time.perf_counter()
for q in range(n):
    assert O.merkle_verify(tree.root, log_h, int(idx[q]), rows[q], paths[q])
t_cpu = time.perf_counter() - t0
comp = n * (width // 8 - 1 + log_h)
print(f"verify_openings: {n} openings, tree 2^{log_h} x {width}: device call {t_gpu*1e6:.0f} us best of {reps} "
      f"({np.median(ts)*1e6:.0f} us median; {comp} compressions, {n*(4*width+32*log_h)/1e3:.0f} KB in), "
      f"oracle loop on one host core {t_cpu*1e3:.2f} ms")
pruned = prune([int(i) for i in idx], rows, paths)
h = V.DeviceHasher(ctx)
assert V.restore(pruned, h) is not None
ts = []
for _ in range(3):
    t0 = time.perf_counter(); V.restore(pruned, h); ts.append(time.perf_counter() - t0)
print(f"restore of the pruned batch ({len(pruned.paths)} distinct leaves, {pruned.n_digests()} digests kept): "
      f"{min(ts)*1e3:.2f} ms with {width // 8 - 1 + log_h} batched hash launches")
tree.free(); ctx.close()

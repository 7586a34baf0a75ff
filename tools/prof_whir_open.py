"""WHIR open of a committed 2^n polynomial alone (the config-3 block of bench_extra.py), for launch lists / ncu:
python tools/prof_whir_open.py [n_vars=28] [n_statements=8] [reps=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_extra as B
import leanmultisig_b200 as lm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = lm.Context(0, 24)
B.cpu_whir_sample = lambda *a: (1.0, 1)
out = B.measure_whir_open(ctx, torch, 6451.8, n, k, reps, 0)
print({key: out[key] for key in ("ms", "ms_best") if key in out})
ctx.close()

"""WHIR open of the committed 2^n_vars polynomial alone (the `whir_open` block of the bench line), for A/B runs of a switch:
LM_WHIR_PY_ROUNDS=1 python tools/time_whir_open.py [n_vars=28] [statements=8] [reps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leanmultisig_b200 as lm
import bench_extra as BX

n_vars = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n_st = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = lm.Context(0, 24)
r = BX.measure_whir_open(ctx, torch, 6451.8, n_vars, n_st, reps, 14)
print(f"WHIR open 2^{n_vars}, {n_st} statements, LM_WHIR_PY_ROUNDS={os.environ.get('LM_WHIR_PY_ROUNDS', '(unset: rounds in the C++ spine)')}: "
      f"{r['ms']:.2f} ms mean, {r['ms_best']:.2f} best of {reps}; transcript words {r['transcript_words']}")
ctx.close()

"""AIR constraint ORDER and polynomials pinned by the reference's own source text.

tests/golden/air_constraints.json (execution table, extension_op precompile) and air_constraints_poseidon16.json (the
poseidon16 precompile, tools/gen_air_golden_poseidon16.py) hold the value of every constraint at a random point, in `assert_zero` call order, obtained by mechanically translating and EXECUTING the
reference's `Air::eval` bodies (tools/gen_air_golden.py - no formula retyped).  Constraint k is multiplied by alpha^k in the
sumcheck (constraint_folder/normal.rs:49-62), so alpha = the k-th unit vector isolates it: the oracle (CPU tier) and the
CUDA sessions (GPU tier) must reproduce each value, which pins the alpha-power assignment that "all constraints vanish on
a valid trace" cannot see."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "air_constraints.json")))
# the poseidon16 table (100 constraints over 109 columns) comes from its own translator, tools/gen_air_golden_poseidon16.py
GOLDEN["tables"] = GOLDEN["tables"] + json.load(open(os.path.join(ROOT, "tests", "golden", "air_constraints_poseidon16.json")))["tables"]
TABLE_ID = {"execution": 0, "extension_op": 1, "poseidon16": 2}


def _m(x):
    return O.to_monty(np.array(x, dtype=np.uint64))


def _unit_alphas(n, k):
    ap = np.zeros((n, 5), dtype=np.uint32)
    ap[k, 0] = int(O.to_monty(1))
    return ap


def _expected(c):
    v = np.zeros(5, dtype=np.uint64)
    if c["kind"] == "ext":
        v[:] = c["value"]
    else:
        v[0] = c["value"]
    return _m(v)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present (GPU box)")
def test_golden_file_is_what_the_reference_source_evaluates_to():
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_air_golden.py"), "--check"], stdout=subprocess.DEVNULL)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_air_golden_poseidon16.py"), "--check"],
                          stdout=subprocess.DEVNULL)


@pytest.mark.parametrize("t", GOLDEN["tables"], ids=lambda t: t["table"])
def test_oracle_constraints_match_reference_source(t):
    tid = TABLE_ID[t["table"]]
    n_cols, n_shift, _ = O.air_shape(tid)
    assert n_cols == len(t["flat"]) and n_shift == len(t["shift"])
    point = np.zeros((n_cols + n_shift, 5), dtype=np.uint32)
    point[:, 0] = _m(t["flat"] + t["shift"])
    la, beta = _m(t["logup_alphas_eq_poly"]), _m(t["bus_beta"])
    n = len(t["constraints"])
    for k, c in enumerate(t["constraints"]):
        got = O.air_eval(tid, point, _unit_alphas(n, k), la, beta)
        assert np.array_equal(got, _expected(c)), f"{t['table']}: constraint {k} differs from the reference source"


@pytest.mark.gpu
@pytest.mark.parametrize("t", GOLDEN["tables"], ids=lambda t: t["table"])
def test_gpu_constraints_match_reference_source(rng, t):
    """A 2-row table whose first row is the golden point and whose second row carries the golden shifted values: the first
    round's z = 0 evaluation with alpha = e_k is constraint k at the point (one pair, eq weight 1)."""
    import leanmultisig_b200 as lm

    tid = TABLE_ID[t["table"]]
    ctx = lm.Context(0, 20)
    flat, shift = _m(t["flat"]), _m(t["shift"])
    cols = []
    for c in range(len(flat)):
        row1 = shift[c] if c < len(shift) else O.random_field(rng, 1)[0]
        cols.append(np.array([flat[c], row1], dtype=np.uint32))
    la, beta = _m(t["logup_alphas_eq_poly"]), _m(t["bus_beta"])
    eq_factor = O.random_field(rng, (1, 5))
    n = len(t["constraints"])
    for k, c in enumerate(t["constraints"]):
        sess = lm.AirSumcheckSession(ctx, tid, cols, eq_factor, np.zeros(5, dtype=np.uint32), _unit_alphas(max(n, 14), k), la, beta)
        raw = sess._raw_round()
        sess.free()
        assert np.array_equal(raw[0], _expected(c)), f"{t['table']}: constraint {k} differs from the reference source"
    ctx.close()

"""Committed fixtures (tests/golden/): the reference's Poseidon1 known-answer vector, and oracle-generated regression pins of
small commits (root, codeword / layer hashes, one OOD evaluation).  CPU tier: the oracle reproduces them.  GPU tier: the
CUDA path through the C ABI reproduces them."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint32).tobytes()).hexdigest()


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def case_input(c):
    rng = np.random.default_rng(c["seed"])
    chunk = 1 << (c["n_vars"] - c["folding"])
    ev = np.zeros(1 << c["n_vars"], dtype=np.uint32)
    ev[: c["live_cols"] * chunk] = rng.integers(0, O.P, size=c["live_cols"] * chunk, dtype=np.uint32)
    y = rng.integers(0, O.P, size=5, dtype=np.uint32)
    assert [int(x) for x in y] == c["ood_y"]
    return ev, O.expand_from_univariate(y, c["n_vars"])


def test_oracle_reproduces_the_reference_kat():
    kat = load("poseidon1_kat.json")
    x = O.to_monty(np.array(kat["input"], dtype=np.uint64))
    for dense in (False, True):
        assert [int(v) for v in O.from_monty(O.poseidon1_permute(x[None, :], dense=dense)[0])] == kat["output"]


@pytest.mark.parametrize("idx", range(4))
def test_oracle_reproduces_the_commit_pins(idx):
    c = load("commit_small.json")["cases"][idx]
    ev, point = case_input(c)
    cw = O.reorder_and_dft(ev, c["n_vars"], 1, c["folding"], c["log_inv_rate"], c["live_cols"])
    layers = O.merkle_tree(cw, 1 << c["folding"], c["live_cols"])
    assert [int(x) for x in layers[-1]] == c["root"]
    assert sha(cw) == c["codeword_sha256"] and sha(layers) == c["layers_sha256"]
    assert [int(x) for x in O.mle_eval(ev, point)] == c["ood_value"]


@pytest.mark.gpu
def test_gpu_reproduces_the_golden_fixtures():
    import leanmultisig_b200 as lm

    ctx = lm.Context(0, 20)
    kat = load("poseidon1_kat.json")
    x = O.to_monty(np.array(kat["input"], dtype=np.uint64))
    assert [int(v) for v in O.from_monty(ctx.poseidon1(x[None, :])[0])] == kat["output"]
    for c in load("commit_small.json")["cases"]:
        ev, point = case_input(c)
        live = c["live_cols"] << (c["n_vars"] - c["folding"])
        tree = ctx.commit(ev, c["n_vars"], c["folding"], c["log_inv_rate"], actual_len=live)
        assert [int(v) for v in tree.root] == c["root"], c["seed"]
        assert sha(tree.codeword()[:, : c["live_cols"]]) == c["codeword_sha256"] and sha(tree.layers()) == c["layers_sha256"]
        assert [int(v) for v in tree.evaluate(point)] == c["ood_value"]
        tree.free()
    ctx.close()

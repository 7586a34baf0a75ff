"""Regenerates tests/golden/*.json.

  poseidon1_kat.json   the reference's ONLY stored vector for this path (crates/backend/koala-bear/src/
                       poseidon1_koalabear_16.rs:1082-1092), copied by hand - not generated.
  commit_small.json    regression pins produced by the ORACLE (oracle/, the CPU restatement) on seeded inputs: the reference is
                       Rust and cannot run here, so these are not reference outputs - they freeze today's oracle so that a
                       later change to the oracle and the kernels together cannot go unnoticed.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint32).tobytes()).hexdigest()


def commit_case(seed, n_vars, folding, rate, live_cols):
    rng = np.random.default_rng(seed)
    chunk = 1 << (n_vars - folding)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: live_cols * chunk] = rng.integers(0, O.P, size=live_cols * chunk, dtype=np.uint32)
    cw = O.reorder_and_dft(ev, n_vars, 1, folding, rate, live_cols)
    layers = O.merkle_tree(cw, 1 << folding, live_cols)
    y = rng.integers(0, O.P, size=5, dtype=np.uint32)
    point = O.expand_from_univariate(y, n_vars)
    return dict(seed=seed, n_vars=n_vars, folding=folding, log_inv_rate=rate, live_cols=live_cols,
                generator="numpy default_rng(seed).integers(0, p, live_cols * 2^(n_vars - folding), uint32) = Montgomery words; then 5 words y",
                root=[int(x) for x in layers[-1]], codeword_sha256=sha(cw), layers_sha256=sha(layers),
                ood_y=[int(x) for x in y], ood_value=[int(x) for x in O.mle_eval(ev, point)])


def main():
    cases = [commit_case(1, 12, 4, 1, 16), commit_case(2, 14, 7, 1, 64), commit_case(3, 15, 7, 2, 40), commit_case(4, 16, 7, 1, 128)]
    with open(os.path.join(HERE, "commit_small.json"), "w") as f:
        json.dump(dict(note="oracle-generated regression pins, see make_golden.py", cases=cases), f, indent=1)
    print("wrote commit_small.json:", [c["root"][0] for c in cases])


if __name__ == "__main__":
    main()

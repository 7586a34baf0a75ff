"""Consistency check against numbers the reference PUBLISHES: proof sizes of `xmss --n-signatures 1550` in the proven regime,
338 KiB at rate 1/2 and 228 KiB at rate 1/4 (README.md:35-36; KiB = proof_size_fe x 31 bits / 8 / 1024,
rec_aggregation/src/benchmark.rs:425).  The WHIR opening dominates a proof; its size follows from exactly the pieces this
repository restates — the query / OOD / round schedule of WhirConfig::new, the leaf widths (2^7 base elements, then 2^5
extension elements), the Merkle heights and the path pruning — so for every candidate size 2^n of the stacked polynomial the
script computes the opening's size at both rates (pruned digests averaged over random query sets) and the part of the published
figure it leaves for everything else (sumcheck / GKR transcripts, evaluations), which does not depend on the rate.
    python tools/proof_size_check.py [trials=40]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from leanmultisig_b200.merkle_pruning import lca_level
from leanmultisig_b200.whir_config import WhirConfig

PUBLISHED = {1: 338, 2: 228}  # KiB, proven regime (Johnson bound), README.md:35-36
# all pairs (rate 1/2, rate 1/4) the README publishes for the proven regime: xmss (README.md:35-36), recursion n = 1..4 (:53-60)
PUBLISHED_PAIRS = {"xmss --n-signatures 1550": (338, 228), "recursion --n 1": (278, 188), "recursion --n 2": (293, 194),
                   "recursion --n 3": (312, 203), "recursion --n 4": (308, 206)}


def pruned_digests(idx, h):
    idx = sorted(set(int(i) for i in idx))
    tot = 0
    for i in range(len(idx)):
        lv = h if i == 0 else lca_level(idx[i - 1], idx[i])
        sk = lca_level(idx[i], idx[i + 1]) - 1 if i + 1 < len(idx) else None
        tot += lv - (1 if sk is not None and sk < lv else 0)
    return tot, len(idx)


def whir_opening_fe(nv, rate, trials, rng):
    cfg = WhirConfig(nv, starting_log_inv_rate=rate)
    fe = 8 + cfg.commitment_ood_samples * 5                                        # root + OOD answers
    n_sc = cfg.first_folding + sum(cfg.folding_at(r + 1) for r in range(cfg.n_rounds)) + cfg.final_sumcheck_rounds
    fe += n_sc * 10                                                                 # (c1, c2) per sumcheck round
    fe += (1 << cfg.n_vars_of_final_polynomial()) * 5                               # final coefficients
    pow_rounds = cfg.first_folding * (cfg.starting_folding_pow_bits > 0)
    for r, rp in enumerate(list(cfg.round_parameters) + [cfg.final_round_config()]):
        h = (rp.domain_size >> rp.folding_factor).bit_length() - 1
        leaf = (1 if r == 0 else 5) << rp.folding_factor
        acc = 0
        for _ in range(trials):
            d, n = pruned_digests(rng.integers(0, 1 << h, rp.num_queries), h)
            acc += n * leaf + 8 * d
        fe += acc / trials + (rp.query_pow_bits > 0)
        if r < cfg.n_rounds:
            fe += 8 + rp.ood_samples * 5
            pow_rounds += cfg.folding_at(r + 1) * (rp.folding_pow_bits > 0)
    return fe + pow_rounds


if __name__ == "__main__":
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(0)
    print("stacked 2^n | opening KiB at rate 1/2, 1/4 | difference")
    kib = {}
    for nv in range(19, 29):
        kib[nv] = {r: whir_opening_fe(nv, r, trials, rng) * 31 / 8 / 1024 for r in (1, 2)}
        print(f"   n = {nv}   |  {kib[nv][1]:6.1f}  {kib[nv][2]:6.1f}         |  {kib[nv][1] - kib[nv][2]:6.1f}")
    print()
    print("published pair (KiB at 1/2, 1/4; difference) -> witness sizes that leave the SAME positive remainder at both rates")
    for name, (a, b) in PUBLISHED_PAIRS.items():
        fits = [(nv, a - kib[nv][1], b - kib[nv][2]) for nv in kib
                if a - kib[nv][1] > 0 and b - kib[nv][2] > 0 and abs((a - kib[nv][1]) - (b - kib[nv][2])) < 6]
        print(f"   {name:26s} {a} {b} ({a - b:3d}) -> " + ", ".join(f"2^{nv}: {x:.1f} / {y:.1f} KiB left" for nv, x, y in fits))

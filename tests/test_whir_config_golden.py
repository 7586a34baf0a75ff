"""WhirConfig::new pinned by executing the reference's own source text (tools/gen_whir_config_golden.py translates
crates/whir/src/config.rs statement by statement): every derived number of the product's and of the oracle's parameter
derivation — queries, OOD samples, PoW bits, rates, domain sizes, generators, final-round configuration — for 12..28
variables at the four supported rates equals the committed golden file; with /root/reference present the translation is
re-run and must reproduce the file."""
import dataclasses
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "whir_config.json")))


def _flatten(cfg):
    out = {"num_variables": cfg.num_variables, "starting_log_inv_rate": cfg.starting_log_inv_rate,
           "commitment_ood_samples": cfg.commitment_ood_samples, "starting_folding_pow_bits": cfg.starting_folding_pow_bits,
           "final_queries": cfg.final_queries, "final_query_pow_bits": cfg.final_query_pow_bits,
           "final_log_inv_rate": cfg.final_log_inv_rate, "final_sumcheck_rounds": cfg.final_sumcheck_rounds,
           "rounds": [dataclasses.asdict(r) for r in cfg.round_parameters]}
    if cfg.round_parameters:
        out["final_round_config"] = dataclasses.asdict(cfg.final_round_config())
    return out


def _impls():
    from leanmultisig_b200.whir_config import WhirConfig as Product
    from oracle.whir import WhirConfig as Oracle

    return [("product", Product), ("oracle", Oracle)]


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_derivation_equals_reference_source(which):
    cls = dict(_impls())[which]
    assert len(GOLD["cases"]) >= 60
    for case in GOLD["cases"]:
        cfg = cls(case["num_variables"], starting_log_inv_rate=case["starting_log_inv_rate"])
        got = _flatten(cfg)
        want = {k: v for k, v in case.items() if k != "merkle_tree_heights"}
        assert got == want, (which, case["num_variables"], case["starting_log_inv_rate"])
        if "merkle_tree_heights" in case:
            # config.rs:361-363: log2 of the tree height of round r = log2(domain size) - folding factor of that round
            rounds = cfg.round_parameters + [cfg.final_round_config()]
            heights = [(r.domain_size.bit_length() - 1) - r.folding_factor for r in rounds]
            assert heights == case["merkle_tree_heights"]


def test_headline_schedules():
    """the numbers SURVEY 8(d) quotes: 230 / 74 queries then 32 at 2^22, 256 / 75 / 32 then 21 at 2^28 (rate 1/2)"""
    by = {(c["num_variables"], c["starting_log_inv_rate"]): c for c in GOLD["cases"]}
    assert [r["num_queries"] for r in by[(22, 1)]["rounds"]] + [by[(22, 1)]["final_queries"]] == [230, 74, 32]
    assert [r["num_queries"] for r in by[(28, 1)]["rounds"]] + [by[(28, 1)]["final_queries"]] == [256, 75, 32, 21]


@pytest.mark.skipif(not os.path.exists("/root/reference/crates/whir/src/config.rs"), reason="reference tree not present")
def test_golden_file_is_current():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import gen_whir_config_golden as g

    fresh = g.generate()
    assert fresh["cases"] == GOLD["cases"]
    assert fresh["translated_python"] == GOLD["translated_python"]


def test_published_proof_sizes_are_consistent_with_the_schedule():
    """README.md:35-36 of the reference publishes 338 KiB (rate 1/2) and 228 KiB (rate 1/4) for the same witness (KiB =
    field elements x 31 bits, rec_aggregation/src/benchmark.rs:425).  The size of the WHIR opening computed from this
    repository's schedule, leaf widths and path pruning (tools/proof_size_check.py) must leave the same positive remainder for
    the rest of the proof at both rates for some witness size; it does at 2^24..2^25 and the published difference lies between
    the two."""
    import numpy as np

    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
    import proof_size_check as psc

    rng = np.random.default_rng(0)
    kib = {nv: {r: psc.whir_opening_fe(nv, r, 20, rng) * 31 / 8 / 1024 for r in (1, 2)} for nv in (24, 25, 27)}
    d24, d25 = kib[24][1] - kib[24][2], kib[25][1] - kib[25][2]
    assert d24 < psc.PUBLISHED[1] - psc.PUBLISHED[2] < d25
    for nv in (24, 25):
        rest = [psc.PUBLISHED[r] - kib[nv][r] for r in (1, 2)]
        assert 25 < rest[0] < 60 and 25 < rest[1] < 60 and abs(rest[0] - rest[1]) < 6
    assert psc.PUBLISHED[1] - kib[27][1] < 0  # a 2^27 witness would already exceed the published size
    # the recursion figures (README.md:53-60): each pair fits one witness size with the same remainder at both rates
    for name, nv in (("recursion --n 1", 21), ("recursion --n 4", 23)):
        a, b = psc.PUBLISHED_PAIRS[name]
        ra, rb = (a - psc.whir_opening_fe(nv, 1, 20, rng) * 31 / 8 / 1024, b - psc.whir_opening_fe(nv, 2, 20, rng) * 31 / 8 / 1024)
        assert 20 < ra < 40 and 20 < rb < 40 and abs(ra - rb) < 4, (name, ra, rb)

"""Extra blocks of bench.py's JSON line: BASELINE config 3 (AIR sumcheck, Logup quotient GKR, WHIR open) and the
metric-(i) PROXY (tools/xmss_proxy.py).  Each block carries ms, algorithmic bytes, an HBM roofline fraction, an end-to-end
figure from pinned HOST buffers and a CPU baseline: the oracle (oracle/, the CPU restatement of the reference) timed on a
bounded sample on the host cores.  Imported by bench.py only (rank 0, N = 1).

Algorithmic bytes are SURVEY.md section 8(d)'s figures: AIR sumcheck over a 2^L-row execution table = 22 columns read in
the base field + every folded table written and read once (15.1 GiB at L = 24, the un-fused count the reference's
compute-then-fold does); quotient GKR = 230 bytes per fraction (up pass + down pass).  WHIR open: 24 B per entry in round
0 (4 B polynomial + 20 B weights), 40 B afterwards, halving per round, plus one pass over the weights per statement.
"""
from __future__ import annotations

import importlib.util
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
P = 0x7F000001


def _rf(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint32)


def air_algorithmic_bytes(log_rows: int) -> int:
    n = 1 << log_rows
    return 22 * n * 4 + 22 * (n // 2) * 20 + sum(22 * (n >> k) * 20 + 22 * (n >> (k + 1)) * 20 for k in range(1, log_rows))


def _best_mean(ts):
    return min(ts) * 1e3, sum(ts) / len(ts) * 1e3


# ------------------------------------------------------------------------------------------------ AIR sumcheck
def cpu_air_sample(log_rows: int, n_threads: int):
    """the oracle's execution-table sumcheck (compute_raw_poly + fold_at_bit per round, as the reference does) on a
    2^log_rows-row table; seconds"""
    import oracle as O

    rng = np.random.default_rng(3)
    cols = _rf(rng, (20, 1 << log_rows))
    cur = np.concatenate([cols, np.stack([O.shift_column(cols[0]), O.shift_column(cols[1])])])
    eqf, ap, la, beta = _rf(rng, (log_rows, 5)), _rf(rng, (14, 5)), _rf(rng, (8, 5)), _rf(rng, 5)
    t0 = time.perf_counter()
    for r in range(log_rows):
        O.air_exec_round(cur, eqf[: log_rows - r - 1], ap, la, beta)
        ch = _rf(rng, 5)
        cur = np.stack([O.fold_lsb(cur[c], ch) for c in range(22)])
    return time.perf_counter() - t0


def measure_air(ctx, torch, peak: float, log_rows: int, reps: int, cpu_log_rows: int) -> dict:
    import leanmultisig_b200 as lm

    rng = np.random.default_rng(11)
    n = 1 << log_rows
    gen = torch.Generator(device="cuda").manual_seed(5)
    d_cols = torch.empty((20, n), dtype=torch.int32, device="cuda")
    for c in range(20):
        d_cols[c] = torch.randint(0, P, (n,), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
    host = torch.empty((20, n), dtype=torch.int32).pin_memory()
    host.copy_(d_cols.cpu())
    host_cols = [host[c].numpy().view(np.uint32) for c in range(20)]
    eqf, ap, la, beta, eta = _rf(rng, (log_rows, 5)), _rf(rng, (14, 5)), _rf(rng, (8, 5)), _rf(rng, 5), _rf(rng, 5)
    sum0 = _rf(rng, 5)

    def run(resident: bool):
        ctx.sync()
        t0 = time.perf_counter()
        if resident:
            sess = lm.AirSumcheckSession(ctx, 0, None, eqf, sum0, ap, la, beta, device_columns=(d_cols.data_ptr(), 20))
        else:
            sess = lm.AirSumcheckSession(ctx, 0, host_cols, eqf, sum0, ap, la, beta)
        ps = lm.NativeProverState(ctx)
        lm.prove_batched_air_sumcheck_native([sess], eta, ps)
        finals = sess.final_column_evals()
        ctx.sync()
        dt = time.perf_counter() - t0
        tr = ps.transcript
        sess.free(), ps.free()
        return dt, finals, tr

    run(True)  # warm-up: buffer cache, kernel images
    res = [run(True) for _ in range(reps)]
    e2e = [run(False) for _ in range(max(2, reps - 1))]
    assert np.array_equal(res[0][1], e2e[-1][1]) and res[0][2] == e2e[-1][2], "resident and host-buffer sessions disagree"
    best, mean = _best_mean([r[0] for r in res])
    e_best, e_mean = _best_mean([r[0] for r in e2e[1:]] or [e2e[0][0]])
    nbytes = air_algorithmic_bytes(log_rows)
    ach = nbytes / (mean * 1e-3) / 1e9
    t_cpu = cpu_air_sample(cpu_log_rows, os.cpu_count())
    del d_cols, host
    return {
        "workload": f"execution-table AIR sumcheck, 2^{log_rows} rows x 20 (+2 shifted) columns, 13 constraints, degree 5, {log_rows} rounds "
                    f"(lm_air_new_dev + lm_air_prove_batched + lm_air_final, transcript included)",
        "ms": mean, "ms_best": best, "reps": reps, "value": n / (mean * 1e-3) / 1e6, "unit": "Mrows/s",
        "algorithmic_bytes": nbytes,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "note": "constraint evaluation in the quintic extension is multiplier-pipe bound (ncu: fmaheavy 47 %, "
                             "issue slots 38 %, profiles/r02_*); bytes = SURVEY 8(d) un-fused count"},
        "e2e": {"ms": e_mean, "ms_best": e_best, "value": n / (e_mean * 1e-3) / 1e6, "unit": "Mrows/s",
                "h2d_bytes_per_step": 20 * n * 4, "d2h_bytes_per_step": 22 * 20 + 25 * 4 * log_rows},
        "cpu_baseline": {"value": (1 << cpu_log_rows) / t_cpu / 1e6, "unit": "Mrows/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"2^{cpu_log_rows}-row table, all {cpu_log_rows} rounds (oracle/air.c, OpenMP), {t_cpu:.2f} s"},
    }


# ------------------------------------------------------------------------------------------------ quotient GKR
def cpu_gkr_sample(log_n: int):
    import oracle as O
    from oracle import logup as OL
    from oracle import whir as W

    rng = np.random.default_rng(4)
    n = (1 << log_n) - 99
    nums, dens = _rf(rng, n), _rf(rng, (n, 5))
    ps = W.ProverState()
    t0 = time.perf_counter()
    OL.prove_gkr_quotient_cpu(ps, nums, dens)
    return time.perf_counter() - t0


def measure_gkr(ctx, torch, peak: float, n_fractions: int, reps: int, cpu_log_n: int) -> dict:
    import leanmultisig_b200 as lm

    gen = torch.Generator(device="cuda").manual_seed(6)
    d_nums = torch.randint(0, P, (n_fractions,), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
    d_dens = torch.empty((n_fractions, 5), dtype=torch.int32, device="cuda")
    for k in range(5):
        d_dens[:, k] = torch.randint(0, P, (n_fractions,), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
    h_nums = torch.empty(n_fractions, dtype=torch.int32).pin_memory()
    h_dens = torch.empty((n_fractions, 5), dtype=torch.int32).pin_memory()
    h_nums.copy_(d_nums.cpu()), h_dens.copy_(d_dens.cpu())

    def run(resident: bool):
        ctx.sync()
        t0 = time.perf_counter()
        if resident:
            g = lm.GkrQuotientProver.from_device(ctx, d_nums.data_ptr(), d_dens.data_ptr(), n_fractions)
        else:
            g = lm.GkrQuotientProver(ctx, h_nums.numpy().view(np.uint32), h_dens.numpy().view(np.uint32))
        ctx.sync()
        t1 = time.perf_counter()
        ps = lm.NativeProverState(ctx)
        out = g.prove_native(ps)
        ctx.sync()
        t2 = time.perf_counter()
        tr = ps.transcript
        g.free(), ps.free()
        return t2 - t0, t1 - t0, t2 - t1, out, tr

    run(True)
    res = [run(True) for _ in range(reps)]
    e2e = [run(False) for _ in range(max(2, reps - 1))]
    assert res[0][4] == e2e[-1][4], "resident and host-buffer GKR transcripts disagree"
    best, mean = _best_mean([r[0] for r in res])
    e_best, e_mean = _best_mean([r[0] for r in e2e[1:]] or [e2e[0][0]])
    nbytes = 230 * n_fractions
    ach = nbytes / (mean * 1e-3) / 1e9
    t_cpu = cpu_gkr_sample(cpu_log_n)
    n_vars = (n_fractions - 1).bit_length()
    del d_nums, d_dens, h_nums, h_dens
    return {
        "workload": f"Logup quotient GKR over {n_fractions} fractions (2^{n_vars} padded; base numerators, extension denominators): "
                    f"up pass + {sum(range(5, n_vars))} sumcheck rounds with the challenger on the device (lm_gkr_new_dev + lm_gkr_prove)",
        "ms": mean, "ms_best": best, "reps": reps, "value": n_fractions / (mean * 1e-3) / 1e6, "unit": "Mfractions/s",
        "phases_ms": {"copy + transpose + up pass": sum(r[1] for r in res) / len(res) * 1e3,
                      "down pass (all layer sumchecks)": sum(r[2] for r in res) / len(res) * 1e3},
        "algorithmic_bytes": nbytes,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "note": "small layers are bound by the latency of the transcript step (3 Poseidon1 permutations per round, "
                             "~25 us on one warp), large ones by extension-field products; bytes = SURVEY 8(d) 230 B per fraction"},
        "e2e": {"ms": e_mean, "ms_best": e_best, "value": n_fractions / (e_mean * 1e-3) / 1e6, "unit": "Mfractions/s",
                "h2d_bytes_per_step": n_fractions * 24, "d2h_bytes_per_step": 64 * 20 + n_vars * 20 + 40},
        "cpu_baseline": {"value": (1 << cpu_log_n) / t_cpu / 1e6, "unit": "Mfractions/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"2^{cpu_log_n} fractions, whole prove_gkr_quotient (oracle/gkr.c + Python spine), {t_cpu:.2f} s"},
    }


# ------------------------------------------------------------------------------------------------ WHIR open
def whir_open_bytes(n_vars: int, n_statements: int, folding=(7, 5), batch: int = 12, live: int | None = None) -> int:
    """Bytes the opening has to move in the form the product runs it: the statements are combined `batch` at a time in one
    read-modify-write pass over the weights (the reference makes one pass PER statement, open.rs:518-584: pass batch=1 for
    that count), round 0 reads the live base-field entries and the weights, later rounds read and write extension tables."""
    total = -(-n_statements // batch) * (1 << n_vars) * 40
    n, first = n_vars, True
    while n > 0:
        if first:
            total += (live if live is not None else (1 << n)) * 4 + (1 << n) * 20
        else:
            total += (1 << n) * 40                  # read p and w
        total += (1 << (n - 1)) * 40               # write the folded tables
        first = False
        n -= 1
    return total


def cpu_whir_sample(n_vars: int, n_statements: int):
    import oracle as O
    from oracle import whir as W

    rng = np.random.default_rng(5)
    cfg = W.WhirConfig(n_vars)
    poly = _rf(rng, 1 << n_vars)
    ps = W.ProverState()
    t0 = time.perf_counter()
    wit = W.cpu_commit(cfg, ps, poly, poly.size)
    t_commit = time.perf_counter() - t0
    stmts = []
    for _ in range(n_statements):
        pt = _rf(rng, (n_vars, 5))
        stmts.append(W.SparseStatement.dense([W.fm(x) for x in pt], W.fm(O.mle_eval(poly, pt))))
    t0 = time.perf_counter()
    W.cpu_prove(cfg, ps, stmts, wit, poly)
    return time.perf_counter() - t0, t_commit


def measure_whir_open(ctx, torch, peak: float, n_vars: int, n_statements: int, reps: int, cpu_n_vars: int) -> dict:
    import leanmultisig_b200 as lm
    from leanmultisig_b200.whir import Witness, _sample_ood

    rng = np.random.default_rng(12)
    live = 1 << (n_vars - 1)
    gen = torch.Generator(device="cuda").manual_seed(8)
    d_poly = torch.randint(0, P, (live,), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
    cfg = lm.WhirConfig(n_vars)

    class _Buf:
        ptr = d_poly.data_ptr()

    def run():
        ps = lm.NativeProverState(ctx)
        tree = ctx.commit_dev(_Buf, n_vars, 1, cfg.first_folding, cfg.starting_log_inv_rate, live, retain_evals=True)
        ps.add_base_scalars(tree.root)
        pts, answers = _sample_ood(ps, cfg.commitment_ood_samples, n_vars, tree.evaluate)
        wit = Witness(tree, pts, answers)
        stmts = []
        for _ in range(n_statements):
            pt = _rf(rng, (n_vars, 5))
            stmts.append(lm.SparseStatement.dense(pt, tree.evaluate(pt)))
        ctx.sync()
        t0 = time.perf_counter()
        lm.WhirProver(ctx, cfg).prove(ps, stmts, wit)
        ctx.sync()
        dt = time.perf_counter() - t0
        n_tr = len(ps.transcript)
        wit.free(), ps.free()
        return dt, n_tr

    run()
    res = [run() for _ in range(reps)]
    best, mean = _best_mean([r[0] for r in res])
    nbytes = whir_open_bytes(n_vars, n_statements + cfg.commitment_ood_samples, live=live)
    nbytes_ref_form = whir_open_bytes(n_vars, n_statements + cfg.commitment_ood_samples, batch=1)
    ach = nbytes / (mean * 1e-3) / 1e9
    t_cpu, _ = cpu_whir_sample(cpu_n_vars, n_statements)
    del d_poly
    return {
        "workload": f"WHIR open (WhirConfig::prove, open.rs:37-248) of the committed 2^{n_vars}-variable polynomial ({live} live entries), "
                    f"{n_statements} dense evaluation statements + {cfg.commitment_ood_samples} OOD, production parameters "
                    f"(folding 7/5, rate 1/2, PoW <= 16 bits on the device), round commitments and STIR openings included",
        "ms": mean, "ms_best": best, "reps": reps, "value": (1 << n_vars) / (mean * 1e-3) / 1e9, "unit": "Gelem/s",
        "algorithmic_bytes": nbytes,
        "algorithmic_bytes_one_pass_per_statement": nbytes_ref_form,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "note": "bytes = one batched read-modify-write pass over the weights for all statements + the sumcheck "
                             "rounds (the count with one pass per statement, as the reference runs it, is "
                             "algorithmic_bytes_one_pass_per_statement and is NOT what frac uses); round commitments, PoW and "
                             "STIR openings are inside the time but not in the bytes"},
        "e2e": {"ms": mean, "value": (1 << n_vars) / (mean * 1e-3) / 1e9, "unit": "Gelem/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": None,
                "note": "the witness of an opening is the prover data the commit left on the device (commit.rs:11-57): there is no "
                        "host input to copy; every opening row, path and round polynomial is read back inside the timed region"},
        "cpu_baseline": {"value": (1 << cpu_n_vars) / t_cpu / 1e9, "unit": "Gelem/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"2^{cpu_n_vars}-variable polynomial, same statements and parameters (oracle/whir.py over oracle C "
                                   f"kernels), {t_cpu:.2f} s"},
        "transcript_words": res[0][1],
    }


def measure_config3(ctx, torch, peak: float, quick: bool = False) -> dict:
    out = {}
    log_rows = 20 if quick else 24
    out["air_sumcheck"] = measure_air(ctx, torch, peak, log_rows, 3, 14 if quick else 17)
    torch.cuda.empty_cache()
    n_frac = (7 << (log_rows - 4)) if quick else 7 << 24
    out["logup_gkr"] = measure_gkr(ctx, torch, peak, n_frac, 3, 14 if quick else 18)
    torch.cuda.empty_cache()
    out["whir_open"] = measure_whir_open(ctx, torch, peak, 22 if quick else 28, 8, 2 if quick else 3, 16 if quick else 20)
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ metric (i) proxy
def measure_xmss_proxy(n_signatures: int = 1550, reps: int = 5) -> dict:
    spec = importlib.util.spec_from_file_location("xmss_proxy", os.path.join(ROOT, "tools", "xmss_proxy.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.run(n_signatures, reps)

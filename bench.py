#!/usr/bin/env python3
"""Benchmark of the B200 WHIR-commit hot path (BASELINE.json configs[1]) — one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--log-rows 22]

A step is one WHIR commit of the stacked witness: gather + evals-DFT (Reed-Solomon encode, rate 1/2) into a
2^22 x 64 KoalaBear codeword matrix (full leaf width 128) followed by the Poseidon1 Merkle tree over its rows
(8 sponge compressions per leaf + 2^22 - 1 tree compressions).  `value` is codeword field elements per second
(2^28 per commit and GPU) with the polynomial already resident in HBM; `e2e` is the same metric through the
reference-facing C-ABI call `lm_commit` with a pinned HOST buffer, host->device copy and root read-back inside
the timed region.  N > 1 (default `--multi sharded`): ONE commit of the same shape row-sharded over the ranks
(leanmultisig_b200/sharded.py: exchange inside the NTT, all-gather of the subtree roots) - the total work is fixed, so the
line says "scaling": "strong"; the root is checked against a single-GPU commit of the whole polynomial outside the timed
region.  `--multi replicas` runs one independent commit per GPU instead (weak scaling).

At N = 1 the line also carries `config3` (BASELINE config 3: execution-table AIR sumcheck, Logup quotient GKR, WHIR open,
each with ms, algorithmic bytes, roofline, e2e and cpu_baseline) and `xmss_proxy` (metric (i) PROXY, best and worst of 5
passes); `--no-extras` skips both (bench_extra.py).

`--impl reference` times the reference algorithm's CPU restatement (oracle/, the reference is Rust and cannot
be built in this image) on ALL host cores over a bounded sample of the same workload; under torchrun the launcher exports
OMP_NUM_THREADS=1, which this arm overrides before the OpenMP runtime starts.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 0x7F000001
FOLDING, LOG_INV_RATE = 7, 1
METRIC = "WHIR commit NTT+Merkle throughput (codeword field elements/s)"
UNIT = "Gelem/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-rows", type=int, default=22, help="log2 of codeword rows (BASELINE config: 22)")
    ap.add_argument("--cpu-sample-log-rows", type=int, default=0, help="0 = pick from the core count")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1: one commit row-sharded over the GPUs (strong scaling) or one commit per GPU (weak)")
    ap.add_argument("--no-extras", action="store_true", help="skip the config3 and xmss_proxy blocks of the N = 1 line")
    ap.add_argument("--quick-extras", action="store_true", help="config3 / proxy at reduced sizes (smoke runs of bench.py itself)")
    return ap.parse_args()


def workload(log_rows: int) -> dict:
    n_vars = log_rows + FOLDING - LOG_INV_RATE
    return dict(n_vars=n_vars, live=1 << (n_vars - 1), rows=1 << log_rows)


def peaks() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self) -> dict:
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------- CPU arm
def use_all_host_threads() -> int:
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the reference's Rayon-style
    all-core path, so the thread count is set explicitly - in the environment BEFORE liboracle.so (libgomp) is loaded,
    and through the library's own export in case the OpenMP runtime is already up."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    import oracle as O

    lib = O.lib()
    lib.lm_or_set_num_threads(n)
    return int(lib.lm_or_max_threads())


def cpu_commit_sample(sample_log_rows: int, reps: int):
    """Oracle (CPU restatement of the reference algorithms, OpenMP over rows) on a 2^sample_log_rows x 64 slice
    of the workload; returns (seconds per commit of the sample, elements per sample, cores)."""
    import numpy as np

    import oracle as O

    n_vars = sample_log_rows + FOLDING - LOG_INV_RATE
    live = 1 << (n_vars - 1)
    rng = np.random.default_rng(0)
    ev = rng.integers(0, P, size=live, dtype=np.uint32)
    ev_full = np.zeros(1 << n_vars, dtype=np.uint32)
    ev_full[:live] = ev
    O.lib()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        cw = O.reorder_and_dft(ev_full, n_vars, 1, FOLDING, LOG_INV_RATE, 64)
        layers = O.merkle_tree(cw, 128, 64)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, 1 << n_vars, os.cpu_count(), layers[-1]


def pick_cpu_sample(cores: int) -> int:
    # AVX-512 oracle: ~1 us per Poseidon1 compression per core (16 lanes), 9 compressions per row -> the whole 2^22-row
    # commit is ~10-20 core-seconds; hosts without AVX-512 fall back to the scalar path (~10x slower): smaller slice
    import oracle as O

    if O.lib().lm_or_have_avx512():
        return 22 if cores >= 8 else 20
    return 19 if cores >= 8 else 17


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = use_all_host_threads()
    slr = args.cpu_sample_log_rows or pick_cpu_sample(cores)
    slr = min(slr, args.log_rows)
    for _ in range(max(args.warmup, 1)):
        cpu_commit_sample(min(slr, 14), 1)  # spins up the OpenMP pool and pages the library in
    times = []
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        t, elems, _, _ = cpu_commit_sample(slr, 1)
        times.append(t)
    t_step = sum(times) / len(times)
    value = elems / t_step / 1e9
    import oracle as O

    simd = "AVX-512 16-lane" if O.lib().lm_or_have_avx512() else "scalar"
    sample = (f"2^{slr} x 64 slice of the 2^{args.log_rows} x 64 commit (rate 1/2, full width 128), {simd} port, "
              f"all host threads (OpenMP)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "strong" if (args.gpus > 1 and args.multi == "sharded") else "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"WHIR commit 2^{args.log_rows}x64 KoalaBear, rate 1/2 (evals-DFT + Poseidon1 Merkle)",
                   "n_vars": args.log_rows + FOLDING - LOG_INV_RATE, "folding_factor": FOLDING, "log_inv_rate": LOG_INV_RATE,
                   "live_cols": 64, "full_cols": 128, "elements_per_commit": 1 << (args.log_rows + FOLDING - LOG_INV_RATE),
                   "sample": sample, "omp_threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust (no toolchain in this image): timed the C restatement in oracle/ (kind=port; AVX-512 "
                "vertical Poseidon1 + cache-blocked vector DFT when the host has AVX-512)",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import leanmultisig_b200 as L
    from leanmultisig_b200._lib import check, lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1 and args.multi == "sharded":
        return run_b200_sharded(args, rank, local_rank, world)

    wl = workload(args.log_rows)
    n_vars, live, rows = wl["n_vars"], wl["live"], wl["rows"]
    elems_per_commit = 1 << n_vars  # codeword elements incl. the implicit zero half of every leaf (2^22 x 128)
    l = lib()
    ctx = L.Context(local_rank, 24)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    with torch.cuda.stream(stream):
        ev = torch.randint(0, P, (live,), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
        cw = torch.empty((rows, 64), dtype=torch.int32, device="cuda")
        layers = torch.empty((2 * rows - 1, 8), dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    host = torch.empty(live, dtype=torch.int32).pin_memory()
    host.copy_(ev.cpu())
    root = np.empty(8, dtype=np.uint32)

    ev_ptr, cw_ptr, ly_ptr = ev.data_ptr(), cw.data_ptr(), layers.data_ptr()

    def step_dev(events=None):
        if events:
            events[0].record(stream)
        check(l.lm_dev_reorder_and_dft(ctx.handle, ev_ptr, n_vars, 1, FOLDING, LOG_INV_RATE, 64, cw_ptr))
        if events:
            events[1].record(stream)
        check(l.lm_dev_merkle_leaves(ctx.handle, cw_ptr, rows, 64, 128, 64, ly_ptr))
        if events:
            events[2].record(stream)
        check(l.lm_dev_merkle_levels(ctx.handle, ly_ptr, rows))
        if events:
            events[3].record(stream)

    def step_e2e():
        t = C.c_void_p()
        check(l.lm_commit(ctx.handle, C.c_void_p(host.data_ptr()), n_vars, 1, live, FOLDING, LOG_INV_RATE, C.byref(t),
                          root.ctypes.data_as(C.POINTER(C.c_uint32))))
        check(l.lm_tree_free(t))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_dev()
            flush.zero_()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = l.lm_kernel_launches()
        # ---- kernel-level timed region: K steps, L2 flushed between steps (flush excluded via events per step)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        for k in range(args.steps):
            step_dev(evs[k])
            flush.zero_()
        barrier()
        launches = l.lm_kernel_launches() - launches0
        t_ntt = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
        t_leaf = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
        t_lvl = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
        t_step = sum(e[0].elapsed_time(e[3]) for e in evs) / args.steps
        # ---- end-to-end through lm_commit with host buffers
        for _ in range(min(args.warmup, 3)):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        barrier()
        t_e2e_wall = (time.perf_counter() - t0) * 1e3 / args.steps
        t_e2e = max(e0.elapsed_time(e1) / args.steps, t_e2e_wall)
        sampler.stop_flag.set()
        sampler.join(timeout=3)

    # max over ranks
    if world > 1:
        tt = torch.tensor([t_step, t_e2e, t_ntt, t_leaf, t_lvl], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step, t_e2e, t_ntt, t_leaf, t_lvl = tt.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel: leaf_sponge_kernel — reads the live codeword, writes one digest per row
        leaf_bytes = rows * 64 * 4 + rows * 32
        ach = leaf_bytes / (t_leaf * 1e-3) / 1e9
        n_perm_leaf = rows * 8
        line = {
            "metric": METRIC, "value": world * elems_per_commit / (t_step * 1e-3) / 1e9, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step,
            "higher_is_better": True, "scaling": "weak" if world > 1 else "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"WHIR commit 2^{args.log_rows}x64 KoalaBear, rate 1/2 (evals-DFT + Poseidon1 Merkle)",
                       "n_vars": n_vars, "folding_factor": FOLDING, "log_inv_rate": LOG_INV_RATE,
                       "live_cols": 64, "full_cols": 128, "elements_per_commit": elems_per_commit,
                       "l2": "256 MiB flush write between timed steps; inputs (0.5 GiB) and codeword (1 GiB) exceed L2",
                       "parallelism": "1 commit per GPU (replicas)" if world > 1 else "single GPU"},
            "breakdown_ms": {"reorder_and_dft": t_ntt, "leaf_sponge": t_leaf, "tree_levels": t_lvl},
            "roofline": {"bound": "hbm", "kernel": "leaf_sponge_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": LEAF_DRAM_TRAFFIC_2_22 if args.log_rows == 22 else None,
                         "traffic_source": "profiles/r02c_ncu_commit_kernels.txt (ncu --set full, dram read + write per launch)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": leaf_bytes,
                         "note": "Poseidon1 with its linear maps on tcgen05 (kind::i8, TMEM accumulators); what is left is bound by the IMAD.WIDE / IMAD dispatch of the S-boxes and reductions (fmaheavy 79 %, issue slots 48 %, tensor pipe 19 %), ~3.1k SASS instr per compression (6.7k before), see DESIGN.md 2.1b",
                         "gperm_per_s": n_perm_leaf / (t_leaf * 1e-3) / 1e9},
            # the bound that actually limits the dominant kernel (not measured live: from the committed ncu capture of this
            # command, same kernel build) - SURVEY 8d asks for a second roofline against the integer pipe
            "roofline_alu": {"bound": "fmaheavy pipe (IMAD.WIDE / IMAD dispatch)", "kernel": "leaf_sponge_kernel",
                             "frac": 0.792, "issue_slot_frac": 0.481, "tensor_pipe_frac": 0.195, "unit": "pipe-active fraction",
                             "source": "ncu sm__pipe_fmaheavy_cycles_active / smsp__issue_active / sm__pipe_tensor_cycles_active, profiles/r02c_ncu_commit_kernels.txt"},
            "roofline_commit": {"bound": "hbm", "achieved": (live * 4 + rows * 64 * 4 + (2 * rows - 1) * 32) / (t_step * 1e-3) / 1e9,
                                "peak": peak, "unit": "GB/s"},
            "roofline_ntt": {"bound": "hbm", "achieved": (live * 4 + rows * 64 * 4) / (t_ntt * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s"},
            "e2e": {"value": world * elems_per_commit / (t_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": t_e2e,
                    "h2d_bytes_per_step": live * 4, "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        line["roofline_commit"]["frac"] = line["roofline_commit"]["achieved"] / peak
        line["roofline_ntt"]["frac"] = line["roofline_ntt"]["achieved"] / peak
        if world == 1:
            cores = use_all_host_threads()
            slr = min(args.cpu_sample_log_rows or pick_cpu_sample(cores), args.log_rows)
            cpu_commit_sample(min(slr, 14), 1)  # warm-up: OpenMP pool start-up costs ~1 s on the first call
            t_cpu, elems, cores, cpu_root = cpu_commit_sample(slr, 1)
            import oracle as O

            simd = "AVX-512 (16 lanes, the reference's packing width)" if O.lib().lm_or_have_avx512() else "scalar"
            line["cpu_baseline"] = {"value": elems / t_cpu / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"2^{slr} x 64 slice of the commit, oracle/ C restatement, {simd}, OpenMP on "
                                              f"all host threads, {t_cpu:.2f} s"}
            if slr == args.log_rows:
                # same seed, same shape: the GPU root of the timed commits must be the oracle's (full-size parity, every run)
                line["cpu_baseline"]["root_matches_gpu"] = bool(np.array_equal(np.asarray(cpu_root).reshape(-1), gpu_root_of_cpu_input(
                    l, ctx, torch, n_vars, live)))
            if not args.no_extras:
                del ev, cw, layers, host
                torch.cuda.empty_cache()
                import bench_extra

                line["config3"] = bench_extra.measure_config3(ctx, torch, peak, quick=args.quick_extras)
        print(json.dumps(line), flush=True) if (world > 1 or args.no_extras) else None
    ctx.close()
    if rank == 0 and world == 1 and not args.no_extras:
        # metric (i) PROXY on its own context (the proxy creates and destroys one per run), after the commit context is gone
        import bench_extra

        torch.cuda.empty_cache()
        try:
            line["xmss_proxy"] = bench_extra.measure_xmss_proxy(64 if args.quick_extras else 1550, 5)
        except Exception as exc:  # the proxy must never cost the headline line
            line["xmss_proxy"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gpu_root_of_cpu_input(l, ctx, torch, n_vars, live):
    """lm_commit of the polynomial the CPU arm commits (numpy seed 0), root only"""
    import numpy as np

    from leanmultisig_b200._lib import check

    rng = np.random.default_rng(0)
    ev = rng.integers(0, P, size=live, dtype=np.uint32)
    root = np.empty(8, dtype=np.uint32)
    t = C.c_void_p()
    check(l.lm_commit(ctx.handle, ev.ctypes.data_as(C.c_void_p), n_vars, 1, live, FOLDING, LOG_INV_RATE, C.byref(t),
                      root.ctypes.data_as(C.POINTER(C.c_uint32))))
    check(l.lm_tree_free(t))
    return root


def run_b200_sharded(args, rank, local_rank, world):
    """N > 1: ONE commit of the BASELINE shape, row-sharded over the ranks (leanmultisig_b200/sharded.py):
    local sub-transform, one all-to-all inside the NTT, last log2(N) layers, per-rank Merkle subtrees, all-gather
    of the N^2 subtree roots, replicated top.  Strong scaling: the total work is fixed."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import leanmultisig_b200 as L
    from leanmultisig_b200._lib import lib
    from leanmultisig_b200.sharded import CudaBackend, ShardedCommit

    wl = workload(args.log_rows)
    n_vars, live = wl["n_vars"], wl["live"]
    elems_per_commit = 1 << n_vars
    l = lib()
    ctx = L.Context(local_rank, 24)
    backend = CudaBackend(ctx)
    stream = backend.stream
    sc = ShardedCommit(backend, dist, n_vars, FOLDING, LOG_INV_RATE, live_cols=64)
    shard_len = live // world
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    shard = torch.randint(0, P, (shard_len,), dtype=torch.int64, device="cuda", generator=gen).to(torch.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    host = torch.empty(shard_len, dtype=torch.int32).pin_memory()
    host.copy_(shard.cpu())

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        root = sc.commit(shard)
        flush.zero_()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = l.lm_kernel_launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        evs[k][0].record(stream)
        root = sc.commit(shard)
        evs[k][1].record(stream)
        flush.zero_()
    barrier()
    launches = l.lm_kernel_launches() - launches0
    t_step = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    phases = backend.phase_times() if os.environ.get("LM_SHARD_TIMING") else None
    # e2e: pinned host shard -> device -> commit -> root on the host
    for _ in range(min(args.warmup, 3)):
        sc.commit_host(host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        root = sc.commit_host(host)   # copy of column group k+1 overlaps transform / exchange / hash of group k
    barrier()
    t_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    tt = torch.tensor([t_step, t_e2e], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step, t_e2e = tt.tolist()
    roots = [None] * world
    dist.all_gather_object(roots, [int(x) for x in np.asarray(root).reshape(-1)])
    # outside the timed region: the sharded root must be the root of ONE single-GPU commit of the whole polynomial
    # (rank q holds, for every column chunk of 2^(n - k) entries, the positions whose top log2(N) bits are q)
    gathered = [torch.empty_like(shard) for _ in range(world)] if rank == 0 else None
    dist.gather(shard, gathered, dst=0)
    single_root = None
    if rank == 0:
        n_cols = live >> (n_vars - FOLDING)
        full = torch.stack([g.view(n_cols, -1) for g in gathered], dim=1).contiguous().view(-1)  # [column][rank][position]
        del gathered

        class _Buf:
            ptr = full.data_ptr()

        torch.cuda.synchronize()
        tree = ctx.commit_dev(_Buf, n_vars, 1, FOLDING, LOG_INV_RATE, live)
        single_root = [int(x) for x in tree.root]
        tree.free()
        del full
    if rank == 0:
        assert all(r == roots[0] for r in roots), "ranks disagree on the Merkle root"
        assert single_root == roots[0], "sharded Merkle root differs from the single-GPU commit of the same polynomial"
        peak, peak_src = peaks()
        rows = wl["rows"]
        commit_bytes = live * 4 + rows * 64 * 4 + (2 * rows - 1) * 32
        ach = commit_bytes / (t_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": elems_per_commit / (t_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"WHIR commit 2^{args.log_rows}x64 KoalaBear, rate 1/2 (evals-DFT + Poseidon1 Merkle)",
                       "n_vars": n_vars, "folding_factor": FOLDING, "log_inv_rate": LOG_INV_RATE, "live_cols": 64,
                       "full_cols": 128, "elements_per_commit": elems_per_commit,
                       "l2": "256 MiB flush write between timed steps",
                       "parallelism": f"one commit row-sharded over {world} GPUs: all-to-all inside the NTT, "
                                      f"all-gather of {world * world} subtree roots"},
            "roofline": {"bound": "hbm", "kernel": "whole sharded commit (aggregate over ranks)", "achieved": ach,
                         "peak": peak * world, "unit": "GB/s", "frac": ach / (peak * world), "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": commit_bytes},
            "e2e": {"value": elems_per_commit / (t_e2e * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": t_e2e,
                    "h2d_bytes_per_step": live * 4, "d2h_bytes_per_step": 32 * world},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "root_check": "sharded root == single-GPU lm_commit_dev root of the gathered polynomial (outside the timed region)",
        }
        if phases:
            line["phases_ms_rank0_last_step"] = phases
        print(json.dumps(line), flush=True)
    ctx.close()
    dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE leaf_sponge_kernel launch at the default workload (2^22 rows), from
# the committed ncu capture; the algorithmic figure is 1 207 959 552 B, i.e. no re-reads.
LEAF_DRAM_TRAFFIC_2_22 = 1_203_177_000  # ncu --set full: dram__bytes_read.sum 1.074693 GB + dram__bytes_write.sum 0.128484 GB per launch


def main():
    args = parse()
    # Libraries (NCCL's version banner, torchrun notices) may print to fd 1; the contract is ONE JSON line on stdout,
    # so everything else is sent to stderr and the line is written to the saved descriptor.
    global print
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    import builtins

    def print(*a, **k):  # noqa: A001
        k.pop("flush", None)
        builtins.print(*a, file=real_stdout, flush=True, **k)

    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

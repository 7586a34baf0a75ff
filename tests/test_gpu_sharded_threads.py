"""The sharded AIR sumcheck and the sharded quotient GKR on ONE GPU: G "ranks" run as threads of this process, each with
its own library context (own stream and scratch), and exchange through an in-memory stand-in for torch.distributed.
Exercises lm_air_new_shard / lm_air_new_folded / lm_gkr_new_shard / lm_gkr_layer_begin_shard and the orchestration of
leanmultisig_b200/sharded.py on boxes that have a single device; the multi-process NCCL versions are in
test_sharded_air.py / test_sharded_gkr.py."""
import threading

import numpy as np
import pytest

import oracle as O
from oracle import logup as OL
from oracle import whir as W

P = 0x7F000001


class ThreadGroup:
    def __init__(self, world):
        self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def exchange(self, rank, value):
        self.slots[rank] = value
        self.barrier.wait()
        out = list(self.slots)
        self.barrier.wait()
        return out


class ThreadDist:
    def __init__(self, group, rank):
        self.group, self.rank = group, rank

    def get_rank(self):
        return self.rank

    def get_world_size(self):
        return self.group.world


class ThreadBackend:
    """CUDA compute through the C ABI, collectives through the ThreadGroup"""

    def __init__(self, ctx):
        self.ctx = ctx

    def air_session(self, table_id, columns, eq_factor, ap, la, beta, **kw):
        import leanmultisig_b200 as lm

        return lm.AirSumcheckSession(self.ctx, table_id, columns, eq_factor, np.zeros(5, dtype=np.uint32), ap, la, beta, **kw)

    def gkr_session(self, nums, dens, n_vars, top_vars):
        from leanmultisig_b200.logup import GkrShardSession

        return GkrShardSession(self.ctx, nums, dens, n_vars, top_vars)

    def sumcheck(self, evals, n_vars, live_len=None):
        return self.ctx.sumcheck(evals, n_vars, live_len)

    def sumcheck_gather(self, d, local, n_vars_total):
        import torch

        n_local = 5 << local.n_vars
        mine = torch.empty((2, n_local), dtype=torch.int32, device="cuda")
        local.export_dev(mine[0].data_ptr(), mine[1].data_ptr())
        torch.cuda.synchronize()
        everyone = torch.stack(d.group.exchange(d.rank, mine))       # same device: the tensors are shared between threads
        tables = everyone.permute(1, 0, 2).contiguous()
        torch.cuda.synchronize()
        return self.ctx.sumcheck_from_dev(tables[0].data_ptr(), tables[1].data_ptr(), n_vars_total)

    def all_reduce_field(self, d, words):
        parts = d.group.exchange(d.rank, np.asarray(words).astype(np.int64))
        return (sum(parts) % P).astype(np.uint32)

    def all_gather_words(self, d, words):
        return np.stack(d.group.exchange(d.rank, np.array(words, dtype=np.uint32)))


def run_ranks(world, fn):
    import leanmultisig_b200 as lm

    group, results, errors = ThreadGroup(world), [None] * world, []

    def body(rank):
        try:
            ctx = lm.Context(0, 16)
            try:
                results[rank] = fn(ThreadBackend(ctx), ThreadDist(group, rank), rank)
            finally:
                ctx.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            group.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.gpu
@pytest.mark.parametrize("world,n_vars,active", [(2, 10, 1000), (4, 11, 1030), (8, 12, 4096)])
def test_gpu_sharded_gkr_threads(rng, world, n_vars, active):
    from leanmultisig_b200.sharded import ShardedGkrQuotientProver

    nums, dens = O.random_field(rng, active), O.random_field(rng, (active, 5))
    ps_ref = W.ProverState()
    q_ref, pt_ref, cn_ref, cd_ref = OL.prove_gkr_quotient_cpu(ps_ref, nums, dens)
    per = (1 << n_vars) // world

    def rank_main(backend, dist, rank):
        lo, hi = min(rank * per, active), min((rank + 1) * per, active)
        ps = W.ProverState()
        prover = ShardedGkrQuotientProver(backend, dist, nums[lo:hi], dens[lo:hi], n_vars)
        out = prover.prove_with_state(ps)
        prover.free()
        return ps.transcript, out

    for transcript, (q, pt, cn, cd) in run_ranks(world, rank_main):
        assert transcript == ps_ref.transcript
        assert np.array_equal(q, W.tm(q_ref)) and np.array_equal(cn, W.tm(cn_ref)) and np.array_equal(cd, W.tm(cd_ref))
    vs = W.VerifierState(ps_ref.transcript, [])
    OL.verify_gkr_quotient(vs, n_vars)


@pytest.mark.gpu
@pytest.mark.parametrize("world,table,log_rows", [(2, 0, 10), (4, 1, 7), (8, 2, 6)])
def test_gpu_sharded_air_threads(rng, world, table, log_rows):
    from leanmultisig_b200.sharded import ShardedAirSumcheckSession
    from test_air_tables import extras, oracle_rounds, with_shifts
    from leanmultisig_b200 import field as F

    n_cols, n_shift, deg = O.air_shape(table)
    base = O.random_field(rng, (n_cols, 1 << log_rows))
    eq_factor = O.random_field(rng, (log_rows, 5))
    ap, la, beta = extras(rng)
    challenges = O.random_field(rng, (log_rows, 5))
    sum0 = O.random_field(rng, 5)
    raws, finals = oracle_rounds(table, with_shifts(table, base), eq_factor, ap, la, beta, challenges)
    # expected bare polynomials from the oracle's raw round sums (air_sumcheck.rs:242-265)
    expected, s, mmf = [], F.from_monty(sum0), F.ONE
    for r in range(log_rows):
        alpha = F.from_monty(eq_factor[log_rows - 1 - r])
        p_evals = [F.mul(F.from_monty(v), mmf) for v in raws[r]]
        p1 = F.mul(F.sub(s, F.mul(F.sub(F.ONE, alpha), p_evals[0])), F.inv(alpha))
        coeffs = F.lagrange_interpolation_at_integers([p_evals[0], p1] + p_evals[1:])
        expected.append(np.stack([F.to_monty(c) for c in coeffs]))
        ch = F.from_monty(challenges[r])
        eq_eval = F.add(F.mul(F.sub(F.ONE, alpha), F.sub(F.ONE, ch)), F.mul(alpha, ch))
        s, mmf = F.mul(F.poly_eval(coeffs, ch), eq_eval), F.mul(mmf, eq_eval)
    per = (1 << log_rows) // world

    def rank_main(backend, dist, rank):
        shard = [base[c, rank * per:(rank + 1) * per] for c in range(n_cols)]
        sess = ShardedAirSumcheckSession(backend, dist, table, shard, eq_factor, sum0, ap, la, beta)
        bares = []
        for r in range(log_rows):
            bare = sess.compute_bare_round_poly()
            bares.append(bare)
            sess.process_challenge(challenges[r], bare)
        out = sess.final_column_evals()
        sess.free()
        return bares, out

    for bares, out in run_ranks(world, rank_main):
        for r in range(log_rows):
            assert np.array_equal(bares[r], expected[r]), f"round {r}"
        assert np.array_equal(out, finals)


@pytest.mark.gpu
@pytest.mark.parametrize("world,n_vars,folding,live_cols", [(2, 14, 5, 32), (4, 13, 4, 11), (8, 16, 7, 64)])
def test_gpu_sharded_product_sumcheck_threads(rng, world, n_vars, folding, live_cols):
    from leanmultisig_b200.sharded import ShardedProductSumcheck, shard_of

    chunk = 1 << (n_vars - folding)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: live_cols * chunk] = O.random_field(rng, live_cols * chunk)
    g = world.bit_length() - 1
    low = n_vars - folding - g
    shapes = [(0, n_vars), (3 % (1 << folding), n_vars - folding), (5, low), (1, n_vars - 1), (2, low + 1), (0, low - 1)]
    stmts = [(sel % (1 << (n_vars - m)), O.random_field(rng, (m, 5)), O.random_field(rng, 5)) for sel, m in shapes]
    n_rounds = folding + 3
    challenges = O.random_field(rng, (n_rounds, 5))
    # single-process oracle session over the whole polynomial
    w = np.zeros((1 << n_vars, 5), dtype=np.uint32)
    for sel, pt, sc in stmts:
        O.weights_add_eq(w, sel, pt, sc)
    p, expected = ev, []
    expected.append(O.prod_round(p, w))
    for k in range(n_rounds):
        p, w = O.fold_msb(p, challenges[k]), O.fold_msb(w, challenges[k])
        expected.append(O.prod_round(p, w))

    def rank_main(backend, dist, rank):
        shard = shard_of(ev, n_vars, folding, rank, world).reshape(1 << folding, -1)[:live_cols].reshape(-1)
        sc = ShardedProductSumcheck(backend, dist, shard, n_vars, folding)
        for sel, pt, s in stmts:
            sc.add_eq(sel, pt, s)
        rounds = [sc.round()]
        for k in range(n_rounds):
            if k % 2 == 0:
                rounds.append(sc.fold_round(challenges[k]))
            else:
                sc.fold(challenges[k])
                rounds.append(sc.round())
        tables = sc.read()
        sc.free()
        return rounds, tables

    for rounds, (gp, gw) in run_ranks(world, rank_main):
        for k, ((c0, c2), (e0, e2)) in enumerate(zip(rounds, expected)):
            assert np.array_equal(c0, e0) and np.array_equal(c2, e2), f"round {k}"
        assert np.array_equal(gp, p) and np.array_equal(gw, w)


# ---- the row-sharded COMMIT and WHIR commit + open with the real CudaBackend ------------------------------------------
class TorchLikeThreadDist(ThreadDist):
    """torch.distributed's collectives (the subset CudaBackend / sharded.py call) between the rank-threads of one process:
    tensors live on the one device, so a collective is a synchronise + device-to-device copies around two barriers."""

    class ReduceOp:
        SUM, MIN = "sum", "min"

    def _sync(self):
        import torch

        torch.cuda.synchronize()

    def barrier(self):
        self._sync()
        self.group.barrier.wait()

    def all_gather_object(self, out_list, obj):
        out_list[:] = self.group.exchange(self.rank, obj)

    def all_reduce(self, t, op="sum"):
        import torch

        self._sync()
        parts = self.group.exchange(self.rank, t.clone())
        acc = torch.stack(parts)
        t.copy_(acc.min(dim=0).values if op == "min" else acc.sum(dim=0))
        self._sync()
        self.group.barrier.wait()

    def all_gather_into_tensor(self, out, t):
        import torch

        self._sync()
        parts = self.group.exchange(self.rank, t)
        out.copy_(torch.stack(parts).reshape(out.shape))
        self._sync()
        self.group.barrier.wait()

    def all_to_all_single(self, out, inp):
        self._sync()
        parts = self.group.exchange(self.rank, inp)
        world = self.group.world
        chunk = inp.shape[0] // world
        for q in range(world):
            out[q * chunk:(q + 1) * chunk].copy_(parts[q][self.rank * chunk:(self.rank + 1) * chunk])
        self._sync()
        self.group.barrier.wait()


def run_cuda_ranks(world, fn):
    """fn(CudaBackend, dist, rank) on `world` threads of this process, one library context and one torch stream each"""
    import leanmultisig_b200 as lm
    from leanmultisig_b200.sharded import CudaBackend

    group, results, errors = ThreadGroup(world), [None] * world, []

    def body(rank):
        try:
            ctx = lm.Context(0, 20)
            try:
                backend = CudaBackend(ctx)
                backend.scatter_dft = None  # CUDA IPC cannot map a buffer into its own process: all-to-all exchange path
                results[rank] = fn(backend, TorchLikeThreadDist(group, rank), rank)
                backend.torch.cuda.synchronize()
            finally:
                ctx.close()
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            group.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.gpu
@pytest.mark.parametrize("world,n_vars,folding,live_cols", [(2, 13, 4, 16), (4, 15, 5, 20), (8, 19, 7, 64)])
def test_gpu_sharded_commit_threads(rng, world, n_vars, folding, live_cols):
    """ShardedCommit (local transform, exchange, last layers, subtree forest, root all-gather, replicated top) with the
    CUDA backend on a single device: root, every rank's codeword runs and openings equal the single-process oracle commit"""
    from leanmultisig_b200.sharded import ShardedCommit, shard_of

    chunk = 1 << (n_vars - folding)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: live_cols * chunk] = O.random_field(rng, live_cols * chunk)
    cw = O.reorder_and_dft(ev, n_vars, 1, folding, 1, live_cols)
    layers = O.merkle_tree(cw, 1 << folding, live_cols)
    h = cw.shape[0]
    queries = [0, 1, h // 2 + 3, h - 1] + [int(x) for x in rng.integers(0, h, 6)]

    def fn(backend, dist, rank):
        sc = ShardedCommit(backend, dist, n_vars, folding, 1, live_cols=live_cols)
        shard = shard_of(ev, n_vars, folding, rank, world).reshape(1 << folding, -1)[:live_cols].reshape(-1)
        root = sc.commit(backend.to_device(shard))
        mine = backend.to_host(sc.codeword)
        geo = sc.geo
        for m in range(world):  # block m of the local matrix = global rows m * block + rank * run ...
            lo = m * geo.block + rank * geo.run
            assert np.array_equal(mine[m * geo.run:(m + 1) * geo.run], cw[lo:lo + geo.run]), f"rank {rank} run {m}"
        opened = {i: sc.open_local(i) for i in queries if geo.owner(i) == rank}
        dist.barrier()
        return np.asarray(root), opened

    out = run_cuda_ranks(world, fn)
    for root, opened in out:
        assert np.array_equal(root, layers[-1])
        for i, (row, path) in opened.items():
            orow, opath = O.merkle_open(cw, 1 << folding, layers, i)
            assert np.array_equal(row, orow) and np.array_equal(path, opath)
    assert sorted(i for _, opened in out for i in opened) == sorted(set(queries))


@pytest.mark.gpu
@pytest.mark.parametrize("world,nv,live_frac_16", [(2, 12, 16), (4, 13, 8)])
def test_gpu_sharded_whir_threads(world, nv, live_frac_16):
    """ShardedWhirProver (sharded commit, OOD evaluation from the shards, sharded product sumcheck, owner-routed STIR
    openings, replicated tail) with the CUDA backend on a single device: transcript, hints and final point equal the
    single-process oracle prover's, the oracle verifier accepts"""
    import leanmultisig_b200 as lm
    from leanmultisig_b200 import whir_config as WC
    from leanmultisig_b200.sharded import ShardedWhirProver, shard_of
    from test_whir_protocol import SMALL, make_statements, oracle_prove, oracle_verify, to_product_statements

    rs = np.random.default_rng(31 + nv + live_frac_16)
    cfg_o, cfg_p = W.WhirConfig(nv, **SMALL), WC.WhirConfig(nv, **SMALL)
    k = cfg_p.first_folding
    live_cols = (1 << k) * live_frac_16 // 16
    live = live_cols << (nv - k)
    poly = O.random_field(rs, 1 << nv)
    poly[live:] = 0
    stm = make_statements(rs, poly, nv, with_next=True)
    ps_o, point_o = oracle_prove(cfg_o, poly, stm, live)

    def fn(backend, dist, rank):
        ps = lm.ProverState(backend.ctx)
        shard = shard_of(poly, nv, k, rank, world).reshape(1 << k, -1)[:live_cols].reshape(-1)
        prover = ShardedWhirProver(backend, dist, cfg_p)
        wit = prover.commit(ps, shard, live_cols)
        point = prover.prove(ps, to_product_statements(stm), wit)
        dist.barrier()
        return ps.transcript, ps.merkle_paths, point

    for transcript, paths, point in run_cuda_ranks(world, fn):
        assert transcript == ps_o.transcript and point == point_o
        assert len(paths) == len(ps_o.merkle_paths)
        for ga, oa in zip(paths, ps_o.merkle_paths):
            for (gl, gp, gi), (ol, op, oi) in zip(ga, oa):
                assert gi == oi and np.array_equal(gl, ol) and np.array_equal(gp, op)
        assert oracle_verify(cfg_o, transcript, paths, stm) == point

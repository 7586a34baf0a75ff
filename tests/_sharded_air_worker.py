"""Worker for tests/test_sharded_air.py (run under torch.distributed.run, gloo on CPU or nccl on GPUs).

Every rank builds the same full table from a seed, keeps its row range, and drives
leanmultisig_b200.sharded.ShardedAirSumcheckSession; the per-round bare polynomials and the final column evaluations
must equal those of the single-process oracle session over the whole table.  CPU mode runs the product's orchestration
with an oracle-backed compute double, GPU mode with the CUDA backend."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from leanmultisig_b200 import field as F  # noqa: E402
from leanmultisig_b200.air import AIR_SHAPES, OuterSumcheckHost  # noqa: E402
from leanmultisig_b200.sharded import P, ShardedAirSumcheckSession  # noqa: E402


class OracleAirSession:
    """one rank's compute on the oracle (test double for AirSumcheckSession in the gloo tier)"""

    def __init__(self, table_id, columns, eq_factor, ap, la, beta, halo_next_row=None, eq_scale=None, folded_columns=None):
        self.table, self.args = table_id, (ap, la, beta)
        _, n_shift, _ = O.air_shape(table_id)
        if folded_columns is not None:
            self.cur = np.ascontiguousarray(folded_columns, dtype=np.uint32)
        else:
            cols = np.stack(columns)
            shifted = [O.shift_column(cols[k]) for k in range(n_shift)]
            if halo_next_row is not None:
                for k in range(n_shift):
                    shifted[k][-1] = halo_next_row[k]
            self.cur = np.concatenate([cols, np.stack(shifted)]) if n_shift else cols
        self.eq = np.ascontiguousarray(eq_factor, dtype=np.uint32).reshape(-1, 5)
        self.scale = eq_scale

    def _raw_round(self):
        n_vars = self.cur.shape[1].bit_length() - 1
        raw = O.air_round(self.table, self.cur, self.eq[: n_vars - 1], *self.args)
        if self.scale is not None:
            raw = np.stack([O.ef_mul(v, self.scale) for v in raw])
        return raw

    def _fold(self, ch):
        self.cur = np.stack([O.fold_lsb(self.cur[c], ch) for c in range(self.cur.shape[0])])

    def final_column_evals(self):
        assert self.cur.shape[1] == 1
        return self.cur[:, 0, :]

    def free(self):
        pass


class OracleBackend:
    def air_session(self, table_id, columns, eq_factor, ap, la, beta, **kw):
        return OracleAirSession(table_id, columns, eq_factor, ap, la, beta, **kw)

    def all_reduce_field(self, d, words):
        t = torch.from_numpy(np.ascontiguousarray(words).astype(np.int64))
        d.all_reduce(t)
        return (t.numpy() % P).astype(np.uint32).reshape(words.shape)

    def all_gather_words(self, d, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        outs = [torch.empty(w.shape, dtype=torch.int32) for _ in range(d.get_world_size())]
        d.all_gather(outs, torch.from_numpy(w.view(np.int32)))
        return np.stack([o.numpy().view(np.uint32) for o in outs])


class SingleOracleSession(OuterSumcheckHost):
    """the whole table in one process: oracle compute + the same host round logic"""

    def __init__(self, table_id, columns, eq_factor, sum_, ap, la, beta):
        self.inner = OracleAirSession(table_id, columns, eq_factor, ap, la, beta)
        n_vars = columns[0].size.bit_length() - 1
        self._init_host(eq_factor, sum_, n_vars, O.air_shape(table_id)[2])

    def _raw_round(self):
        return self.inner._raw_round()

    def _fold(self, ch):
        self.inner._fold(ch)

    def final_column_evals(self):
        return self.inner.final_column_evals()


def main():
    mode, table_id, log_rows = sys.argv[1], int(sys.argv[2], 0), int(sys.argv[3])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if mode == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_cols, n_shift, degree = O.air_shape(table_id)
    assert AIR_SHAPES[table_id & 0xFF] == (n_cols, n_shift, degree)
    rng = np.random.default_rng(7 + table_id + log_rows)
    n = 1 << log_rows
    cols = O.random_field(rng, (n_cols, n))
    eq_factor = O.random_field(rng, (log_rows, 5))
    alpha = O.random_field(rng, 5)
    ap = [np.array([int(O.to_monty(1)), 0, 0, 0, 0], dtype=np.uint32)]
    for _ in range(100):
        ap.append(O.ef_mul(ap[-1], alpha))
    ap, la, beta = np.stack(ap), O.random_field(rng, (6, 5)), O.random_field(rng, 5)
    sum0 = O.random_field(rng, 5)
    challenges = O.random_field(rng, (log_rows, 5))

    if mode == "gpu":
        import leanmultisig_b200 as lm
        from leanmultisig_b200.sharded import CudaBackend

        ctx = lm.Context(local_rank, 20)
        backend = CudaBackend(ctx)
    else:
        backend = OracleBackend()
    per = n // world
    shard = [cols[c, rank * per:(rank + 1) * per] for c in range(n_cols)]
    sess = ShardedAirSumcheckSession(backend, dist, table_id, shard, eq_factor, sum0, ap, la, beta)
    ref = SingleOracleSession(table_id, list(cols), eq_factor, sum0, ap, la, beta)
    assert sess.initial_n_vars() == log_rows and sess.bare_degree() == degree
    for r in range(log_rows):
        bare, exp = sess.compute_bare_round_poly(), ref.compute_bare_round_poly()
        assert np.array_equal(bare, exp), f"rank {rank}: round {r} differs"
        assert np.array_equal(sess.sum(), ref.sum())
        sess.process_challenge(challenges[r], bare)
        ref.process_challenge(challenges[r], exp)
    assert np.array_equal(sess.final_column_evals(), ref.final_column_evals()), f"rank {rank}: final column evaluations differ"
    # the final values are the MLEs of the (shifted) columns at the reversed challenge point
    assert np.array_equal(sess.final_column_evals()[0], O.mle_eval(cols[0], challenges[::-1]))
    sess.free()
    dist.barrier()
    if rank == 0:
        print("SHARDED_AIR_OK", world, mode)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Worker for tests/test_sharded_commit.py (run under torch.distributed.run, gloo on CPU or nccl on GPUs).

CPU mode runs the product's orchestration (leanmultisig_b200.sharded.ShardedCommit) with an oracle-backed compute
backend; GPU mode uses the CUDA backend.  Both compare with the single-process oracle commit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from leanmultisig_b200.sharded import ShardedCommit, ShardGeometry, shard_of  # noqa: E402


class OracleBackend:
    """numpy/oracle stand-in for the CUDA backend (test double for the gloo tier)."""

    def to_device(self, a):
        return np.ascontiguousarray(a, dtype=np.uint32)

    def to_host(self, t):
        return np.asarray(t)

    def empty_like(self, t):
        return np.empty_like(t)

    def rows(self, t, start, count):
        return t[start:start + count]

    def take_rows(self, t, idx):
        return np.asarray(t)[np.asarray(idx, dtype=np.int64)]

    def reorder_and_dft(self, shard, n_vars, folding, log_inv_rate, cols):
        return O.reorder_and_dft(shard, n_vars, 1, folding, log_inv_rate, cols)

    def all_to_all(self, d, out, inp):
        o, i = torch.from_numpy(out.view(np.int32)), torch.from_numpy(inp.view(np.int32))
        d.all_to_all_single(o, i)

    def dft_layers_mapped(self, mat, w, log_h, l_first, n_blocks, run, block, offset):
        mat[...] = O.dft_layers_mapped(mat, log_h, l_first, n_blocks, run, block, offset)

    def merkle_tree(self, rows, full_cols, eff_cols):
        return O.merkle_tree(rows, full_cols, eff_cols)

    def all_gather_roots(self, d, my_roots):
        world = d.get_world_size()
        outs = [torch.empty(my_roots.shape, dtype=torch.int32) for _ in range(world)]
        d.all_gather(outs, torch.from_numpy(np.ascontiguousarray(my_roots).view(np.int32)))
        return np.stack([o.numpy().view(np.uint32) for o in outs])

    def permute_roots(self, all_roots):
        return np.ascontiguousarray(all_roots.transpose(1, 0, 2)).reshape(-1, 8)

    def merkle_levels(self, layer0):
        n = layer0.shape[0]
        layers = [layer0]
        cur = layer0
        while cur.shape[0] > 1:
            st = np.concatenate([cur[0::2], cur[1::2]], axis=1)
            cur = O.poseidon1_compress(st)[:, :8]
            layers.append(cur)
        return np.concatenate(layers)


def main():
    mode = sys.argv[1]
    n_vars, folding, rate, live_cols = (int(x) for x in sys.argv[2:6])
    if mode == "gpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(42)
    cols = 1 << folding
    chunk = 1 << (n_vars - folding)
    ev = np.zeros(1 << n_vars, dtype=np.uint32)
    ev[: live_cols * chunk] = O.random_field(rng, live_cols * chunk)
    # single-process reference
    cw_ref = O.reorder_and_dft(ev, n_vars, 1, folding, rate, live_cols)
    layers_ref = O.merkle_tree(cw_ref, cols, live_cols)
    root_ref = layers_ref[-1]
    shard = shard_of(ev, n_vars, folding, rank, world).reshape(cols, -1)[:live_cols].reshape(-1)
    if mode == "gpu":
        import leanmultisig_b200 as lm
        from leanmultisig_b200.sharded import CudaBackend

        ctx = lm.Context(int(os.environ.get("LOCAL_RANK", "0")), 24)
        backend = CudaBackend(ctx)
    else:
        backend = OracleBackend()
    sc = ShardedCommit(backend, dist, n_vars, folding, rate, live_cols=live_cols)
    root = sc.commit(backend.to_device(shard))
    assert np.array_equal(np.asarray(root).view(np.uint32).reshape(-1), root_ref), f"rank {rank}: root differs"
    geo = ShardGeometry(n_vars, folding, rate, world)
    local = backend.to_host(sc.codeword)
    for m in range(world):
        g0 = m * geo.block + rank * geo.run
        assert np.array_equal(local[m * geo.run:(m + 1) * geo.run], cw_ref[g0:g0 + geo.run]), f"rank {rank}: run {m} differs"
    log_h = geo.log_h
    for row in [r for r in (0, 1, geo.run, geo.h // 2 + 3, geo.h - 1, 5 * geo.run + 1) if r < geo.h and geo.owner(r) == rank]:
        data, path = sc.open_local(row)
        er, ep = O.merkle_open(cw_ref, cols, layers_ref, row)
        assert np.array_equal(data, er) and np.array_equal(path, ep), f"rank {rank}: opening {row} differs"
        assert O.merkle_verify(root_ref, log_h, row, data, path)
    # the host-input path (pipelined over column groups on the CUDA backend, plain upload + commit otherwise).  The
    # codeword matrix of the CUDA backend is written by the OTHER ranks during a commit: every rank must be done reading
    # the previous codeword before anyone starts the next commit.
    dist.barrier()
    if mode == "gpu":
        host = torch.from_numpy(np.ascontiguousarray(shard).view(np.int32)).pin_memory()
        root2 = sc.commit_host(host)
    else:
        root2 = sc.commit_host(np.ascontiguousarray(shard))
    assert np.array_equal(np.asarray(root2).view(np.uint32).reshape(-1), root_ref), f"rank {rank}: root of the host-input path differs"
    local2 = backend.to_host(sc.codeword)
    for m in range(world):
        g0 = m * geo.block + rank * geo.run
        assert np.array_equal(local2[m * geo.run:(m + 1) * geo.run], cw_ref[g0:g0 + geo.run]), f"rank {rank}: run {m} differs (host path)"
    row = rank * geo.run + 1
    data, path = sc.open_local(row)
    er, ep = O.merkle_open(cw_ref, cols, layers_ref, row)
    assert np.array_equal(data, er) and np.array_equal(path, ep), f"rank {rank}: opening {row} differs (host path)"
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK", world, mode)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Row-sharded WHIR commit and row-sharded AIR sumcheck across the GPUs of one node.

One process per GPU (`torch.distributed`, NCCL over NVLink; gloo in the CPU tests).  The stacked polynomial has
index (column c | position s); rank q holds, for every column, the slice of s whose top g = log2(world) bits equal q.

  1. local      T_q = reorder_and_dft of the shard (a polynomial of n_vars - g variables, same folding and rate):
                the first log2(h / G) butterfly layers of the global transform act independently on G row blocks
                and block q is exactly rank q's shard.
  2. all-to-all rank q sends rows [q' run, (q'+1) run) of T_q to rank q' (run = h / G^2): (G-1)/G of its block.
  3. combine    the last g layers pair rows that differ only in the block index m, now all local
                (lm_dev_dft_layers_mapped).  Rank q' ends up with G runs of `run` consecutive codeword rows.
  4. Merkle     leaf sponge over all local rows and the G subtrees (one per run) level by level as one forest, all-gather
                of the G^2 subtree roots (32 B each), the top 2g levels are replicated.

AIR sumcheck (`ShardedAirSumcheckSession`): rank q holds rows [q 2^(n-g), (q+1) 2^(n-g)) of every column of a table.  The
session folds the least-significant row bit first (air_sumcheck.rs:144-151), so the first n - g rounds touch only local
rows: every rank computes the round sums of its shard with the eq weight of its row prefix already multiplied in
(lm_air_new_shard), ONE all-reduce of degree x 5 field words per round adds them up, and the fold is local.  The shifted
("next row") columns need the first row of the next shard once, at setup (compute_shifted_columns, :683-694).  After
n - g rounds every rank is left with one EF value per column: an all-gather of (n_cols + n_shift) x 5 words builds the
2^g-row table on which the last g rounds run replicated (lm_air_new_folded).

Quotient GKR (`ShardedGkrQuotientProver`): the fraction table is split the same way.  The up pass pairs adjacent rows
and is local down to 2^(5-g) fractions per rank; an all-gather of those gives the 2^5 top values the prover sends.  Every
layer sumcheck folds the least-significant variable first: k - g local rounds with ONE all-reduce of (c0, c2) = 10 field
words each, then an all-gather of the four folded values per rank and the last g rounds on 4 x G values on the host.

WHIR product sumcheck (`ShardedProductSumcheck`): the first `folding` rounds fold the column bits (most significant
first, poly/src/utils.rs:161-186), which every rank holds completely: local rounds with ONE all-reduce of (c0, c2) each.
The next variables are the sharding bits; by then the tables are 2^(n - folding) EF entries in total, so one all-gather
(device to device) rebuilds them on every rank and the remaining rounds, the STIR updates and the round commitments run
replicated on the ordinary single-GPU session.

The compute steps go through a backend object so that the CPU test tier can run the same orchestration with the
oracle over gloo; the product backend is the CUDA library (`CudaBackend`), there is no CPU product path.
"""
from __future__ import annotations

import os

import numpy as np

from . import field as F
from .air import AIR_SHAPES, OuterSumcheckHost
from .logup import N_VARS_TO_SEND_GKR_COEFFS, GkrQuotientProver

P = 0x7F000001


def shard_of(evals: np.ndarray, n_vars: int, folding: int, rank: int, world: int) -> np.ndarray:
    """The part of the (host) polynomial rank `rank` owns, laid out as a polynomial of n_vars - g variables."""
    g = world.bit_length() - 1
    assert world == 1 << g and n_vars - folding >= 2 * g
    cols = 1 << folding
    chunk = 1 << (n_vars - folding)
    sub = chunk >> g
    return np.ascontiguousarray(evals.reshape(cols, chunk)[:, rank * sub:(rank + 1) * sub]).reshape(-1)


class ShardGeometry:
    def __init__(self, n_vars: int, folding: int, log_inv_rate: int, world: int):
        self.world = world
        self.g = world.bit_length() - 1
        assert world == 1 << self.g, "world size must be a power of two"
        self.n_vars, self.folding, self.log_inv_rate = n_vars, folding, log_inv_rate
        self.log_h = n_vars + log_inv_rate - folding
        self.h = 1 << self.log_h
        self.block = self.h >> self.g          # rows of one rank's local transform
        self.run = self.block >> self.g        # consecutive codeword rows per (rank, block) pair
        assert self.run >= 2, "domain too small for this many ranks"
        self.log_run = self.run.bit_length() - 1

    def owner(self, row: int) -> int:
        return (row // self.run) % self.world

    def local_row(self, row: int) -> int:
        """index of a global codeword row inside its owner's local matrix ((m, j') order)"""
        m, rest = divmod(row, self.block)
        return m * self.run + rest % self.run

    def subtree_index(self, rank: int, m: int) -> int:
        """position of rank's m-th subtree root in the global layer of G^2 roots"""
        return m * self.world + rank


class ShardedCommit:
    """Collective: every rank calls commit() with its shard; all ranks return the same root.

    With the CUDA backend's fused exchange the matrix the OTHER ranks store into during a commit is one buffer per shape,
    reused by the next commit of that shape; `self.codeword` is a private copy taken before the commit's closing collective,
    so a witness stays valid however many commits of the same shape follow (round-1 advisor finding)."""

    def __init__(self, backend, dist, n_vars: int, folding: int, log_inv_rate: int, live_cols: int | None = None):
        self.b, self.dist = backend, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.geo = ShardGeometry(n_vars, folding, log_inv_rate, self.world)
        self.cols = (1 << folding) if live_cols is None else live_cols
        self.full_cols = 1 << folding

    def commit(self, shard):
        geo, b = self.geo, self.b
        w = self.cols
        mark = getattr(b, "mark", None) or (lambda name: None)  # optional per-phase device timestamps (CudaBackend)
        mark("start")
        scatter = getattr(b, "scatter_dft", None) if self.world > 1 else None
        if scatter is not None and not b.scatter_ready(self.dist, geo.block, w):
            scatter = None
        if scatter is not None:
            # 1 + 2 fused: the last pass of the local transform stores every row into the matrix of the rank that owns it
            # after the exchange, through peer pointers over NVLink; a stream-ordered barrier closes the exchange
            mat = scatter(self.dist, shard, geo.n_vars - geo.g, geo.folding, geo.log_inv_rate, w)
            mark("local_dft")
            mark("all_to_all")
        else:
            # 1. local transform of the shard
            t_local = b.reorder_and_dft(shard, geo.n_vars - geo.g, geo.folding, geo.log_inv_rate, w)  # block x w
            mark("local_dft")
            # 2. exchange: equal splits of `run` rows, received in block order m = source rank
            mat = b.empty_like(t_local)
            if self.world > 1:
                b.all_to_all(self.dist, mat, t_local)
            else:
                mat = t_local
            mark("all_to_all")
        # 3. last g layers on the local rows.  With the fused exchange `mat` is the buffer the peers store the NEXT commit of
        #    this shape into, so the pass writes its result to a private matrix of the witness (same traffic as in place, no
        #    copy); it is ordered before this commit's root all-gather, which no peer can pass - and hence start its next
        #    scatter - before this rank has reached it (round-1 advisor finding)
        if scatter is not None:
            priv = b.empty_like(mat)
            b.dft_layers_mapped(mat, w, geo.log_h, geo.log_h - geo.g, self.world, geo.run, geo.block, self.rank * geo.run, out=priv)
            mat = priv
        elif geo.g:
            b.dft_layers_mapped(mat, w, geo.log_h, geo.log_h - geo.g, self.world, geo.run, geo.block, self.rank * geo.run)
        mark("last_layers")
        self.codeword = mat
        # 4. the G subtrees of this rank as ONE forest: leaf digests of all local rows in one launch, then level by level
        #    over the concatenated runs (a run is a power-of-two block of the local matrix, so level l of the forest is the
        #    concatenation of level l of every subtree); the levels above log2(run) are not used
        self.forest = b.merkle_tree(mat, self.full_cols, w)          # (2 block - 1) x 8, level-major
        mark("subtrees")
        return self._finish(mark)

    def commit_host(self, host_shard):
        """commit() for a shard in (pinned) HOST memory.  Backends that can (CudaBackend with the fused exchange) pipeline
        it over column groups taken right to left, as lm_commit does on one GPU: while group k is transformed, exchanged
        and absorbed by the leaf sponge, the copy of group k+1 crosses PCIe.  Otherwise: upload, then commit()."""
        pipelined = getattr(self.b, "commit_host_pipelined", None) if self.world > 1 else None
        if pipelined is not None:
            out = pipelined(self, host_shard)
            if out is not None:
                self.codeword, self.forest = out
                return self._finish(lambda name: None)
        return self.commit(self.b.to_device(host_shard) if isinstance(host_shard, np.ndarray) else host_shard.cuda(non_blocking=True))

    def _finish(self, mark):
        geo, b = self.geo, self.b
        off = 2 * geo.block - ((2 * geo.block) >> geo.log_run)
        my_roots = b.rows(self.forest, off, self.world)              # world x 8: level log2(run) = the subtree roots
        all_roots = b.all_gather_roots(self.dist, my_roots)          # [rank][m] -> world x world x 8
        top_layer0 = b.permute_roots(all_roots)                      # index m * world + rank
        self.top = b.merkle_levels(top_layer0)                       # (2 G^2 - 1) x 8
        self.root = b.to_host(self.top)[-1]
        mark("top")
        return self.root

    def open_local(self, row: int):
        """Opening of a codeword row this rank owns: (row zero-extended, sibling path leaf level first)."""
        geo, b = self.geo, self.b
        assert geo.owner(row) == self.rank
        m, rest = divmod(row, geo.block)
        jp = rest % geo.run
        path = []
        off, n, idx = 0, geo.block, m * geo.run + jp
        for _ in range(geo.log_run):
            path.append(b.to_host(b.rows(self.forest, off + (idx ^ 1), 1)).reshape(-1))
            off += n
            n >>= 1
            idx >>= 1
        top = b.to_host(self.top)
        off, n, idx = 0, self.world * self.world, geo.subtree_index(self.rank, m)
        while n > 1:
            path.append(top[off + (idx ^ 1)])
            off += n
            n >>= 1
            idx >>= 1
        data = b.to_host(b.rows(self.codeword, m * geo.run + jp, 1)).reshape(-1)
        full = np.zeros(self.full_cols, dtype=np.uint32)
        full[: data.size] = data
        return full, np.stack(path)


def _open_local_batch(sc, rows):
    """open_local for several rows this rank owns with ONE device gather + read-back per array (forest siblings, codeword rows)
    and one download of the replicated top layers — the single-GPU path's lm_open does the same with one gather kernel; a
    proof opens ~230 rows of the first tree.  -> [(row zero-extended, path)] in the order of `rows`."""
    geo, b = sc.geo, sc.b
    take = getattr(b, "take_rows", None)
    if take is None or not rows:
        return [sc.open_local(r) for r in rows]
    f_idx, c_idx, tops = [], [], []
    for row in rows:
        assert geo.owner(row) == sc.rank
        m, rest = divmod(row, geo.block)
        jp = rest % geo.run
        off, n, idx = 0, geo.block, m * geo.run + jp
        c_idx.append(idx)
        for _ in range(geo.log_run):
            f_idx.append(off + (idx ^ 1))
            off += n
            n >>= 1
            idx >>= 1
        tops.append(geo.subtree_index(sc.rank, m))
    forest_rows = take(sc.forest, f_idx).reshape(len(rows), geo.log_run, 8) if geo.log_run else np.zeros((len(rows), 0, 8), np.uint32)
    code_rows = take(sc.codeword, c_idx)
    top = b.to_host(sc.top)
    out = []
    for q in range(len(rows)):
        path = [forest_rows[q]]
        off, n, idx = 0, sc.world * sc.world, tops[q]
        tail = []
        while n > 1:
            tail.append(top[off + (idx ^ 1)])
            off += n
            n >>= 1
            idx >>= 1
        if tail:
            path.append(np.stack(tail))
        full = np.zeros(sc.full_cols, dtype=np.uint32)
        full[: code_rows.shape[1]] = code_rows[q]
        out.append((full, np.concatenate(path, axis=0)))
    return out


class ShardedAirSumcheckSession(OuterSumcheckHost):
    """trait OuterSumcheckSession (air_sumcheck.rs:34-42) over one row-range shard per rank.  Collective: every rank
    constructs it with its rows and then makes the same calls with the same challenges (the transcript is replicated);
    `prove_batched_air_sumcheck` drives it unchanged."""

    def __init__(self, backend, dist, table_id: int, shard_columns, eq_factor, sum_, alpha_powers, logup_alphas_eq_poly,
                 bus_beta):
        self.b, self.dist = backend, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        g = self.world.bit_length() - 1
        assert self.world == 1 << g, "world size must be a power of two"
        n_cols, n_shift, degree = AIR_SHAPES[table_id & 0xFF]
        cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in shard_columns]
        assert len(cols) == n_cols
        local_rows = cols[0].size
        self.local_vars = local_rows.bit_length() - 1
        assert local_rows == 1 << self.local_vars and self.local_vars >= 1, "a shard needs at least two rows"
        self.g = g
        eq = np.ascontiguousarray(eq_factor, dtype=np.uint32).reshape(-1, 5)
        assert eq.shape[0] == self.local_vars + g
        self._args = (table_id, np.ascontiguousarray(alpha_powers, dtype=np.uint32).reshape(-1, 5),
                      np.ascontiguousarray(logup_alphas_eq_poly, dtype=np.uint32).reshape(-1, 5),
                      np.ascontiguousarray(bus_beta, dtype=np.uint32))
        self._eq_top = eq[:g]
        # one-row halo: first row of the shifted columns' sources on the next rank
        halo = None
        if n_shift and self.world > 1:
            first = np.array([c[0] for c in cols[:n_shift]], dtype=np.uint32)
            everyone = backend.all_gather_words(dist, first)
            if self.rank + 1 < self.world:
                halo = everyone[self.rank + 1]
        # eq value of this rank's row prefix: the top g variables are the bits of the rank, most significant first
        scale = F.ONE
        for k in range(g):
            e = F.from_monty(eq[k])
            scale = F.mul(scale, e if (self.rank >> (g - 1 - k)) & 1 else F.sub(F.ONE, e))
        self.local = backend.air_session(table_id, cols, eq[g:], *self._args[1:], halo_next_row=halo, eq_scale=F.to_monty(scale))
        self.tail = None
        self._init_host(eq, sum_, self.local_vars + g, degree)

    def _raw_round(self) -> np.ndarray:
        if self.rounds_done < self.local_vars:
            raw = self.local._raw_round()
            return self.b.all_reduce_field(self.dist, raw) if self.world > 1 else raw
        return self.tail._raw_round()

    def _fold(self, challenge) -> None:
        if self.rounds_done < self.local_vars:
            self.local._fold(challenge)
            if self.rounds_done + 1 == self.local_vars and self.g:
                mine = self.local.final_column_evals()                         # (n_cols + n_shift) x 5
                table = self.b.all_gather_words(self.dist, mine)               # world x (n_cols + n_shift) x 5
                folded = np.ascontiguousarray(table.transpose(1, 0, 2))        # column x row (= rank) x 5
                self.tail = self.b.air_session(self._args[0], None, self._eq_top, *self._args[1:], folded_columns=folded)
        else:
            self.tail._fold(challenge)

    def final_column_evals(self) -> np.ndarray:
        return (self.tail if self.g else self.local).final_column_evals()

    def free(self):
        for s in (self.local, self.tail):
            if s is not None:
                s.free()
        self.local = self.tail = None


def prefix_eq(point, g: int, rank: int):
    """eq(point[:g], bits of rank), most significant bit first; point: list of EF tuples"""
    scale = F.ONE
    for k in range(g):
        scale = F.mul(scale, point[k] if (rank >> (g - 1 - k)) & 1 else F.sub(F.ONE, point[k]))
    return scale


def _eq_table_small(point) -> list:
    """[eq(point, j) for j < 2^len(point)], x_0 = most significant bit of j"""
    table = [F.ONE]
    for x in point:
        table = [v for t in table for v in (F.mul(t, F.sub(F.ONE, x)), F.mul(t, x))]
    return table


class ShardedGkrQuotientProver(GkrQuotientProver):
    """prove_gkr_quotient (quotient_gkr/mod.rs:31-141) over a fraction table split into one row range per rank.
    Collective: every rank constructs it with its rows (`nums`, `dens` = the active part of rows
    [rank 2^(n_vars-g), (rank+1) 2^(n_vars-g)), possibly empty) and drives prove() with the same transcript."""

    def __init__(self, backend, dist, nums, dens, n_vars: int):
        self.b, self.dist = backend, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.g = self.world.bit_length() - 1
        assert self.world == 1 << self.g and self.g < N_VARS_TO_SEND_GKR_COEFFS
        assert n_vars - self.g > N_VARS_TO_SEND_GKR_COEFFS, "table too small for this many ranks"
        self.n_vars = n_vars
        self.local = backend.gkr_session(nums, dens, n_vars - self.g, N_VARS_TO_SEND_GKR_COEFFS - self.g)
        self.handle = None

    def top(self):
        tn, td = self.local.top()
        if self.world == 1:
            return tn, td
        both = self.b.all_gather_words(self.dist, np.stack([tn, td]))      # world x 2 x 2^(5-g) x 5
        return (np.ascontiguousarray(both[:, 0]).reshape(-1, 5), np.ascontiguousarray(both[:, 1]).reshape(-1, 5))

    def _layer_begin(self, k, point_m, alpha_m):
        g = self.g
        point = [F.from_monty(x) for x in point_m]
        self._local_rounds, self._round_no = k - g, 0
        self._top_point, self._alpha, self._tail = point[:g], F.from_monty(alpha_m), None
        self.local.layer_begin(k - g, point_m[g:], alpha_m, F.to_monty(prefix_eq(point, g, self.rank)))

    def _round(self):
        if self._round_no < self._local_rounds:
            c0, c2 = self.local.round()
            if self.world == 1:
                return c0, c2
            both = self.b.all_reduce_field(self.dist, np.stack([c0, c2]))
            return both[0], both[1]
        # tail: 4 columns of 2^m values on the host; G(nl, nr, dl, dr) = nl dr + nr dl + alpha dl dr  (sumcheck_utils.rs:65-79)
        nl, nr, dl, dr = self._tail
        m = len(nl).bit_length() - 1
        eq = _eq_table_small(self._top_point[: m - 1])
        al = self._alpha

        def big_g(a, b, c, d):
            return F.add(F.add(F.mul(a, d), F.mul(b, c)), F.mul(al, F.mul(c, d)))

        c0 = c2 = F.ZERO
        for j, e in enumerate(eq):
            lo = [col[2 * j] for col in self._tail]
            df = [F.sub(col[2 * j + 1], col[2 * j]) for col in self._tail]
            c0 = F.add(c0, F.mul(e, big_g(*lo)))
            c2 = F.add(c2, F.mul(e, big_g(*df)))
        return F.to_monty(c0), F.to_monty(c2)

    def _fold(self, r_m):
        if self._round_no < self._local_rounds:
            self.local.fold(r_m)
            if self._round_no + 1 == self._local_rounds and self.g:
                mine = self.local.layer_end()                                # 4 x 5
                everyone = self.b.all_gather_words(self.dist, mine)           # world x 4 x 5, row index = rank
                self._tail = [[F.from_monty(everyone[q, c]) for q in range(self.world)] for c in range(4)]
        else:
            r = F.from_monty(r_m)
            self._tail = [[F.add(col[2 * j], F.mul(r, F.sub(col[2 * j + 1], col[2 * j]))) for j in range(len(col) // 2)]
                          for col in self._tail]
        self._round_no += 1

    def _layer_end(self):
        if self.g == 0:
            return self.local.layer_end()
        return np.stack([F.to_monty(col[0]) for col in self._tail])

    def free(self):
        if self.local is not None:
            self.local.free()
            self.local = None


def localize_statement(n_vars: int, folding: int, g: int, rank: int, selector: int, point):
    """Restriction of the weight statement  w[x] += s * eq(point, x_inner) * [x_outer == selector]  (x = outer | inner,
    inner = the low m = len(point) bits) to rank's shard.  Global index bits, most significant first: column (folding
    bits), rank (g bits), position inside the shard.  Returns (local selector, local point, eq factor of the rank bits)
    or None when the selector excludes this rank."""
    m = len(point)
    low = n_vars - folding - g                      # bits of the position inside the shard
    keep, scale, sel = list(range(m)), F.ONE, selector
    drop_sel = []
    for t in range(g):                              # rank bit t sits at global bit position low + t
        pos, bit = low + t, (rank >> t) & 1
        if pos < m:                                 # inside the eq part: coordinate m - 1 - pos
            x = point[m - 1 - pos]
            scale = F.mul(scale, x if bit else F.sub(F.ONE, x))
            keep.remove(m - 1 - pos)
        else:                                       # inside the selector: bit pos - m
            if ((selector >> (pos - m)) & 1) != bit:
                return None
            drop_sel.append(pos - m)
    for b in sorted(drop_sel, reverse=True):
        sel = ((sel >> (b + 1)) << b) | (sel & ((1 << b) - 1))
    return sel, [point[j] for j in keep], scale


class ShardedProductSumcheck:
    """SumcheckSingle (crates/whir/src/open.rs:323-446) over the row-sharded stacked polynomial.  Collective: every rank
    builds it from its shard (`shard_of`) and makes the same calls with the same challenges.  Statements (`add_eq`) must
    be added before the first round; `add_base_eq`, `read`, `eval_poly`, `commit_poly` are available once the tables are
    replicated, i.e. after `folding` folds."""

    def __init__(self, backend, dist, shard_evals, n_vars: int, folding: int, live_len: int | None = None):
        self.b, self.dist = backend, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.g = self.world.bit_length() - 1
        assert self.world == 1 << self.g and n_vars - folding >= 2 * self.g and folding >= 1
        self.n_vars_total, self.folding = n_vars, folding
        self.local = backend.sumcheck(shard_evals, n_vars - self.g, live_len)
        self.rep = None
        self.folds = 0

    @property
    def n_vars(self) -> int:
        return self.n_vars_total - self.folds

    def add_eq(self, selector: int, point, scalar) -> None:
        if self.rep is not None:          # replicated tables (the STIR / OOD updates of the later WHIR rounds)
            return self.rep.add_eq(selector, point, scalar)
        assert self.folds == 0, "statements are added before the first fold"
        pt = [F.from_monty(x) for x in np.ascontiguousarray(point, dtype=np.uint32).reshape(-1, 5)]
        loc = localize_statement(self.n_vars_total, self.folding, self.g, self.rank, selector, pt)
        if loc is None:
            return
        sel, lp, scale = loc
        lpm = np.stack([F.to_monty(x) for x in lp]) if lp else np.zeros((0, 5), dtype=np.uint32)
        self.local.add_eq(sel, lpm, F.to_monty(F.mul(scale, F.from_monty(scalar))))

    def add_next(self, selector: int, point, scalar) -> None:
        """The next-row statement (SparseStatement::new_next, what the stacked-PCS opening emits for every table column,
        sub_protocols/src/stacked_pcs.rs:73-82) on shards.  matrix_next_mle_folded (poly/src/next_mle.rs:35-58) is the sum of
        m + 1 terms; term k lives on the inner indices x = (b << (k + 1)) | (1 << k) with weight
        (1 - oc[m-k-1]) prod_{j >= m-k} oc[j] eq(oc[0 .. m-k-1), b), the last term on x = 2^m - 1 with weight prod_j oc[j].
        On a row-range shard the g index bits that select the rank are constants: a term survives on this rank only if its
        fixed bits agree with the rank's, the coordinates of b that fall on rank bits turn into scalar factors, and the
        remaining index set is again of the strided form `add_strided_eq` takes."""
        if self.rep is not None:
            return self.rep.add_next(selector, point, scalar)
        assert self.folds == 0, "statements are added before the first fold"
        oc = [F.from_monty(x) for x in np.ascontiguousarray(point, dtype=np.uint32).reshape(-1, 5)]
        m, g = len(oc), self.g
        low = self.n_vars_total - self.folding - g
        rank_pos = [(low + t, (self.rank >> t) & 1) for t in range(g)]          # (global bit position, required value)
        sel_local = selector
        for pos, bit in sorted(rank_pos, reverse=True):                          # rank bits inside the selector
            if pos >= m:
                if ((selector >> (pos - m)) & 1) != bit:
                    return
                b = pos - m
                sel_local = ((sel_local >> (b + 1)) << b) | (sel_local & ((1 << b) - 1))
        inner = [(pos, bit) for pos, bit in rank_pos if pos < m]
        m_local = m - len(inner)
        base = sel_local << m_local
        s0 = F.from_monty(scalar)

        def emit(shift, offset, pt, coeff):
            ptm = np.stack([F.to_monty(x) for x in pt]) if pt else np.zeros((0, 5), dtype=np.uint32)
            self.local.add_strided_eq(base, shift, offset, ptm, F.to_monty(coeff))

        for k in range(m):
            pre = m - k - 1
            coeff = F.mul(s0, F.sub(F.ONE, oc[pre]))
            for j in range(pre + 1, m):
                coeff = F.mul(coeff, oc[j])
            keep, ok = list(range(pre)), True
            for pos, bit in inner:
                if pos > k:                      # a bit of b: coordinate oc[m - 1 - pos]
                    c = oc[m - 1 - pos]
                    coeff = F.mul(coeff, c if bit else F.sub(F.ONE, c))
                    keep.remove(m - 1 - pos)
                elif pos == k:                   # the set bit of the pattern
                    ok = ok and bit == 1
                else:                            # one of the zero bits below it
                    ok = ok and bit == 0
            if not ok:
                continue
            below = sum(1 for pos, _ in inner if pos < k)
            at = any(pos == k for pos, _ in inner)
            shift = k + 1 - below - (1 if at else 0)
            offset = 0 if at else 1 << (k - below)
            emit(shift, offset, [oc[i] for i in keep], coeff)
        if all(bit == 1 for _, bit in inner):     # the all-ones index (the last row repeats)
            coeff = s0
            for c in oc:
                coeff = F.mul(coeff, c)
            emit(0, (1 << m_local) - 1, [], coeff)

    def _reduce(self, c0, c2):
        if self.world == 1:
            return c0, c2
        both = self.b.all_reduce_field(self.dist, np.stack([c0, c2]))
        return both[0], both[1]

    def round(self):
        return self.rep.round() if self.rep is not None else self._reduce(*self.local.round())

    def _after_local_fold(self):
        self.folds += 1
        if self.folds == self.folding:
            self.rep = self.b.sumcheck_gather(self.dist, self.local, self.n_vars_total - self.folding)
            self.local.free()
            self.local = None

    def fold(self, r) -> None:
        if self.rep is not None:
            self.rep.fold(r)
            self.folds += 1
        else:
            self.local.fold(r)
            self._after_local_fold()

    def fold_round(self, r):
        if self.rep is not None:
            self.folds += 1
            return self.rep.fold_round(r)
        if self.folds + 1 == self.folding:
            self.local.fold(r)
            self._after_local_fold()
            return self.rep.round()
        out = self._reduce(*self.local.fold_round(r))
        self.folds += 1
        return out

    def __getattr__(self, name):
        # add_base_eq / read / eval_poly / commit_poly / export_dev: the replicated single-GPU session
        rep = self.__dict__.get("rep")
        if rep is None or name.startswith("_"):
            raise AttributeError(f"{name} is only available after the {self.__dict__.get('folding')} local folds")
        return getattr(rep, name)

    def free(self):
        for s in (self.local, self.rep):
            if s is not None:
                s.free()
        self.local = self.rep = None


# ======================================================================================================
# WhirConfig::commit / prove over the row-sharded initial commitment
# ======================================================================================================
class ShardedTree:
    """The `Tree` surface WhirProver.prove uses (root, open, elem_dim) on a ShardedCommit: openings are served by the rank
    that owns the codeword row and exchanged, so that every rank holds every hint (the transcript is replicated)."""

    elem_dim = 1

    def __init__(self, commit: ShardedCommit):
        self.sc, self.dist = commit, commit.dist
        self.root, self.height, self.log_height = commit.root, commit.geo.h, commit.geo.log_h

    def open(self, indices):
        idx = [int(i) for i in indices]
        own = [(q, i) for q, i in enumerate(idx) if self.sc.geo.owner(i) == self.sc.rank]
        mine = {q: o for (q, _), o in zip(own, _open_local_batch(self.sc, [i for _, i in own]))}
        everyone = [None] * self.sc.world
        self.dist.all_gather_object(everyone, mine)
        merged = {}
        for part in everyone:
            merged.update(part)
        rows = np.stack([merged[q][0] for q in range(len(idx))]) if idx else np.zeros((0, self.sc.full_cols), dtype=np.uint32)
        paths = np.stack([merged[q][1] for q in range(len(idx))]) if idx else np.zeros((0, self.log_height, 8), dtype=np.uint32)
        return rows, paths

    def free(self):
        pass


class ShardedWhirProver:
    """`WhirConfig::commit` / `WhirConfig::prove` (crates/whir/src/commit.rs:64-99, open.rs:37-248) for a stacked polynomial
    that is row-sharded over the ranks.  Round 0 — the commit (ShardedCommit), the out-of-domain evaluations (one all-reduce
    of 5 words each), the first `first_folding` sumcheck rounds (ShardedProductSumcheck) and the STIR openings of the first
    tree (owner-routed) — runs on the shards; from the first fold on the tables are 2^first_folding times smaller and every
    rank continues on the ordinary single-device sessions (the first commit is >= 8x all later ones).  The host
    logic is WhirProver.prove itself; every rank runs it with its own (identical) transcript.  Next-row statements
    (`is_next`, what the stacked-PCS opening emits for table columns) are restricted to the shards term by term
    (ShardedProductSumcheck.add_next)."""

    def __init__(self, backend, dist, cfg):
        from .whir import WhirProver

        self.b, self.dist, self.cfg = backend, dist, cfg
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.g = self.world.bit_length() - 1
        outer = self

        class _Prover(WhirProver):
            def _session(self, witness):
                return ShardedProductSumcheck(outer.b, outer.dist, witness.shard, cfg.num_variables, cfg.first_folding,
                                              witness.live_len)

        self._prover = _Prover(getattr(backend, "ctx", None), cfg)

    def evaluate(self, shard_padded, point_m) -> np.ndarray:
        """value of the whole polynomial at `point_m` (n x 5 Montgomery words) from the shards: the g coordinates that
        select the rank become an eq factor, the local evaluations are added up with one all-reduce"""
        cfg, g = self.cfg, self.g
        k = cfg.first_folding
        pt = np.ascontiguousarray(point_m, dtype=np.uint32).reshape(-1, 5)
        local_pt = np.concatenate([pt[:k], pt[k + g:]])
        scale = prefix_eq([F.from_monty(x) for x in pt[k:k + g]], g, self.rank)
        local = F.mul(F.from_monty(self.b.mle_eval(shard_padded, local_pt)), scale)
        if self.world == 1:
            return F.to_monty(local)
        return self.b.all_reduce_field(self.dist, F.to_monty(local))

    def commit(self, prover_state, shard, live_cols: int | None = None):
        """shard: this rank's part of the polynomial (shard_of layout restricted to the live columns), host numpy"""
        from .whir import Witness, _sample_ood

        cfg = self.cfg
        nv = cfg.num_variables
        cols = (1 << cfg.first_folding) if live_cols is None else live_cols
        shard = np.ascontiguousarray(shard, dtype=np.uint32)
        sc = ShardedCommit(self.b, self.dist, nv, cfg.first_folding, cfg.starting_log_inv_rate, live_cols=cols)
        root = np.asarray(sc.commit(self.b.to_device(shard))).view(np.uint32).reshape(-1)
        prover_state.add_base_scalars(root)
        padded = np.zeros(1 << (nv - self.g), dtype=np.uint32)
        padded[: shard.size] = shard
        pts, answers = _sample_ood(prover_state, cfg.commitment_ood_samples, nv, lambda pt: self.evaluate(padded, pt))
        w = Witness(ShardedTree(sc), pts, answers)
        w.shard, w.live_len = shard, shard.size
        return w

    def prove(self, prover_state, statements, witness):
        return self._prover.prove(prover_state, statements, witness)


class CudaBackend:
    """Compute steps on one GPU through the C ABI; tensors are torch CUDA int32 (device memory + NCCL plumbing)."""

    def __init__(self, ctx):
        import torch

        from ._lib import check, lib

        self.torch, self.ctx, self.lib, self.check = torch, ctx, lib(), check
        # one stream for torch ops, NCCL and the library's kernels
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        ctx.set_stream(self.stream.cuda_stream)
        import os

        self._timing, self._marks = bool(os.environ.get("LM_SHARD_TIMING")), []
        self._scatter = {}   # (rows, cols) -> (own matrix, work buffer, peer tensors kept alive, pointer table)
        self._scatter_ok, self._scatter_error = {}, None
        if os.environ.get("LM_SHARD_EXCHANGE", "p2p") != "p2p":
            self.scatter_dft = None  # fall back to the NCCL all-to-all after the local transform

    def to_device(self, a: np.ndarray):
        return self.torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()

    def to_host(self, t) -> np.ndarray:
        return t.cpu().numpy().view(np.uint32)

    def empty_like(self, t):
        return self.torch.empty_like(t)

    def rows(self, t, start: int, count: int):
        return t[start:start + count]

    def take_rows(self, t, idx) -> np.ndarray:
        """rows `idx` of a device matrix on the host: one gather kernel, one read-back"""
        sel = self.torch.as_tensor(idx, dtype=self.torch.int64, device=t.device)
        return t.index_select(0, sel).cpu().numpy().view(np.uint32)

    def reorder_and_dft(self, shard, n_vars, folding, log_inv_rate, cols):
        h = 1 << (n_vars + log_inv_rate - folding)
        out = self.torch.empty((h, cols), dtype=self.torch.int32, device=shard.device)
        self.check(self.lib.lm_dev_reorder_and_dft(self.ctx.handle, shard.data_ptr(), n_vars, 1, folding, log_inv_rate, cols,
                                                   out.data_ptr()))
        return out

    def _scatter_buffers(self, dist, rows, cols):
        """One matrix per rank (lm_dev_alloc), mapped into every other rank's address space by CUDA IPC (lm_dev_ipc_*: the
        64-byte handles travel through all_gather_object), allocated once per shape: the commit's codeword lives in it."""
        key = (rows, cols)
        if key not in self._scatter:
            import ctypes as C

            torch, lib = self.torch, self.lib
            world, rank = dist.get_world_size(), dist.get_rank()
            # every rank reaches the handle all-gather, whatever failed locally before it (an out-of-memory here must not
            # leave the other ranks blocked in the collective): a failed rank contributes None
            own = work = None
            payload = None
            try:
                own = self.ctx.alloc(rows * cols * 4)
                work = torch.empty((rows, cols), dtype=torch.int32, device="cuda")
                hbuf = C.create_string_buffer(64)
                if lib.lm_dev_ipc_export(self.ctx.handle, own.ptr, hbuf) == 0:
                    payload = bytes(hbuf.raw)
            except Exception:  # noqa: BLE001
                payload = None
            handles = [None] * world
            dist.all_gather_object(handles, payload)
            if any(h is None for h in handles):
                if own is not None:
                    own.free()
                raise RuntimeError("CUDA IPC export (or the exchange buffer allocation) failed on a rank")
            table = np.zeros(world, dtype=np.uint64)
            opened = []
            try:
                for q in range(world):
                    if q == rank:
                        table[q] = own.ptr.value
                    else:
                        p = C.c_void_p()
                        self.check(lib.lm_dev_ipc_open(self.ctx.handle, handles[q], C.byref(p)))
                        table[q] = p.value
                        opened.append(p)
            except Exception:
                for p in opened:  # do not leak the mappings opened so far, nor the own buffer
                    lib.lm_dev_ipc_close(self.ctx.handle, p)
                own.free()
                raise

            class _Raw:  # zero-copy torch view of the library allocation
                __cuda_array_interface__ = {"shape": (rows, cols), "typestr": "<i4", "data": (int(own.ptr.value), False),
                                            "version": 3, "strides": None}

            mat = torch.as_tensor(_Raw(), device="cuda")
            self._scatter[key] = (mat, work, (own, opened), table)
        return self._scatter[key]

    def close(self):
        """unmap the peer matrices and free the own ones (collective: call on every rank before the process group goes)"""
        for mat, work, (own, opened), _ in self._scatter.values():
            for p in opened:
                self.lib.lm_dev_ipc_close(self.ctx.handle, p)
            own.free()
        self._scatter = {}

    def scatter_ready(self, dist, rows, cols) -> bool:
        """Collective: map the peer matrices for this shape; False on EVERY rank when any rank could not (no peer access, IPC
        disabled in the container, ...) — the commit then takes the NCCL all-to-all path instead."""
        key = (rows, cols)
        if key in self._scatter_ok:
            return self._scatter_ok[key]
        ok = 1
        try:
            self._scatter_buffers(dist, rows, cols)
        except Exception as e:  # noqa: BLE001
            ok = 0
            self._scatter_error = e
        flag = self.torch.tensor([ok], dtype=self.torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self._scatter_ok[key] = bool(int(flag.item()))
        return self._scatter_ok[key]

    def scatter_dft(self, dist, shard, n_vars, folding, log_inv_rate, cols):
        import ctypes as C

        rows = 1 << (n_vars + log_inv_rate - folding)
        mat, work, _, table = self._scatter_buffers(dist, rows, cols)
        self.check(self.lib.lm_dev_reorder_and_dft_scatter(self.ctx.handle, shard.data_ptr(), n_vars, folding, log_inv_rate, cols,
                                                           work.data_ptr(), table.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                           dist.get_world_size(), dist.get_rank()))
        # every rank's stores must have landed before anyone reads its matrix: NCCL barrier, ordered on the stream
        dist.all_reduce(self._flag())
        return mat

    def commit_host_pipelined(self, sc, host_shard):
        """-> (codeword matrix, forest) or None when the shape is not eligible.  host_shard: pinned torch int32 tensor, the
        shard polynomial (live columns x positions)."""
        import ctypes as C

        torch, lib, geo, dist = self.torch, self.lib, sc.geo, sc.dist
        w, full = sc.cols, sc.full_cols
        n_chunks = w // 8
        if self.scatter_dft is None or w % 8 or n_chunks < 2 or (full - w) // 8 < 2 or not torch.is_tensor(host_shard):
            return None
        if not self.scatter_ready(dist, geo.block, w):
            return None
        mat, work, _, table = self._scatter_buffers(dist, geo.block, w)
        sub = host_shard.numel() // w                       # positions per column in the shard
        d_evals = torch.empty(host_shard.numel(), dtype=torch.int32, device="cuda")
        forest = torch.empty((2 * geo.block - 1, 8), dtype=torch.int32, device="cuda")
        priv = self.empty_like(mat)                         # the witness's codeword (the exchange buffer is reused by the next commit)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
        self._copy_stream.wait_stream(self.stream)
        n_vars, world, rank = geo.n_vars - geo.g, dist.get_world_size(), dist.get_rank()
        tptr = table.ctypes.data_as(C.POINTER(C.c_uint64))
        # up to 8 equal column groups of whole rate chunks, right to left (as lm_commit: with the tensor-core sponge a group's
        # transform + hash takes about as long as its copy, so equal groups keep PCIe and the SMs busy and leave one group's
        # compute after the last byte; LM_COMMIT_DOUBLING_GROUPS=1: the round-1 schedule 1, 1, 2, 4, ..)
        doubling = bool(os.environ.get("LM_COMMIT_DOUBLING_GROUPS"))
        # a group is at least as wide as the tile of the scattering pass (32 columns at 4+ ranks, 16 at 2: 128- / 64-byte row
        # pieces over NVLink instead of 32-byte ones, csrc/ntt.cu ntt_pass_wide_kernel), which also bounds the number of
        # exchange barriers per commit
        wide_chunks = (4 if world >= 4 else 2) if n_chunks % (4 if world >= 4 else 2) == 0 else 1
        chunk_end, take, first = n_chunks, (1 if doubling else max((n_chunks + 7) // 8, wide_chunks)), True
        while chunk_end > 0:
            take = min(take, chunk_end)
            col_begin, count = (chunk_end - take) * 8, take * 8
            with torch.cuda.stream(self._copy_stream):
                d_evals[col_begin * sub:(col_begin + count) * sub].copy_(host_shard[col_begin * sub:(col_begin + count) * sub],
                                                                         non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self.stream.wait_event(ev)
            self.check(lib.lm_dev_reorder_and_dft_scatter_cols(self.ctx.handle, d_evals.data_ptr(), n_vars, geo.folding,
                                                               geo.log_inv_rate, w, work.data_ptr(), tptr, world, rank, col_begin, count))
            dist.all_reduce(self._flag())                   # every rank's stores of this group have landed
            self.check(lib.lm_dev_dft_layers_mapped_out(self.ctx.handle, mat.data_ptr(), priv.data_ptr(), w, geo.log_h, geo.log_h - geo.g,
                                                        world, geo.run, geo.block, rank * geo.run, col_begin, count))
            self.check(lib.lm_dev_merkle_absorb(self.ctx.handle, priv.data_ptr(), geo.block, w, full, w, chunk_end - 1, take,
                                                forest.data_ptr()))
            chunk_end -= take
            if doubling and not first:
                take *= 2
            first = False
        self.check(lib.lm_dev_merkle_levels(self.ctx.handle, forest.data_ptr(), geo.block))
        d_evals.record_stream(self._copy_stream)
        return priv, forest

    def _flag(self):
        if not hasattr(self, "_flag_t"):
            self._flag_t = self.torch.zeros(1, dtype=self.torch.int32, device="cuda")
        return self._flag_t

    def all_to_all(self, dist, out, inp):
        # torch's current stream IS the library's stream: NCCL orders itself after the transform and the next kernel after
        # the exchange through stream events, no host synchronisation
        dist.all_to_all_single(out, inp)

    def mark(self, name):
        """per-phase device timestamps when LM_SHARD_TIMING is set (read with phase_times())"""
        if not self._timing:
            return
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.stream)
        if name == "start":
            self._marks = []
        self._marks.append((name, ev))

    def phase_times(self) -> dict:
        self.torch.cuda.synchronize()
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self._marks, self._marks[1:])}

    def dft_layers_mapped(self, mat, w, log_h, l_first, n_blocks, run, block, offset, out=None):
        if out is None:
            self.check(self.lib.lm_dev_dft_layers_mapped(self.ctx.handle, mat.data_ptr(), w, log_h, l_first, n_blocks, run, block, offset))
        else:
            self.check(self.lib.lm_dev_dft_layers_mapped_out(self.ctx.handle, mat.data_ptr(), out.data_ptr(), w, log_h, l_first, n_blocks,
                                                             run, block, offset, 0, 0))

    def merkle_tree(self, rows, full_cols, eff_cols):
        h, w = rows.shape
        layers = self.torch.empty((2 * h - 1, 8), dtype=self.torch.int32, device=rows.device)
        self.check(self.lib.lm_dev_merkle_tree(self.ctx.handle, rows.data_ptr(), h, w, full_cols, eff_cols, layers.data_ptr()))
        return layers

    def all_gather_roots(self, dist, my_roots):
        world = dist.get_world_size()
        out = self.torch.empty((world,) + tuple(my_roots.shape), dtype=my_roots.dtype, device=my_roots.device)
        if world > 1:
            dist.all_gather_into_tensor(out, my_roots.contiguous())
        else:
            out[0] = my_roots
        return out

    def permute_roots(self, all_roots):
        # all_roots[rank][m] -> layer index m * world + rank
        return all_roots.permute(1, 0, 2).contiguous().reshape(-1, 8)

    def merkle_levels(self, layer0):
        n = layer0.shape[0]
        layers = self.torch.empty((2 * n - 1, 8), dtype=self.torch.int32, device=layer0.device)
        layers[:n] = layer0
        if n > 1:
            self.check(self.lib.lm_dev_merkle_levels(self.ctx.handle, layers.data_ptr(), n))
        return layers

    # ---- AIR sumcheck -----------------------------------------------------------------------------------------
    def air_session(self, table_id, columns, eq_factor, alpha_powers, la, beta, **kw):
        from .air import AirSumcheckSession

        zero = np.zeros(5, dtype=np.uint32)  # the running sum lives in the sharded session, not in the per-rank ones
        return AirSumcheckSession(self.ctx, table_id, columns, eq_factor, zero, alpha_powers, la, beta, **kw)

    def all_reduce_field(self, dist, words: np.ndarray) -> np.ndarray:
        """sum mod p over the ranks of an array of Montgomery residues (one NCCL all-reduce on widened words)"""
        t = self.torch.from_numpy(np.ascontiguousarray(words).astype(np.int64)).cuda()
        dist.all_reduce(t)
        return (t.cpu().numpy() % P).astype(np.uint32).reshape(words.shape)

    def all_gather_words(self, dist, words: np.ndarray) -> np.ndarray:
        world = dist.get_world_size()
        t = self.torch.from_numpy(np.ascontiguousarray(words).view(np.int32)).cuda()
        out = self.torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().view(np.uint32)

    # ---- quotient GKR ---------------------------------------------------------------------------------------------
    def gkr_session(self, nums, dens, n_vars, top_vars):
        from .logup import GkrShardSession

        return GkrShardSession(self.ctx, nums, dens, n_vars, top_vars)

    # ---- WHIR product sumcheck ------------------------------------------------------------------------------------
    def sumcheck(self, evals, n_vars, live_len=None):
        return self.ctx.sumcheck(evals, n_vars, live_len)

    def sumcheck_gather(self, dist, local, n_vars_total):
        """all-gather (device to device, rank order = the sharding bits) of the folded tables of `local` and a session on
        the full tables"""
        world = dist.get_world_size()
        n_local = 5 << local.n_vars
        mine = self.torch.empty((2, n_local), dtype=self.torch.int32, device="cuda")
        local.export_dev(mine[0].data_ptr(), mine[1].data_ptr())
        everyone = self.torch.empty((world, 2, n_local), dtype=self.torch.int32, device="cuda")
        if world > 1:
            dist.all_gather_into_tensor(everyone, mine)
        else:
            everyone[0] = mine
        tables = everyone.permute(1, 0, 2).contiguous()          # [poly | weights][rank][local index]
        self.torch.cuda.synchronize()
        return self.ctx.sumcheck_from_dev(tables[0].data_ptr(), tables[1].data_ptr(), n_vars_total)

    def mle_eval(self, evals, point_m) -> np.ndarray:
        return self.ctx.mle_eval(evals, point_m)

"""Merkle-path pruning of the proof's opening hints, prover side.

Reference: MerklePaths::prune (crates/backend/fiat-shamir/src/merkle_pruning.rs:18-84).  The openings of one query batch
(what `Tree.open` / `lm_open` returns: zero-extended rows + sibling paths) are sorted by leaf index and de-duplicated; a
path only keeps the siblings below its lowest common ancestor with the previous leaf, minus the one sibling that is the
subtree root of the next leaf's side (the verifier recomputes it); the all-zero tail shared by every leaf is dropped.
Host logic on integers and arrays — no field arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def lca_level(a: int, b: int) -> int:
    """merkle_pruning.rs:14-16: number of levels below the lowest common ancestor of leaves a != b"""
    return (a ^ b).bit_length()


@dataclass
class PrunedMerklePaths:
    merkle_height: int
    original_order: list          # position of every original query in the de-duplicated, sorted list
    leaf_data: list               # per kept leaf: the row without the common zero tail
    paths: list                   # per kept leaf: (leaf_index, [sibling digests kept])
    n_trailing_zeros: int

    def n_digests(self) -> int:
        return sum(len(s) for _, s in self.paths)


def prune(leaf_indices, rows, sibling_paths) -> PrunedMerklePaths:
    """leaf_indices: n ints; rows: n x width (zero-extended leaves); sibling_paths: n x height x 8 (leaf level first)"""
    idx = [int(i) for i in leaf_indices]
    rows = np.asarray(rows)
    paths = np.asarray(sibling_paths)
    assert len(idx) > 0 and rows.shape[0] == len(idx) == paths.shape[0]
    height = paths.shape[1]
    order = sorted(range(len(idx)), key=lambda q: idx[q])  # stable, like sort_by_key
    original_order = [0] * len(idx)
    kept = []                                              # positions (into the originals) of the de-duplicated leaves
    for q in order:
        if kept and idx[kept[-1]] == idx[q]:
            original_order[q] = len(kept) - 1
        else:
            original_order[q] = len(kept)
            kept.append(q)
    leaf_len = rows.shape[1]
    n_trailing_zeros = 0
    for off in range(leaf_len - 1, -1, -1):
        if any(rows[q, off] != 0 for q in kept):
            break
        n_trailing_zeros += 1
    out_paths = []
    for i, q in enumerate(kept):
        levels = height if i == 0 else lca_level(idx[kept[i - 1]], idx[q])
        skip = lca_level(idx[q], idx[kept[i + 1]]) - 1 if i + 1 < len(kept) else None
        out_paths.append((idx[q], [paths[q, lvl].copy() for lvl in range(levels) if lvl != skip]))
    return PrunedMerklePaths(height, original_order, [rows[q, : leaf_len - n_trailing_zeros].copy() for q in kept], out_paths,
                             n_trailing_zeros)

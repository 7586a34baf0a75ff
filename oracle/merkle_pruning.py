"""ORACLE — TEST INFRASTRUCTURE ONLY.  Verifier side of the Merkle-path pruning: PrunedMerklePaths::restore
(crates/backend/fiat-shamir/src/merkle_pruning.rs:86-176) with the tree's own hash functions (hash_slice for leaves,
compress for nodes) — rebuilds every full opening from the pruned hint, or returns None when the hint is malformed."""
from __future__ import annotations

import numpy as np

import oracle as O


def lca_level(a: int, b: int) -> int:
    return (a ^ b).bit_length()


def _combine(left, right):
    st = np.concatenate([left, right]).astype(np.uint32)[None, :]
    return O.poseidon1_compress(st)[0, :8]


def restore(pruned, hash_leaf=None, hash_combine=None):
    """-> list of (leaf_index, row, siblings[height x 8]) in the ORIGINAL query order, or None"""
    hash_leaf = hash_leaf or (lambda row: O.hash_slice(np.asarray(row, dtype=np.uint32)))
    hash_combine = hash_combine or _combine
    n, h = len(pruned.paths), pruned.merkle_height
    if h >= 32 or pruned.n_trailing_zeros > 1024:
        return None
    leaf_data = [np.concatenate([np.asarray(d), np.zeros(pruned.n_trailing_zeros, dtype=np.asarray(d).dtype)]) for d in pruned.leaf_data]
    if len(leaf_data) != n:
        return None

    def levels(i):
        return h if i == 0 else lca_level(pruned.paths[i - 1][0], pruned.paths[i][0])

    def skip(i):
        return lca_level(pruned.paths[i][0], pruned.paths[i + 1][0]) - 1 if i + 1 < n else None

    subtree = [[] for _ in range(n)]
    for i in range(n - 1, -1, -1):               # backward pass: subtree hashes that restore the skipped siblings
        leaf_idx, stored = pruned.paths[i]
        if leaf_idx >= 1 << h:
            return None
        it = iter(stored)
        cur = hash_leaf(leaf_data[i])
        subtree[i].append(cur)
        for lvl in range(levels(i)):
            if skip(i) == lvl:
                if lvl >= len(subtree[i + 1]):
                    return None
                sib = subtree[i + 1][lvl]
            else:
                sib = next(it, None)
                if sib is None:
                    return None
            cur = hash_combine(cur, sib) if ((leaf_idx >> lvl) & 1) == 0 else hash_combine(sib, cur)
            subtree[i].append(cur)
    restored = []
    for i in range(n):                            # forward pass: full sibling arrays
        leaf_idx, stored = pruned.paths[i]
        it = iter(stored)
        sibs = []
        for lvl in range(levels(i)):
            sibs.append(subtree[i + 1][lvl] if skip(i) == lvl else next(it))
        if restored:
            sibs.extend(restored[-1][2][levels(i):])
        if len(sibs) != h:
            return None
        restored.append((leaf_idx, leaf_data[i], sibs))
    out = []
    for pos in pruned.original_order:
        if pos >= len(restored):
            return None
        li, row, sibs = restored[pos]
        out.append((li, row, np.stack(sibs) if sibs else np.zeros((0, 8), dtype=np.uint32)))
    return out

"""Proof wire format: postcard + lz4, so that a proof produced on the GPU path is the byte
string the unchanged reference verifier parses.

Reference:
  Proof<F> { transcript: Vec<F>, merkle_paths: Vec<PrunedMerklePaths<F, F>> }   crates/backend/fiat-shamir/src/transcript.rs:31-35
  PrunedMerklePaths { merkle_height: usize, original_order: Vec<usize>, leaf_data: Vec<Vec<F>>,
                      paths: Vec<(usize, Vec<[F; 8]>)>, n_trailing_zeros: usize }  fiat-shamir/src/merkle_pruning.rs:5-12
  F serialises as serialize_u32(Montgomery value)                               koala-bear/src/monty_31/monty_31.rs:152-157
  compress = lz4_flex::compress_prepend_size(postcard::to_allocvec(..))         rec_aggregation/src/type_1_aggregation.rs:81-89

postcard 1.1.3 and lz4_flex 0.13.0 are crates.io dependencies (Cargo.lock), absent from the reference tree; their published
formats are restated here: postcard writes u32 / u64 / usize as LEB128 varints (7 bits per byte, low bits first), a
sequence as varint(len) + its elements, tuples / structs / fixed-size arrays as their elements with no prefix.  The LZ4
block codec is csrc/wire.cpp.  Host logic only; the oracle verifier (tests) consumes `Proof.restore()`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._lib import check, lib
from .merkle_pruning import PrunedMerklePaths, prune

P = 0x7F000001


# ---------------------------------------------------------------------------------------------------- postcard
def varint(x: int) -> bytes:
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def varints_u32(a) -> bytes:
    """postcard encoding of a u32 array, element after element (vectorised: a proof holds ~10^5 field elements)"""
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1).astype(np.uint64)
    if a.size == 0:
        return b""
    n_bytes = np.ones(a.size, dtype=np.int64)
    for k in range(1, 5):
        n_bytes += (a >= (1 << (7 * k))).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(n_bytes)])
    out = np.zeros(int(offs[-1]), dtype=np.uint8)
    for k in range(5):
        sel = n_bytes > k
        byte = ((a[sel] >> np.uint64(7 * k)) & np.uint64(0x7F)).astype(np.uint8)
        cont = (n_bytes[sel] > k + 1).astype(np.uint8) << 7
        out[offs[:-1][sel] + k] = byte | cont
    return out.tobytes()


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.i = data, 0

    def varint(self, max_bits: int = 64) -> int:
        x, shift = 0, 0
        while True:
            if self.i >= len(self.d):
                raise ValueError("postcard: unexpected end of input")
            b = self.d[self.i]
            self.i += 1
            x |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
            if shift > 70:
                raise ValueError("postcard: varint too long")
        if x >> max_bits:
            raise ValueError("postcard: varint overflows its type")
        return x

    def field(self) -> int:
        v = self.varint(32)
        if v >= P:
            raise ValueError("non-canonical MontyField31 value")  # monty_31.rs:159-167
        return v

    def fields(self, n: int) -> np.ndarray:
        return np.array([self.field() for _ in range(n)], dtype=np.uint32)


def encode_pruned(p: PrunedMerklePaths) -> bytes:
    out = [varint(p.merkle_height), varint(len(p.original_order))]
    out += [varint(int(o)) for o in p.original_order]
    out.append(varint(len(p.leaf_data)))
    for d in p.leaf_data:
        d = np.asarray(d, dtype=np.uint32).reshape(-1)
        out += [varint(d.size), varints_u32(d)]
    out.append(varint(len(p.paths)))
    for leaf_index, sibs in p.paths:
        out += [varint(int(leaf_index)), varint(len(sibs))]
        if len(sibs):
            out.append(varints_u32(np.stack([np.asarray(s, dtype=np.uint32).reshape(8) for s in sibs])))
    out.append(varint(p.n_trailing_zeros))
    return b"".join(out)


def _decode_pruned(r: _Reader) -> PrunedMerklePaths:
    height = r.varint()
    order = [r.varint() for _ in range(r.varint())]
    leaf_data = [r.fields(r.varint()) for _ in range(r.varint())]
    paths = []
    for _ in range(r.varint()):
        leaf_index = r.varint()
        paths.append((leaf_index, [r.fields(8) for _ in range(r.varint())]))
    return PrunedMerklePaths(height, order, leaf_data, paths, r.varint())


# ---------------------------------------------------------------------------------------------------- lz4 (csrc/wire.cpp)
def lz4_compress_prepend_size(data: bytes) -> bytes:
    L = lib()
    cap = int(L.lm_lz4_compress_bound(len(data)))
    src = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
    dst = (C.c_uint8 * cap)()
    n = C.c_uint64()
    check(L.lm_lz4_compress_prepend_size(src, len(data), dst, cap, C.byref(n)))
    return bytes(dst[: n.value])


def lz4_decompress_size_prepended(data: bytes) -> bytes:
    L = lib()
    src = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
    n = C.c_uint64()
    check(L.lm_lz4_decompress_size_prepended(src, len(data), None, 0, C.byref(n)))
    if n.value > (1 << 30):
        raise ValueError("lz4: declared size too large")
    dst = (C.c_uint8 * max(n.value, 1))()
    check(L.lm_lz4_decompress_size_prepended(src, len(data), dst, n.value, C.byref(n)))
    return bytes(dst[: n.value])


# ---------------------------------------------------------------------------------------------------- Proof
@dataclass
class Proof:
    """fiat-shamir/src/transcript.rs:31-35; field elements are Montgomery-form u32 as everywhere in this package"""

    transcript: np.ndarray
    merkle_paths: list = field(default_factory=list)  # PrunedMerklePaths, one per hint_merkle_paths call

    @staticmethod
    def from_prover_state(prover_state) -> "Proof":
        """ProverState::into_proof (fiat-shamir/src/prover.rs:48-53); every query batch is pruned as hint_merkle_paths_base
        does (prover.rs:116-118).  prover_state.merkle_paths: lists of (row, sibling path, leaf index)."""
        pruned = []
        for batch in prover_state.merkle_paths:
            if isinstance(batch, PrunedMerklePaths):
                pruned.append(batch)
                continue
            rows = np.stack([np.asarray(r, dtype=np.uint32) for r, _, _ in batch])
            paths = np.stack([np.asarray(p, dtype=np.uint32) for _, p, _ in batch])
            pruned.append(prune([int(i) for _, _, i in batch], rows, paths))
        return Proof(np.array(prover_state.transcript, dtype=np.uint32), pruned)

    def proof_size_fe(self) -> int:
        """transcript.rs:37-53"""
        return int(self.transcript.size) + sum(sum(len(d) for d in p.leaf_data) + 8 * p.n_digests() for p in self.merkle_paths)

    def to_postcard(self) -> bytes:
        out = [varint(int(self.transcript.size)), varints_u32(self.transcript), varint(len(self.merkle_paths))]
        out += [encode_pruned(p) for p in self.merkle_paths]
        return b"".join(out)

    @staticmethod
    def from_postcard(data: bytes) -> "Proof":
        r = _Reader(data)
        transcript = r.fields(r.varint())
        paths = [_decode_pruned(r) for _ in range(r.varint())]
        if r.i != len(data):
            raise ValueError("postcard: trailing bytes")
        return Proof(transcript, paths)

    def compress(self) -> bytes:
        return lz4_compress_prepend_size(self.to_postcard())

    @staticmethod
    def decompress(data: bytes) -> "Proof":
        return Proof.from_postcard(lz4_decompress_size_prepended(data))

"""Proof wire format: postcard + lz4 (reference: fiat-shamir/src/transcript.rs:31-35,
rec_aggregation/src/type_1_aggregation.rs:81-89).  Byte-level checks against encodings written out by hand from the
published postcard / LZ4-block formats, round trips, rejection of malformed input, and an end-to-end pass: a proof goes
prover state -> pruned paths -> postcard -> lz4 -> bytes -> back -> restored openings -> verifier accepts."""
import numpy as np
import pytest

import oracle as O
from leanmultisig_b200 import wire
from leanmultisig_b200.merkle_pruning import PrunedMerklePaths


def test_varint_known_values():
    assert wire.varint(0) == b"\x00" and wire.varint(127) == b"\x7f" and wire.varint(128) == b"\x80\x01"
    assert wire.varint(300) == b"\xac\x02" and wire.varint(2**32 - 1) == b"\xff\xff\xff\xff\x0f"
    vals = [0, 1, 127, 128, 16383, 16384, 2**21 - 1, 2**21, 2**28 - 1, 2**28, O.P - 1, 2**32 - 1]
    assert wire.varints_u32(np.array(vals, dtype=np.uint32)) == b"".join(wire.varint(v) for v in vals)
    rng = np.random.default_rng(1)
    a = rng.integers(0, O.P, size=5000, dtype=np.uint32)
    assert wire.varints_u32(a) == b"".join(wire.varint(int(v)) for v in a)


def test_postcard_of_a_hand_encoded_proof():
    d0 = np.arange(8, dtype=np.uint32) + 1000
    proof = wire.Proof(np.array([1, 200, 0x7F000000], dtype=np.uint32),
                       [PrunedMerklePaths(2, [0, 1], [np.array([5, 6], dtype=np.uint32), np.array([7], dtype=np.uint32)],
                                          [(1, [d0]), (3, [])], 3)])
    want = bytes([3, 1]) + b"\xc8\x01" + b"\x80\x80\x80\xf8\x07"  # transcript: len 3; 1; 200; 0x7f000000
    want += bytes([1])                                             # one PrunedMerklePaths
    want += bytes([2])                                             # merkle_height
    want += bytes([2, 0, 1])                                       # original_order: len 2; 0; 1
    want += bytes([2, 2, 5, 6, 1, 7])                              # leaf_data: 2 leaves: [5, 6], [7]
    want += bytes([2, 1, 1]) + b"".join(wire.varint(1000 + k) for k in range(8))  # paths[0] = (1, [d0]): index 1, 1 sibling of 8 F
    want += bytes([3, 0])                                          # paths[1] = (3, [])
    want += bytes([3])                                             # n_trailing_zeros
    assert proof.to_postcard() == want
    back = wire.Proof.from_postcard(want)
    assert np.array_equal(back.transcript, proof.transcript) and back.merkle_paths[0].merkle_height == 2
    assert back.merkle_paths[0].original_order == [0, 1] and back.merkle_paths[0].n_trailing_zeros == 3
    assert [list(d) for d in back.merkle_paths[0].leaf_data] == [[5, 6], [7]]
    assert back.merkle_paths[0].paths[0][0] == 1 and np.array_equal(back.merkle_paths[0].paths[0][1][0], d0)
    assert back.merkle_paths[0].paths[1] == (3, [])
    assert proof.proof_size_fe() == 3 + 3 + 8
    with pytest.raises(ValueError):
        wire.Proof.from_postcard(want + b"\x00")                   # trailing bytes
    with pytest.raises(ValueError):
        wire.Proof.from_postcard(bytes([1]) + b"\x81\x80\x80\xf8\x07" + bytes([0]))  # 0x7f000001 = p: non-canonical


def test_lz4_hand_encoded_blocks():
    # 20 x 'a': one literal, a match of 14 at offset 1 (token 0x1A), then the mandatory 5 trailing literals (token 0x50)
    block = bytes([20, 0, 0, 0, 0x1A]) + b"a" + bytes([1, 0, 0x50]) + b"aaaaa"
    assert wire.lz4_decompress_size_prepended(block) == b"a" * 20
    assert wire.lz4_compress_prepend_size(b"a" * 20) == block
    # literals only (incompressible / short input): token 0xB0 + 11 bytes
    assert wire.lz4_compress_prepend_size(b"hello world") == bytes([11, 0, 0, 0, 0xB0]) + b"hello world"
    # a literal run of 15 + 3 uses one length-extension byte; long matches use 255-runs
    data = bytes(range(18))
    assert wire.lz4_compress_prepend_size(data) == bytes([18, 0, 0, 0, 0xF0, 3]) + data
    long_run = b"xy" * 400
    comp = wire.lz4_compress_prepend_size(long_run)
    assert len(comp) < 40 and wire.lz4_decompress_size_prepended(comp) == long_run
    assert wire.lz4_compress_prepend_size(b"") == bytes([0, 0, 0, 0, 0]) and wire.lz4_decompress_size_prepended(bytes([0, 0, 0, 0, 0])) == b""


def test_lz4_round_trips_and_rejects_malformed_blocks():
    rng = np.random.default_rng(2)
    for n in list(range(0, 40)) + [255, 256, 270, 4096, 65535, 65536, 70001, 300000]:
        raw = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        structured = (rng.integers(0, 4, size=n, dtype=np.uint8) * 17).tobytes()
        for data in (raw, structured):
            comp = wire.lz4_compress_prepend_size(data)
            assert wire.lz4_decompress_size_prepended(comp) == data
    good = wire.lz4_compress_prepend_size(b"abcd" * 100)
    from leanmultisig_b200 import LmError

    for bad in (good[:-1], good[:4], bytes([9, 0, 0, 0]) + good[4:], bytes([4, 0, 0, 0, 0x10, 65, 5, 0])):
        with pytest.raises((LmError, ValueError)):
            wire.lz4_decompress_size_prepended(bad)


@pytest.mark.parametrize("nv", [10, 13])
def test_proof_bytes_round_trip_through_the_verifier(nv):
    """oracle CPU prover -> Proof (pruned) -> compress -> decompress -> restore -> oracle verifier accepts, and the restored
    openings are exactly the prover's hints"""
    from oracle import merkle_pruning as OMP
    from oracle import whir as W
    from test_whir_protocol import SMALL, make_statements, oracle_prove, oracle_verify

    rng = np.random.default_rng(40 + nv)
    cfg = W.WhirConfig(nv, **SMALL)
    poly = O.random_field(rng, 1 << nv)
    stm = make_statements(rng, poly, nv)
    ps, point = oracle_prove(cfg, poly, stm, poly.size)
    proof = wire.Proof.from_prover_state(ps)
    blob = proof.compress()
    back = wire.Proof.decompress(blob)
    assert back.to_postcard() == proof.to_postcard() and len(blob) <= len(proof.to_postcard()) * 256 // 255 + 32
    assert np.array_equal(back.transcript, np.array(ps.transcript, dtype=np.uint32))
    restored = []
    for pruned, original in zip(back.merkle_paths, ps.merkle_paths):
        batch = OMP.restore(pruned)
        assert batch is not None and len(batch) == len(original)
        for (i, row, sibs), (orow, opath, oi) in zip(batch, original):
            assert i == oi and np.array_equal(row, orow) and np.array_equal(sibs, opath)
        restored.append([(row, sibs, i) for i, row, sibs in batch])
    assert oracle_verify(cfg, [int(x) for x in back.transcript], restored, stm) == point
    assert proof.proof_size_fe() < sum(len(r) + 8 * len(p) for b in ps.merkle_paths for r, p, _ in b) + len(ps.transcript)

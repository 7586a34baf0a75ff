"""Oracle pinned against the reference's only stored vector (Poseidon1 KAT) and cross-formulation checks."""
import numpy as np
import pytest

import oracle as O

# reference: crates/backend/koala-bear/src/poseidon1_koalabear_16.rs:1082-1092
KAT_OUT = [610090613, 935319874, 1893335292, 796792199, 356405232, 552237741, 55134556, 1215104204,
           1823723405, 1133298033, 1780633798, 1453946561, 710069176, 1128629550, 1917333254, 1175481618]


def test_kat_dense_and_sparse():
    x = O.to_monty(np.arange(16))
    for dense in (True, False):
        y = O.poseidon1_permute(x, dense=dense)
        assert O.from_monty(y).tolist() == KAT_OUT


def test_dense_equals_sparse_random(rng):
    x = O.random_field(rng, (257, 16))
    assert np.array_equal(O.poseidon1_permute(x, dense=True), O.poseidon1_permute(x, dense=False))


def test_compress_is_permute_plus_input(rng):
    x = O.random_field(rng, (33, 16))
    y = O.poseidon1_permute(x)
    z = ((y.astype(np.uint64) + x) % O.P).astype(np.uint32)
    assert np.array_equal(O.poseidon1_compress(x), z)


def test_field_constants():
    # two-adic generators: g_k^2 = g_{k-1}, g_1 = -1   (koala_bear.rs:46-54)
    for k in range(1, 25):
        g = O.two_adic_generator(k)
        assert O.kb_mul(g, g) == O.two_adic_generator(k - 1)
    assert int(O.from_monty(O.two_adic_generator(1))) == O.P - 1
    # 16^-1 mod p as quoted at poseidon1_koalabear_16.rs:603
    assert int(O.from_monty(O.kb_inv(int(O.to_monty(16))))) == 1997537281


def test_monty_roundtrip(rng):
    x = rng.integers(0, O.P, size=1000, dtype=np.uint32)
    assert np.array_equal(O.from_monty(O.to_monty(x)), x)
    a, b = int(x[0]), int(x[1])
    assert int(O.from_monty(O.kb_mul(int(O.to_monty(a)), int(O.to_monty(b))))) == a * b % O.P


def test_quintic_extension(rng):
    # X^5 = 1 - X^2  (extension.rs:26): X * X^4 = -X^2 + 1
    one = int(O.to_monty(1))
    X = np.array([0, one, 0, 0, 0], dtype=np.uint32)
    X4 = np.array([0, 0, 0, 0, one], dtype=np.uint32)
    exp = np.array([one, 0, O.P - one, 0, 0], dtype=np.uint32)
    assert np.array_equal(O.ef_mul(X, X4), exp)
    a = O.random_field(rng, 5)
    b = O.random_field(rng, 5)
    c = O.random_field(rng, 5)
    assert np.array_equal(O.ef_mul(a, b), O.ef_mul(b, a))
    assert np.array_equal(O.ef_mul(O.ef_mul(a, b), c), O.ef_mul(a, O.ef_mul(b, c)))
    assert np.array_equal(O.ef_mul(a, O.ef_inv(a)), np.array([one, 0, 0, 0, 0], dtype=np.uint32))
    # Frobenius matrix row 0 = X^p (quintic_extension/mod.rs:19-27)
    frob1 = [1576402667, 1173144480, 1567662457, 1206866823, 2428146]
    acc = np.array([one, 0, 0, 0, 0], dtype=np.uint32)
    base = X.copy()
    e = O.P
    while e:
        if e & 1:
            acc = O.ef_mul(acc, base)
        base = O.ef_mul(base, base)
        e >>= 1
    assert O.from_monty(acc).tolist() == frob1


# ---- AVX-512 batch path (the CPU baseline of bench.py) against the scalar oracle -------------------------------
def _need_avx512():
    if not O.lib().lm_or_have_avx512():
        pytest.skip("host has no AVX-512 (the oracle then uses its scalar path everywhere)")


def test_avx512_compress_matches_scalar(rng):
    import ctypes as C

    _need_avx512()
    states = O.random_field(rng, (16, 16))
    states[3] = 0
    states[5] = 0x7F000000  # p - 1 in every lane of one state
    exp = O.poseidon1_compress(states.copy())
    got = np.ascontiguousarray(states.copy())
    O.lib().lm_or_poseidon1_compress_x16(got.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("stored,full,eff", [(64, 128, 64), (128, 128, 128), (24, 32, 24), (16, 160, 16), (40, 48, 40)])
def test_avx512_leaf_digests_match_hash_slice(rng, stored, full, eff):
    _need_avx512()
    h = 48 + 5  # three 16-row blocks through the vector path and a scalar tail
    mat = O.random_field(rng, (h, stored))
    mat[:, eff:] = 0
    got = O.first_digest_layer(mat, full, eff)
    for r in (0, 7, 15, 16, 47, 48, 52):
        row = np.zeros(full, dtype=np.uint32)
        row[:stored] = mat[r]
        assert np.array_equal(got[r], O.hash_slice(row)), f"row {r}"


def test_avx512_tree_layers_match_scalar_compress(rng):
    _need_avx512()
    h = 64
    mat = O.random_field(rng, (h, 16))
    layers = O.merkle_tree(mat, 16, 16)
    off, n = 0, h
    while n > 1:
        prev = layers[off:off + n]
        st = np.concatenate([prev[0::2], prev[1::2]], axis=1)
        assert np.array_equal(layers[off + n:off + n + n // 2], O.poseidon1_compress(st)[:, :8])
        off += n
        n //= 2

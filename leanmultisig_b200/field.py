"""Host-side KoalaBear / quintic-extension arithmetic on Python ints.

Only for the handful of per-round values the reference also computes on the host (p(1), Lagrange
interpolation, eq factors of a challenge): crates/backend/poly/src/dense_poly.rs:33,
crates/sub_protocols/src/air_sumcheck.rs:250-275.  Boundary values are Montgomery-form u32 exactly like the
reference's memory; internally canonical residues are used.
"""
from __future__ import annotations

import numpy as np

P = 0x7F000001
_R = (1 << 32) % P
_RINV = pow(_R, -1, P)


def from_monty(v) -> tuple:
    return tuple(int(x) * _RINV % P for x in np.asarray(v, dtype=np.uint64).reshape(-1))


def to_monty(e) -> np.ndarray:
    return np.array([x % P * _R % P for x in e], dtype=np.uint32)


ZERO = (0, 0, 0, 0, 0)
ONE = (1, 0, 0, 0, 0)


def add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


def sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


def neg(a):
    return tuple((-x) % P for x in a)


def scal(a, k: int):
    return tuple(x * k % P for x in a)


def mul(a, b):
    """product in F[X]/(X^5 + X^2 - 1)  (quintic_extension/extension.rs:26)"""
    d = [0] * 9
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                d[i + j] += x * y
    return ((d[0] + d[5] - d[8]) % P, (d[1] + d[6]) % P, (d[2] - d[5] + d[7] + d[8]) % P, (d[3] - d[6] + d[8]) % P,
            (d[4] - d[7]) % P)


def power(a, e: int):
    r = ONE
    while e:
        if e & 1:
            r = mul(r, a)
        a = mul(a, a)
        e >>= 1
    return r


# X^(p i) for i = 1..4 (crates/backend/koala-bear/src/quintic_extension/mod.rs:19-48), canonical residues;
# checked against exponentiation in tests/test_air_sumcheck.py::test_host_field_matches_oracle
_FROBENIUS = (
    (1576402667, 1173144480, 1567662457, 1206866823, 2428146),
    (1680345488, 1381986, 615237464, 1380104858, 295431824),
    (441230756, 323126830, 704986542, 1445620072, 503505220),
    (1364444097, 1144738982, 2008416047, 143367062, 1027410849),
)


def frobenius(a):
    """a -> a^p: sum_i a_i (X^p)^i, a linear map with the matrix above"""
    out = [a[0], 0, 0, 0, 0]
    for i in range(1, 5):
        if a[i]:
            for k in range(5):
                out[k] += a[i] * _FROBENIUS[i - 1][k]
    return tuple(x % P for x in out)


def inv(a):
    """a^-1 = (a^p a^(p^2) a^(p^3) a^(p^4)) / Norm(a)   (quintic_extension/extension.rs:585-613)"""
    f1 = frobenius(a)
    f12 = frobenius(mul(a, f1))          # a^(p + p^2)
    conj = mul(f12, frobenius(frobenius(f12)))  # a^(p + p^2 + p^3 + p^4)
    norm = mul(a, conj)
    assert norm[1:] == (0, 0, 0, 0) and norm[0], "inverse of zero"
    return scal(conj, pow(norm[0], -1, P))


def poly_eval(coeffs, x):
    acc = ZERO
    for c in reversed(coeffs):
        acc = add(mul(acc, x), c)
    return acc


def lagrange_interpolation_at_integers(values):
    """coefficients of the unique polynomial of degree < len(values) with p(i) = values[i], i = 0..len-1
    (DensePolynomial::lagrange_interpolation, crates/backend/poly/src/dense_poly.rs:33)."""
    n = len(values)
    coeffs = [ZERO] * n
    for i, v in enumerate(values):
        # numerator polynomial prod_{j != i} (X - j), base-field coefficients
        num = [1]
        denom = 1
        for j in range(n):
            if j == i:
                continue
            num = [(a - j * b) % P for a, b in zip([0] + num, num + [0])]
            denom = denom * (i - j) % P
        dinv = pow(denom, -1, P)
        for k in range(n):
            coeffs[k] = add(coeffs[k], scal(v, num[k] * dinv % P))
    return coeffs


def two_adic_generator(bits: int) -> int:
    """canonical generator of the 2^bits-th roots of unity: successive squares of the order-2^24 generator
    (crates/backend/koala-bear/src/koala_bear.rs, TWO_ADIC_GENERATORS)"""
    assert 0 <= bits <= 24
    return pow(0x6AC49F88, 1 << (24 - bits), P)


# ---- numpy batches of canonical EF values (arrays [..., 5] of uint64) ----------------------------------------
def np_from_monty(words) -> np.ndarray:
    return (np.asarray(words, dtype=np.uint64) * np.uint64(_RINV)) % np.uint64(P)


def np_to_monty(canon) -> np.ndarray:
    return ((np.asarray(canon, dtype=np.uint64) * np.uint64(_R)) % np.uint64(P)).astype(np.uint32)


def np_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """elementwise product of EF arrays (broadcasting over the leading axes)"""
    p = np.uint64(P)
    d = [None] * 9
    for i in range(5):
        for j in range(5):
            t = (a[..., i] * b[..., j]) % p
            d[i + j] = t if d[i + j] is None else d[i + j] + t
    pp = np.uint64(2 * P)
    out = np.stack([d[0] + d[5] + (pp - d[8] % p), d[1] + d[6], d[2] + d[7] + d[8] + (pp - d[5] % p),
                    d[3] + d[8] + (pp - d[6] % p), d[4] + (pp - d[7] % p)], axis=-1)
    return out % p


def np_mle_eval_rows(rows: np.ndarray, point) -> np.ndarray:
    """rows: [n, 2^k, 5] canonical; point: k canonical EF tuples, first coordinate = most significant index bit
    (MleRef::evaluate on each row, crates/backend/poly/src/evals.rs:142)"""
    p = np.uint64(P)
    cur = rows
    for x in point:
        h = cur.shape[1] // 2
        lo, hi = cur[:, :h], cur[:, h:]
        xv = np.array(x, dtype=np.uint64).reshape(1, 1, 5)
        cur = (lo + np_mul(hi + (p - lo), xv)) % p
    return cur[:, 0]

#!/usr/bin/env python3
"""Generate leanmultisig_b200/csrc/poseidon1_tables.inc.

Derives, with exact integer arithmetic mod p, every constant the CUDA Poseidon1
permutation needs and writes them as Montgomery-form u32 literals so that nvcc sees
compile-time immediates:

  * full-round constants (8 x 16)
  * the sparse partial-round factorisation: first_rc (16), dense transition m_i
    (16 x 16), per-round first rows (20 x 16), rank-1 column vectors v (20 x 15),
    lane-0 scalar constants (19)

The factorisation follows the construction the reference performs at start-up
(crates/backend/koala-bear/src/poseidon1_koalabear_16.rs:399-505,575-600): it is the
standard Poseidon "equivalent sparse matrices" rewrite, so the permutation's outputs
are unchanged (tests/test_poseidon1_consts.py checks the generated tables against the
oracle's dense permutation on CPU).

Input data: leanmultisig_b200/csrc/poseidon1_rc.inc (canonical round constants).
"""
from __future__ import annotations

import os
import re
import sys

P = 0x7F000001
R = (1 << 32) % P
W, RF_HALF, RP = 16, 4, 20
MDS_COL = [1, 3, 13, 22, 67, 2, 15, 63, 101, 1, 2, 17, 11, 1, 51, 1]

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "leanmultisig_b200", "csrc")


def load_rc():
    txt = open(os.path.join(CSRC, "poseidon1_rc.inc")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", txt)]
    assert len(vals) == (2 * RF_HALF + RP) * W
    return [vals[r * W:(r + 1) * W] for r in range(2 * RF_HALF + RP)]


def mat_mul(a, b):
    n = len(a)
    return [[sum(a[i][k] * b[k][j] for k in range(n)) % P for j in range(n)] for i in range(n)]


def mat_vec(m, v):
    return [sum(m[i][j] * v[j] for j in range(len(v))) % P for i in range(len(m))]


def mat_inv(m):
    n = len(m)
    a = [row[:] + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(m)]
    for c in range(n):
        piv = next(r for r in range(c, n) if a[r][c] % P)
        a[c], a[piv] = a[piv], a[c]
        inv = pow(a[c][c], -1, P)
        a[c] = [x * inv % P for x in a[c]]
        for r in range(n):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % P for x, y in zip(a[r], a[c])]
    return [row[n:] for row in a]


def derive():
    rc = load_rc()
    mds = [[MDS_COL[(i - j) % W] for j in range(W)] for i in range(W)]
    partial = rc[RF_HALF:RF_HALF + RP]

    # Push the partial-round constant vectors backwards through MDS^-1: afterwards only
    # lane 0 carries a per-round constant and one full vector is added up front.
    mds_inv = mat_inv(mds)
    opt = [0] * RP
    tmp = partial[RP - 1][:]
    for i in range(RP - 2, -1, -1):
        back = mat_vec(mds_inv, tmp)
        opt[i + 1] = back[0]
        tmp = partial[i][:]
        for j in range(1, W):
            tmp[j] = (tmp[j] + back[j]) % P
    first_rc = tmp
    scalar_rc = opt[1:]

    # Sparse factorisation M^RP = m_i * S_0 * ... * S_{RP-1}, S_r = [[m00, w_hat^T], [v, I]].
    mds_t = [list(col) for col in zip(*mds)]
    m_mul = [row[:] for row in mds_t]
    vs, ws = [], []
    m_i = None
    for _ in range(RP):
        v = [m_mul[0][j + 1] for j in range(W - 1)]
        w = [m_mul[i + 1][0] for i in range(W - 1)]
        hat_inv = mat_inv([row[1:] for row in m_mul[1:]])
        w_hat = mat_vec(hat_inv, w)
        vs.append(v)
        ws.append(w_hat)
        m_i = [row[:] for row in m_mul]
        m_i[0][0] = 1
        for k in range(1, W):
            m_i[k][0] = 0
            m_i[0][k] = 0
        m_mul = mat_mul(mds_t, m_i)
    m_i = [list(col) for col in zip(*m_i)]
    vs.reverse()
    ws.reverse()
    first_row = [[mds[0][0]] + w_hat for w_hat in ws]
    return dict(rc=rc, first_rc=first_rc, scalar_rc=scalar_rc, m_i=m_i, first_row=first_row, v=vs)


def permute_sparse(state, c):
    """Pure-Python evaluation of the permutation from the derived tables (canonical ints)."""
    s = [x % P for x in state]
    mds = [[MDS_COL[(i - j) % W] for j in range(W)] for i in range(W)]

    def full(s, rc):
        s = [pow((x + k) % P, 3, P) for x, k in zip(s, rc)]
        return mat_vec(mds, s)

    for r in range(RF_HALF):
        s = full(s, c["rc"][r])
    s = [(x + k) % P for x, k in zip(s, c["first_rc"])]
    s = mat_vec(c["m_i"], s)
    for r in range(RP):
        s0 = pow(s[0], 3, P)
        if r < RP - 1:
            s0 = (s0 + c["scalar_rc"][r]) % P
        s[0] = s0
        dot = sum(a * b for a, b in zip(s, c["first_row"][r])) % P
        for i in range(1, W):
            s[i] = (s[i] + s0 * c["v"][r][i - 1]) % P
        s[0] = dot
    for r in range(RF_HALF):
        s = full(s, c["rc"][RF_HALF + RP + r])
    return s


def rpow(x, e):
    """x * R^e mod p (e may be negative)."""
    return (x % P) * pow(R, e, P) % P


# Scale drift of the cheap full-round arithmetic (see poseidon1.cuh).  A lane is held as v * s for a known field
# element s: the Montgomery cube maps s -> s^3 R^-2, the MDS is evaluated without the two halvings of its
# even/odd (Karatsuba) splitting, s -> 4 s, and the single Montgomery reduction gives s -> s R^-1.  Every step is
# homogeneous, so constants are simply pre-multiplied by the scale in force where they are added.
MDS_GAIN = 4
RINV = pow(R, -1, P)


def full_round_scales(s):
    """(scale at which the next round constant is added inside the MDS accumulator, scale after the reduction)"""
    acc_scale = pow(s, 3, P) * RINV * RINV * MDS_GAIN % P
    return acc_scale, acc_scale * RINV % P


def restructure(c):
    """Tables for the CUDA kernel's formulation.

    Partial rounds are rewritten so that lanes 1..15 are never materialised per round:
      x'      = state after the 4 initial full rounds + first_rc            (held at scale s4)
      s0_0    = (m_i x')_0
      z_k     = s0_k^3                                                      (k = 0..19)
      s0_{r+1}= fr[r][0] z_r + D_r + sum_{k<r} g[r][k] z_k
      D_r     = sum_{i>=1} fr[r][i] (m_i x')_i   + (all lane-0 scalar constants pushed through)
      lane_i  = (m_i x')_i + sum_k v[k][i-1] z_k + const_i                  (i = 1..15)
    which is the same linear algebra as the reference's sparse loop
    (poseidon1_koalabear_16.rs:893-906) with the sums regrouped.
    """
    rc, m_i, fr, v, sc = c["rc"], c["m_i"], c["first_row"], c["v"], c["scalar_rc"]
    sc = sc + [0]  # no constant after the last partial round
    t = {}
    t["RC0"] = [x * R % P for x in rc[0]]
    # constants folded into the MDS accumulators of initial rounds 0..2 (for rounds 1..3), then first_rc
    s = R
    t["RC_INIT"] = []
    for r in range(4):
        acc_scale, s = full_round_scales(s)
        nxt = rc[r + 1] if r < 3 else c["first_rc"]
        t["RC_INIT"].append([x * acc_scale % P for x in nxt])
    s4 = s
    lin = R * R % P * pow(s4, -1, P) % P  # constant factor that brings (held x') * const * R^-1 to Montgomery scale R
    # s0_0 and D_r as linear forms in x'
    g_rows = [m_i[0][:]]
    for r in range(RP):
        g_rows.append([sum(fr[r][i] * m_i[i][j] for i in range(1, W)) % P for j in range(W)])
    t["G"] = [[x * lin % P for x in row] for row in g_rows]           # 21 x 16
    gtri = [[sum(fr[r][i] * v[k][i - 1] for i in range(1, W)) % P for k in range(r)] for r in range(RP)]
    # constant part of s0_{r+1}: fr0_r sc_r + sum_{k<r} g[r][k] sc_k   (held at R^2 inside the accumulator)
    t["G_CONST"] = [0] + [rpow(fr[r][0] * sc[r] + sum(gtri[r][k] * sc[k] for k in range(r)), 2) for r in range(RP)]
    t["FR0"] = [rpow(fr[r][0], 1) for r in range(RP)]
    t["GTRI"] = [[rpow(x, 1) for x in row] + [0] * (RP - len(row)) for row in gtri]
    # final lanes 1..15: (m_i x')_i + sum_k v[k][i-1] (z_k + sc_k) + rc_terminal0[i]
    t["MI"] = [[x * lin % P for x in m_i[i]] for i in range(1, W)]      # 15 x 16
    t["V"] = [[rpow(v[k][i - 1], 1) for k in range(RP)] for i in range(1, W)]  # 15 x 20 (lane-major)
    rct = rc[RF_HALF + RP:]
    t["LANE_CONST"] = [rpow(sum(v[k][i - 1] * sc[k] for k in range(RP)) + rct[0][i], 2) for i in range(1, W)]
    # lane 0 after the last partial round also needs the first terminal round constant
    t["G_CONST"][RP] = (t["G_CONST"][RP] + rpow(rct[0][0], 2)) % P
    s = R
    t["RC_TERM"] = []
    for r in range(4):
        acc_scale, s = full_round_scales(s)
        if r < 3:
            t["RC_TERM"].append([x * acc_scale % P for x in rct[r + 1]])
    t["FIX"] = R * R % P * pow(s, -1, P) % P  # held * FIX * R^-1 = v * R
    return t


def redc(x):
    return x * pow(R, -1, P) % P


def permute_restructured(state_monty, t):
    """Integer model of the CUDA kernel (Montgomery-form in, Montgomery-form out)."""
    def sbox(a):
        return redc(redc(a * a) * a)

    def mds_redc(a3, init):
        return [redc(init[i] + MDS_GAIN * sum(MDS_COL[(i - j) % W] * a3[j] for j in range(W))) for i in range(W)]

    zeros = [0] * W
    s = [(x + k) % P for x, k in zip(state_monty, t["RC0"])]
    for r in range(4):
        s = mds_redc([sbox(a) for a in s], t["RC_INIT"][r])
    x = s
    z = []
    s0 = redc(sum(t["G"][0][j] * x[j] for j in range(W)))
    D = [t["G_CONST"][r + 1] + sum(t["G"][r + 1][j] * x[j] for j in range(W)) for r in range(RP)]
    for r in range(RP):
        z.append(sbox(s0))
        s0 = redc(D[r] + t["FR0"][r] * z[r] + sum(t["GTRI"][r][k] * z[k] for k in range(r)))
    lanes = [s0] + [redc(t["LANE_CONST"][i] + sum(t["MI"][i][j] * x[j] for j in range(W)) +
                         sum(t["V"][i][k] * z[k] for k in range(RP))) for i in range(W - 1)]
    s = lanes
    for r in range(4):
        s = mds_redc([sbox(a) for a in s], t["RC_TERM"][r] if r < 3 else zeros)
    return [redc(a * t["FIX"]) for a in s]


def cfmt(vals, per_line=8):
    lines = []
    for i in range(0, len(vals), per_line):
        lines.append("    " + ", ".join("0x%08xu" % (v % P) for v in vals[i:i + per_line]) + ",")
    return "\n".join(lines)


def emit(t) -> str:
    out = []
    out.append("// GENERATED by tools/gen_poseidon1_consts.py — do not edit.")
    out.append("// Poseidon1-KoalaBear-16 tables for the restructured permutation in poseidon1.cuh.")
    out.append("// Instance: crates/backend/koala-bear/src/poseidon1_koalabear_16.rs:11-22,699-815 (reference).")
    out.append("// Every entry is an integer in [0,p) already multiplied by the power of R = 2^32 its use site needs.")

    def arr2(name, rows):
        out.append("  /* %s */ {" % name)
        for row in rows:
            out.append("  {\n" + cfmt(row) + "\n  },")
        out.append("  },")

    def arr1(name, row):
        out.append("  /* %s */ {\n%s\n  }," % (name, cfmt(row)))

    out.append("{")
    arr1("RC0[16]", t["RC0"])
    arr2("RC_INIT[4][16]", t["RC_INIT"])
    arr2("G[21][16]", t["G"])
    arr1("G_CONST[21]", t["G_CONST"])
    arr1("FR0[20]", t["FR0"])
    arr2("GTRI[20][20]", t["GTRI"])
    arr2("MI[15][16]", t["MI"])
    arr2("V[15][20]", t["V"])
    arr1("LANE_CONST[15]", t["LANE_CONST"])
    arr2("RC_TERM[3][16]", t["RC_TERM"])
    out.append("  /* FIX */ 0x%08xu," % t["FIX"])

    def darr2(name, rows):
        out.append("  /* %s (same integers as doubles, for the FP64-pipe MDS) */ {" % name)
        for row in rows:
            out.append("  {" + ", ".join("%d.0" % (v % P) for v in row) + "},")
        out.append("  },")

    darr2("RC_INIT_D[4][16]", t["RC_INIT"])
    darr2("RC_TERM_D[3][16]", t["RC_TERM"])
    out.append("}")
    return "\n".join(out) + "\n"


def emit_sparse(c) -> str:
    """Plain sparse-form tables (Montgomery form) for the Poseidon16 AIR and trace generator (air_tables.cuh):
    the column values of the lean_vm Poseidon table are DEFINED through this formulation
    (crates/lean_vm/src/tables/poseidon_16/{mod.rs:384-420, trace_gen.rs:46-95})."""
    out = ["// GENERATED by tools/gen_poseidon1_consts.py — do not edit.",
           "// Poseidon1-KoalaBear-16 sparse-form constants, Montgomery form (x * 2^32 mod p)."]
    m = lambda row: [x * R % P for x in row]
    rc = c["rc"]
    out.append("{")
    out.append("  /* RC_FULL[8][16]: initial rounds 0..3, final rounds 0..3 */ {")
    for r in list(range(RF_HALF)) + list(range(RF_HALF + RP, 2 * RF_HALF + RP)):
        out.append("  {\n" + cfmt(m(rc[r])) + "\n  },")
    out.append("  },")
    out.append("  /* FIRST_RC[16] */ {\n%s\n  }," % cfmt(m(c["first_rc"])))
    out.append("  /* M_I[16][16] */ {")
    for row in c["m_i"]:
        out.append("  {\n" + cfmt(m(row)) + "\n  },")
    out.append("  },")
    out.append("  /* FIRST_ROW[20][16] */ {")
    for row in c["first_row"]:
        out.append("  {\n" + cfmt(m(row)) + "\n  },")
    out.append("  },")
    out.append("  /* V[20][16] (entry 15 unused) */ {")
    for row in c["v"]:
        out.append("  {\n" + cfmt(m(row + [0])) + "\n  },")
    out.append("  },")
    out.append("  /* SCALAR_RC[20] (entry 19 unused) */ {\n%s\n  }," % cfmt(m(c["scalar_rc"] + [0])))
    out.append("}")
    return "\n".join(out) + "\n"


def main():
    c = derive()
    t = restructure(c)
    ok = True
    for name, txt in (("poseidon1_tables.inc", emit(t)), ("poseidon1_sparse_tables.inc", emit_sparse(c))):
        path = os.path.join(CSRC, name)
        if "--check" in sys.argv:
            ok = ok and os.path.exists(path) and open(path).read() == txt
        else:
            open(path, "w").write(txt)
            print("wrote", path)
    if "--check" in sys.argv:
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

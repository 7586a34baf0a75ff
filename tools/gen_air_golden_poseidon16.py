"""Golden per-constraint values of the poseidon16 AIR (100 constraints, 109 columns), produced BY EXECUTING THE REFERENCE'S
SOURCE TEXT of crates/lean_vm/src/tables/poseidon_16/mod.rs — the third table next to tools/gen_air_golden.py's two.

What is translated mechanically (statement syntax only, with the Rust-subset translator of tools/gen_whir_config_golden.py):
  * `Air::eval` of Poseidon16Precompile (mod.rs:316-362), `eval_poseidon1_16`, `eval_2_full_rounds_16`,
    `eval_last_2_full_rounds_16`, `dense_mat_vec_air_16`, `sparse_mat_air_16` (mod.rs:383-548);
  * the column layout: the field list of `#[repr(C)] struct Poseidon1Cols16<T>` (mod.rs:364-381) is parsed and `flat` is cut
    into those fields in declaration order (what `align_to` does).
Reference elimination is syntactic: `for (s, r) in A.iter_mut().zip(B.iter())` becomes an index loop with `*s` -> A[k],
`*r` -> B[k]; `add_kb(x, c)` (mod.rs:48-63: `*x += c`) becomes `x = x + c`; `mul_kb(a, c)` (mod.rs:67-82) is `a * c`;
`builder.low_degree_block(&mut state, |b, state| {..})` (air/src/lib.rs:73-78: `block(self, state)`) is inlined with b = builder.
What is NOT taken from the source text: `mds_air_16` is the product with the circulant MDS matrix whose first column is parsed
from koala-bear's MDS constant, and the sparse partial-round tables (`poseidon1_sparse_*`) come from this repository's
derivation tools/gen_poseidon1_consts.py::derive() — the reference computes them at start-up (poseidon1_koalabear_16.rs:399-481)
and stores no copy; the round constants are parsed from the reference (poseidon1_koalabear_16.rs:699-815) and must equal the
table that derivation starts from.  The constraint polynomials over the committed columns do not depend on which equivalent
sparse factorisation is used as long as it reproduces the permutation, which the KAT pins.

    python tools/gen_air_golden_poseidon16.py [--check]     # tests/golden/air_constraints_poseidon16.json
"""
from __future__ import annotations

import json
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_air_golden as G  # noqa: E402  Fp, Ef, Builder, consts_of
import gen_poseidon1_consts as PC  # noqa: E402
import gen_whir_config_golden as T  # noqa: E402  Emitter, U, match_close, split_top

P = 0x7F000001
REF = "/root/reference/crates"
SRC = f"{REF}/lean_vm/src/tables/poseidon_16/mod.rs"
KB = f"{REF}/backend/koala-bear/src/poseidon1_koalabear_16.rs"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "air_constraints_poseidon16.json")
FUNCS = ["eval", "eval_poseidon1_16", "eval_2_full_rounds_16", "eval_last_2_full_rounds_16", "dense_mat_vec_air_16",
         "sparse_mat_air_16"]


class Fp(G.Fp):
    def cube(self):
        return self * self * self


G.Fp.cube = Fp.cube  # results of G.Fp arithmetic are G.Fp


class U(T.U):
    __truediv__ = T.U.__floordiv__  # usize division


class Builder(G.Builder):
    def assert_eq(self, x, y):          # air/src/lib.rs:59-61
        self.assert_zero(x - y)

    def assert_eq_low(self, x, y):      # air/src/lib.rs:67-69
        self.assert_eq(x, y)


def fn_text(src: str, name: str):
    m = re.search(r"\bfn\s+" + name + r"\s*(<[^(]*>)?\s*\(", src)
    po = src.index("(", m.end() - 1)
    pc = T.match_close(src, po)
    params = []
    for p in T.split_top(src[po + 1:pc], ","):
        p = p.strip()
        if p:
            params.append("self" if p in ("&self", "self") else p.split(":")[0].strip().replace("mut ", ""))
    bo = src.index("{", pc)
    return params, src[bo + 1:T.match_close(src, bo)]


def deref_zip_loops(body: str) -> str:
    """for (x, y) in A.iter_mut().zip(B[.iter()]) { .. }  ->  for _k in 0..WIDTH { .. } with the references resolved"""
    while True:
        m = re.search(r"for\s+\((\w+),\s*(&?)(\w+)\)\s+in\s+(\w+)\.iter_mut\(\)\.zip\(([\w.]+?)(?:\.iter\(\))?\)\s*\{", body)
        if not m:
            return body
        x, amp, y, a, b = m.groups()
        o = m.end() - 1
        c = T.match_close(body, o)
        inner = body[o + 1:c]
        inner = re.sub(r"add_kb\(" + x + r",\s*", f"add_kb({a}[_k], ", inner)
        inner = re.sub(r"\*" + x + r"\b", f"{a}[_k]", inner)
        inner = re.sub(r"\b" + x + r"\.", f"{a}[_k].", inner)
        inner = re.sub((r"\b" if amp else r"\*") + y + r"\b", f"{b}[_k]", inner)
        body = body[:m.start()] + "for _k in 0..WIDTH {" + inner + "}" + body[c + 1:]


def prepare(body: str) -> str:
    body = T.strip_comments(body)
    # the #[repr(C)] view of `flat`
    i = body.find("let cols:")
    if i >= 0:
        o = body.index("{", i)
        c = T.match_close(body, o)
        assert body[c + 1:].lstrip().startswith(";")
        body = body[:i] + "let cols = make_cols(builder.flat())" + body[c + 1:]
    # closure of low_degree_block, inlined
    m = re.search(r"builder\.low_degree_block\(&mut state,\s*\|b,\s*state\|\s*\{", body)
    if m:
        o = m.end() - 1
        c = T.match_close(body, o)
        tail = body[c + 1:].lstrip()
        assert tail.startswith(");")
        body = body[:m.start()] + "let b = builder;" + body[o + 1:c] + body[body.index(");", c) + 2:]
    body = re.sub(r"let\s+state:\s*&mut\s*\[AB::IF;\s*WIDTH\]\s*=\s*state\.try_into\(\)\.unwrap\(\);", "", body)
    body = deref_zip_loops(body)
    body = re.sub(r"let\s+mut\s+(\w+):\s*\[_;\s*WIDTH\]\s*=\s*([\w.]+);", r"let mut \1 = list(\2);", body)
    body = re.sub(r"let\s+(\w+)\s*=\s*\*(\w+);", r"let \1 = list(\2);", body)
    body = re.sub(r"add_kb\((?:&mut\s+)?([^,;]+?),\s*([^;]+?)\);", r"\1 = \1 + (\2);", body)
    body = re.sub(r"(?m)^(\s*)([\w\[\] +\-*/]+?)\s*\+=\s*([^;]+);", r"\1\2 = \2 + (\3);", body)
    body = re.sub(r"::<[^>]*>", "", body)                              # turbofish
    body = body.replace("std::slice::from_ref(", "list_of(")
    body = re.sub(r"&mut\s+", "", body)
    body = re.sub(r"&(?=[\[a-z])", "", body)
    body = re.sub(r"\bAB::IF::ONE\b|\bA::ONE\b", "Fp(1)", body)
    body = re.sub(r"\bA::ZERO\b|\bAB::IF::ZERO\b", "Fp(0)", body)
    body = body.replace("AB::F::from_usize(", "Fp(")
    body = re.sub(r"let\s+mut\s+", "let ", body)
    return body


def translate(src: str) -> str:
    out = []
    for name in FUNCS:
        params, body = fn_text(src, name)
        em = T.Emitter("Poseidon16")
        em.emit(0, f"def {name}({', '.join(params)}):")
        em.block(prepare(body), 1, "return")
        out += em.lines + [""]
    return "\n".join(out)


def struct_fields(src: str, consts: dict):
    i = src.index("struct Poseidon1Cols16<T>")
    o = src.index("{", i)
    fields = []
    for f in T.split_top(T.strip_comments(src[o + 1:T.match_close(src, o)]), ","):
        f = f.strip()
        if not f:
            continue
        name, ty = f.replace("pub ", "").split(":", 1)
        dims = [int(eval(d.replace("/", "//"), {"__builtins__": {}}, dict(consts))) for d in re.findall(r";\s*([^\];]+)\]", ty)]
        fields.append((name.strip(), dims))  # `[[T; W]; N]` lists the inner dimension first
    return fields


def make_cols_fn(fields):
    def make_cols(flat):
        class Cols:
            pass
        cols, pos = Cols(), 0
        for name, dims in fields:
            if not dims:
                setattr(cols, name, flat[pos])
                pos += 1
            elif len(dims) == 1:
                setattr(cols, name, list(flat[pos:pos + dims[0]]))
                pos += dims[0]
            else:
                inner, outer = dims
                setattr(cols, name, [list(flat[pos + r * inner:pos + (r + 1) * inner]) for r in range(outer)])
                pos += inner * outer
        assert pos == len(flat), (pos, len(flat))
        return cols
    return make_cols


def reference_round_constants():
    txt = open(KB).read()
    i = txt.index("const POSEIDON1_RC:")
    o = txt.index("new_2d_array(", i)
    c = txt.index("]);", o)
    vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", T.strip_comments(txt[o:c]))]
    assert len(vals) == 28 * 16, len(vals)
    return [vals[16 * r:16 * r + 16] for r in range(28)]


def mds_first_column():
    txt = open(f"{REF}/backend/koala-bear/src/poseidon1_koalabear_16.rs").read()
    m = re.search(r"const MDS_CIRC_COL[^=]*=\s*KoalaBear::new_array\(\[([^\]]*)\]", txt)
    col = [int(x) for x in re.findall(r"\d+", m.group(1))]
    assert len(col) == 16
    return col


def generate():
    src = open(SRC).read()
    consts = G.consts_of(KB, SRC, f"{REF}/lean_vm/src/core/constants.rs", f"{REF}/lean_vm/src/tables/mod.rs")
    for m in re.finditer(r"(?m)^(?:pub(?:\([a-z]+\))?\s+)?const\s+([A-Z0-9_]+):\s*usize\s*=\s*([^;]+);", src):
        if m.group(1) not in consts:
            consts[m.group(1)] = int(eval(m.group(2).replace("/", "//"), {"__builtins__": {}}, dict(consts)))
    fields = struct_fields(src, consts)
    n_cols = sum(1 if not d else (d[0] if len(d) == 1 else d[0] * d[1]) for _, d in fields)
    code = translate(src)
    tabs = PC.derive()
    rc = reference_round_constants()
    assert rc == [list(r) for r in tabs["rc"]], "round constants differ from the reference's table"
    col = mds_first_column()
    half = consts["POSEIDON1_HALF_FULL_ROUNDS"]
    rp = consts["POSEIDON1_PARTIAL_ROUNDS"]
    fp = lambda t: [Fp(x) for x in t]  # noqa: E731

    def mds_air_16(state):  # mod.rs:12-30: the circulant MDS layer, y_i = sum_j col[(i - j) mod 16] x_j
        x = list(state)
        for i in range(16):
            acc = Fp(0)
            for j in range(16):
                acc = acc + x[j] * col[(i - j) % 16]
            state[i] = acc

    rng = random.Random(20261017)
    flat = [Fp(rng.randrange(P)) for _ in range(n_cols)]
    la = [G.Ef([rng.randrange(P) for _ in range(5)]) for _ in range(8)]
    beta = G.Ef([rng.randrange(P) for _ in range(5)])

    def eval_virtual_bus_column(_extra, flag, data):  # tables/utils.rs:5-21
        s = G.Ef([0] * 5)
        for c, d in zip(la, data):
            s = s + c * d
        return (s + la[-1] * Fp(consts["LOGUP_PRECOMPILE_DOMAINSEP"])) * beta + flag

    b = Builder(flat, [])
    env = {"Fp": Fp, "U": U, "BUS": True, "list": list, "len": len, "range": range, "make_cols": make_cols_fn(fields),
           "mds_air_16": mds_air_16, "mul_kb": lambda a, c: a * c, "eval_virtual_bus_column": eval_virtual_bus_column,
           "list_of": lambda x: [x],
           "poseidon1_initial_constants": lambda: [fp(r) for r in rc[:half]],
           "poseidon1_final_constants": lambda: [fp(r) for r in rc[half + rp:]],
           "poseidon1_sparse_first_round_constants": lambda: fp(tabs["first_rc"]),
           "poseidon1_sparse_m_i": lambda: [fp(r) for r in tabs["m_i"]],
           "poseidon1_sparse_first_row": lambda: [fp(r) for r in tabs["first_row"]],
           "poseidon1_sparse_v": lambda: [fp(list(r) + [0]) for r in tabs["v"]],
           "poseidon1_sparse_scalar_round_constants": lambda: fp(tabs["scalar_rc"])}
    env.update({k: U(v) for k, v in consts.items()})
    exec(compile(code, "<poseidon_16/mod.rs translated>", "exec"), env)
    env["eval"](None, b, None)
    return {
        "note": "generated by tools/gen_air_golden_poseidon16.py from the reference's poseidon_16/mod.rs source text; canonical "
                "residues mod 2^31 - 2^24 + 1",
        "tables": [{
            "table": "poseidon16", "source": os.path.relpath(SRC, "/root/reference"), "flat": [x.v for x in flat], "shift": [],
            "columns": [[n, d] for n, d in fields],
            "logup_alphas_eq_poly": [x.c for x in la], "bus_beta": beta.c,
            "constraints": [{"kind": k, "value": v} for k, v in b.log], "translated_python": code.split("\n"),
        }],
    }


if __name__ == "__main__":
    g = generate()
    t = g["tables"][0]
    print(t["table"], len(t["flat"]), "columns,", len(t["constraints"]), "constraints")
    if "--check" in sys.argv:
        assert json.load(open(OUT)) == g, "tests/golden/air_constraints_poseidon16.json is stale"
    else:
        json.dump(g, open(OUT, "w"), indent=1)
        print("wrote", OUT)

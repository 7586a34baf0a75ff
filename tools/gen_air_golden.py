"""Golden per-constraint values of the lean_vm AIRs, produced BY EXECUTING THE REFERENCE'S SOURCE TEXT.

The reference (Rust) cannot be compiled in this image, but the `Air::eval` bodies of the execution table and of the
extension_op precompile are straight-line field arithmetic over `flat[...]` / `shift[...]`.  This script reads them from
`/root/reference`, translates the statement syntax mechanically (let -> assignment, `std::array::from_fn(|k| ..)` -> list
comprehension, `for k in 0..5 {..}` -> loop, `AB::F::ONE` -> constants, Rust if-expressions -> Python conditional
expressions; no formula is retyped) and runs the result on a random point with a builder that records every
`assert_zero` / `assert_bool` / `assert_zero_ef` in call order.  The recorded values pin, independently of oracle/air.c
and of the CUDA kernels,
  * the ORDER of the constraints (= which alpha power multiplies which constraint), and
  * every constraint's polynomial, at a random point.
`eval_virtual_bus_column` (crates/lean_vm/src/tables/utils.rs:5-21) and `quintic_mul_air` are evaluated from their
definitions: (sum_i la_i data_i + la_last * DOMAINSEP) * beta + flag, and the product in F[X]/(X^5 + X^2 - 1).

    python tools/gen_air_golden.py            # rewrites tests/golden/air_constraints.json (needs /root/reference)

tests/test_air_golden.py compares the oracle (CPU tier) and the CUDA sessions (GPU tier) with the file, and - when
/root/reference is present - re-runs this translation and checks the committed file is current.
"""
from __future__ import annotations

import json
import os
import random
import re
import sys

P = 0x7F000001
REF = "/root/reference/crates/lean_vm/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "air_constraints.json")


class Fp:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v % P

    def _c(self, o):
        return o if isinstance(o, Fp) else Fp(o)

    def __add__(self, o):
        if isinstance(o, Ef):
            return o + self
        return Fp(self.v + self._c(o).v)

    __radd__ = __add__

    def __sub__(self, o):
        return Fp(self.v - self._c(o).v)

    def __rsub__(self, o):
        return Fp(self._c(o).v - self.v)

    def __mul__(self, o):
        if isinstance(o, Ef):
            return o * self
        return Fp(self.v * self._c(o).v)

    __rmul__ = __mul__

    def __neg__(self):
        return Fp(-self.v)

    def halve(self):
        return Fp(self.v * pow(2, -1, P))

    def double(self):
        return Fp(2 * self.v)

    def bool_check(self):  # field.rs:207-210: x * (1 - x)
        return self * (Fp(1) - self)


class Ef:
    """F[X]/(X^5 + X^2 - 1), schoolbook"""

    def __init__(self, c):
        self.c = [x % P for x in c]

    def __add__(self, o):
        if isinstance(o, Ef):
            return Ef([a + b for a, b in zip(self.c, o.c)])
        v = o.v if isinstance(o, Fp) else o
        return Ef([self.c[0] + v] + self.c[1:])

    __radd__ = __add__

    def __mul__(self, o):
        if not isinstance(o, Ef):
            v = o.v if isinstance(o, Fp) else o
            return Ef([a * v for a in self.c])
        d = [0] * 9
        for i, a in enumerate(self.c):
            for j, b in enumerate(o.c):
                d[i + j] += a * b
        for k in range(8, 4, -1):  # X^k = X^(k-5) - X^(k-3)
            d[k - 5] += d[k]
            d[k - 3] -= d[k]
            d[k] = 0
        return Ef(d[:5])

    __rmul__ = __mul__


def quintic_mul_air(a, b):
    """product of two elements given by 5 base coordinates each, coordinates are AIR values (Fp here)"""
    d = [Fp(0)] * 9
    for i in range(5):
        for j in range(5):
            d[i + j] = d[i + j] + a[i] * b[j]
    for k in range(8, 4, -1):
        d[k - 5] = d[k - 5] + d[k]
        d[k - 3] = d[k - 3] - d[k]
    return d[:5]


class Builder:
    def __init__(self, flat, shift):
        self._flat, self._shift, self.log = flat, shift, []

    def flat(self):
        return self._flat

    def shift(self):
        return self._shift

    def assert_zero(self, x):
        self.log.append(("base", x.v))

    def assert_bool(self, x):
        self.assert_zero(x.bool_check())

    def assert_zero_ef(self, x):
        self.log.append(("ext", list(x.c)))

    def declare_values(self, _):
        raise AssertionError("BUS = false branch is not the golden one")


def consts_of(*paths):
    out = {}
    for p in paths:
        for m in re.finditer(r"pub(?:\([a-z]+\))?\s+const\s+([A-Z0-9_]+):\s*usize\s*=\s*([^;]+);", open(p).read()):
            out[m.group(1)] = m.group(2).strip()
    # resolve expressions over earlier constants
    res = {}
    for _ in range(4):
        for k, v in out.items():
            if k in res:
                continue
            try:
                res[k] = int(eval(v, {"__builtins__": {}}, dict(res)))
            except Exception:
                pass
    return res


def eval_body(path):
    txt = open(path).read()
    i = txt.index("fn eval<AB: AirBuilder>")
    i = txt.index("{", i)
    depth, j = 0, i
    while True:
        depth += {"{": 1, "}": -1}.get(txt[j], 0)
        if depth == 0:
            break
        j += 1
    return txt[i + 1:j]


def split_top(s, sep=";"):
    """split at top-level separators, keeping brace blocks (for / if) as single statements"""
    out, depth, cur = [], 0, ""
    k = 0
    while k < len(s):
        ch = s[k]
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        cur += ch
        if depth == 0 and (ch == sep or (ch == "}" and re.match(r"\s*(for|if)\b", cur))):
            # an if block may be followed by else
            rest = s[k + 1:]
            if ch == "}" and re.match(r"\s*else\b", rest):
                k += 1
                continue
            out.append(cur.strip().rstrip(";").strip())
            cur = ""
        k += 1
    if cur.strip():
        out.append(cur.strip())
    return [x for x in out if x]


def tr_expr(e):
    e = re.sub(r"//[^\n]*", "", e)
    e = e.replace("AB::F::ONE", "Fp(1)").replace("AB::F::TWO", "Fp(2)").replace("AB::IF::ONE", "Fp(1)")
    e = re.sub(r"AB::F::from_usize\(\s*(?:crate::)?([A-Za-z0-9_]+)\s*\)", r"Fp(\1)", e)
    e = re.sub(r"eval_virtual_bus_column::<AB, EF>", "eval_virtual_bus_column", e)
    e = e.replace("&[", "[").replace("&", "")
    # closures: std::array::from_fn(|k| BODY)
    while "std::array::from_fn(" in e:
        a = e.index("std::array::from_fn(")
        b = a + len("std::array::from_fn(")
        depth, j = 1, b
        while depth:
            depth += {"(": 1, ")": -1}.get(e[j], 0)
            j += 1
        inner = e[b:j - 1].strip()
        m = re.match(r"\|(\w+)\|\s*(.*)$", inner, re.S)
        var, body = m.group(1), m.group(2).strip()
        e = e[:a] + f"[({tr_closure_body(body)}) for {var} in range(5)]" + e[j:]
    return tr_if(e)


def tr_closure_body(body):
    body = body.strip()
    if body.startswith("{"):
        stmts = split_top(body[1:-1].strip())
        # let x = E; ... ; final expression  ->  nested lambdas
        expr = tr_if(stmts[-1])
        for st in reversed(stmts[:-1]):
            m = re.match(r"let\s+(\w+)\s*=\s*(.*)$", st, re.S)
            expr = f"(lambda {m.group(1)}: {expr})({tr_expr(m.group(2))})"
        return expr
    return tr_if(body)


def tr_if(e):
    """Rust `if c { a } else { b }` expressions -> Python conditional expressions (innermost first)"""
    pat = re.compile(r"if\s+([^{}]+?)\s*\{([^{}]*)\}\s*else\s*\{([^{}]*)\}", re.S)
    while True:
        m = pat.search(e)
        if not m:
            return e
        e = e[:m.start()] + f"(({m.group(2).strip()}) if ({m.group(1).strip()}) else ({m.group(3).strip()}))" + e[m.end():]


def tr_stmts(stmts, indent=""):
    py = []
    for st in stmts:
        st = re.sub(r"//[^\n]*", "", st).strip()
        if not st:
            continue
        m = re.match(r"for\s+(\w+)\s+in\s+(\d+)\.\.(\d+)\s*\{(.*)\}\s*$", st, re.S)
        if m:
            py.append(f"{indent}for {m.group(1)} in range({m.group(2)}, {m.group(3)}):")
            py += tr_stmts(split_top(m.group(4)), indent + "    ")
            continue
        m = re.match(r"if\s+BUS\s*\{(.*)\}\s*else\s*\{(.*)\}\s*$", st, re.S)
        if m:
            py += tr_stmts(split_top(m.group(1)), indent)
            continue
        m = re.match(r"let\s+(\([^)]*\)|\w+)\s*(?::\s*\[[^\]]*\])?\s*=\s*(.*)$", st, re.S)
        if m:
            py.append(f"{indent}{m.group(1)} = {' '.join(tr_expr(m.group(2)).split())}")
            continue
        py.append(f"{indent}{' '.join(tr_expr(st).split())}")
    return py


def run_table(name, path, const_paths, n_cols, n_shift, rng):
    consts = consts_of(path, *const_paths)
    body = eval_body(path)
    code = "\n".join(tr_stmts(split_top(body)))
    flat = [Fp(rng.randrange(P)) for _ in range(n_cols)]
    shift = [Fp(rng.randrange(P)) for _ in range(n_shift)]
    la = [Ef([rng.randrange(P) for _ in range(5)]) for _ in range(8)]
    beta = Ef([rng.randrange(P) for _ in range(5)])

    def eval_virtual_bus_column(_extra, flag, data):  # tables/utils.rs:5-21
        s = Ef([0] * 5)
        for c, d in zip(la, data):
            s = s + c * d
        return (s + la[-1] * Fp(consts["LOGUP_PRECOMPILE_DOMAINSEP"])) * beta + flag

    b = Builder(flat, shift)
    env = {"Fp": Fp, "builder": b, "extra_data": None, "eval_virtual_bus_column": eval_virtual_bus_column,
           "quintic_mul_air": quintic_mul_air, "range": range}
    env.update(consts)
    exec(code, env)
    return {
        "table": name, "source": os.path.relpath(path, "/root/reference"), "flat": [x.v for x in flat], "shift": [x.v for x in shift],
        "logup_alphas_eq_poly": [x.c for x in la], "bus_beta": beta.c,
        "constraints": [{"kind": k, "value": v} for k, v in b.log], "translated_python": code.split("\n"),
    }


def generate():
    rng = random.Random(20260117)
    common = [f"{REF}/core/constants.rs", f"{REF}/tables/extension_op/mod.rs"]
    return {
        "note": "generated by tools/gen_air_golden.py from the reference's Air::eval source text; canonical residues mod 2^31 - 2^24 + 1",
        "tables": [
            run_table("execution", f"{REF}/tables/execution/air.rs", common, 20, 2, rng),
            run_table("extension_op", f"{REF}/tables/extension_op/air.rs", common, 29, 13, rng),
        ],
    }


if __name__ == "__main__":
    g = generate()
    for t in g["tables"]:
        print(t["table"], len(t["constraints"]), "constraints:", [c["kind"] for c in t["constraints"]].count("ext"), "extension-valued")
    if "--check" in sys.argv:
        assert json.load(open(OUT)) == g, "tests/golden/air_constraints.json is stale"
    else:
        json.dump(g, open(OUT, "w"), indent=1)
        print("wrote", OUT)

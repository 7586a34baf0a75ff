"""Golden WHIR parameter schedules, produced BY EXECUTING THE REFERENCE'S SOURCE TEXT of crates/whir/src/config.rs.

The reference (Rust) cannot be compiled in this image.  `WhirConfig::new`, `FoldingFactor` and `SecurityAssumption` are plain
integer / f64 code, so this script translates their statement syntax mechanically into Python (fn -> def, let -> assignment,
`match self` -> if chain, expression-valued if / match -> branches assigning the same name, struct literals -> keyword
constructors, `x as f64` / `as usize` -> casts, float literals -> an f64 wrapper whose methods are Rust's: log2, ceil, sqrt,
powf, powi (square-and-multiply like compiler-rt's __powidf2), max, min) and runs the result.  No formula is retyped; the
translated text is stored next to the values so a reader can diff it against the Rust.

    python tools/gen_whir_config_golden.py        # rewrites tests/golden/whir_config.json (needs /root/reference)

tests/test_whir_config_golden.py compares the product's `WhirConfig` (leanmultisig_b200/whir_config.py), the oracle's
(oracle/whir.py) and — when /root/reference is present — a fresh run of this translation with the committed file.
Parameters: lean_prover/src/lib.rs:22-49 (default_whir_config: 124 bits, 16 grinding bits, folding 7 / 5, initial domain
reduction 5, coefficients sent at <= 8 variables, Johnson bound), EF::bits() = bit length of p^5 = 155.
"""
from __future__ import annotations

import json
import math
import os
import re
import sys

REF = "/root/reference/crates/whir/src/config.rs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "whir_config.json")
P = 0x7F000001


# ------------------------------------------------------------------------------------------ Rust number semantics
class F64(float):
    def _w(f):  # noqa: N805
        def g(self, o):
            return F64(f(float(self), float(o)))
        return g

    __add__ = _w(lambda a, b: a + b)
    __radd__ = _w(lambda a, b: b + a)
    __sub__ = _w(lambda a, b: a - b)
    __rsub__ = _w(lambda a, b: b - a)
    __mul__ = _w(lambda a, b: a * b)
    __rmul__ = _w(lambda a, b: b * a)
    __truediv__ = _w(lambda a, b: a / b)
    __rtruediv__ = _w(lambda a, b: b / a)

    def __neg__(self):
        return F64(-float(self))

    def log2(self):
        return F64(math.log2(self))

    def sqrt(self):
        return F64(math.sqrt(self))

    def ceil(self):
        return F64(math.ceil(self))

    def powf(self, e):
        return F64(math.pow(self, e))

    def powi(self, n):  # compiler-rt __powidf2
        a, b, r = float(self), int(n), 1.0
        recip = b < 0
        b = abs(b)
        while True:
            if b & 1:
                r *= a
            b //= 2
            if b == 0:
                break
            a *= a
        return F64(1.0 / r if recip else r)

    def max(self, o):
        return F64(max(float(self), float(o)))

    def min(self, o):
        return F64(min(float(self), float(o)))


class U(int):
    """usize"""

    def _w(f):  # noqa: N805
        def g(self, o):
            if isinstance(o, float):
                raise TypeError("usize mixed with f64 without a cast")
            r = f(int(self), int(o))
            if r < 0:
                raise OverflowError("usize underflow")
            return U(r)
        return g

    __add__ = _w(lambda a, b: a + b)
    __radd__ = _w(lambda a, b: b + a)
    __sub__ = _w(lambda a, b: a - b)
    __rsub__ = _w(lambda a, b: b - a)
    __mul__ = _w(lambda a, b: a * b)
    __rmul__ = _w(lambda a, b: b * a)
    __lshift__ = _w(lambda a, b: a << b)
    __rlshift__ = _w(lambda a, b: b << a)
    __rshift__ = _w(lambda a, b: a >> b)
    __rrshift__ = _w(lambda a, b: b >> a)
    __floordiv__ = _w(lambda a, b: a // b)

    def saturating_sub(self, o):
        return U(max(int(self) - int(o), 0))

    def div_ceil(self, o):
        return U(-(-int(self) // int(o)))

    def ilog2(self):
        return U(int(self).bit_length() - 1)


def as_usize(x):
    return U(int(x))  # f64 -> usize truncates (the sources only cast ceil()ed values and integers)


def as_f64(x):
    return F64(float(x))


# ------------------------------------------------------------------------------------------ source -> Python
def strip_comments(src: str) -> str:
    return re.sub(r"//[^\n]*", "", src)


def match_close(s: str, i: int) -> int:
    """index of the bracket closing s[i]"""
    pairs = {"(": ")", "[": "]", "{": "}"}
    depth, j = 0, i
    while j < len(s):
        c = s[j]
        if c in pairs:
            depth += 1
        elif c in pairs.values():
            depth -= 1
            if depth == 0:
                return j
        j += 1
    raise ValueError("unbalanced")


def split_top(s: str, sep: str) -> list[str]:
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += c
    if cur.strip():
        out.append(cur)
    return out


def operand_start(s: str, end: int) -> int:
    """start of the postfix expression ending at s[:end] (identifier / call / method chain / parenthesised group)"""
    j = end
    while j > 0:
        c = s[j - 1]
        if c == ")":
            depth, k = 0, j - 1
            while True:
                if s[k] == ")":
                    depth += 1
                elif s[k] == "(":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
            j = k
        elif c.isalnum() or c in "_.":
            j -= 1
        else:
            break
    return j


def tr_casts(e: str) -> str:
    while True:
        m = re.search(r"\s+as\s+(f64|usize)\b", e)
        if not m:
            return e
        a = operand_start(e, m.start())
        e = e[:a] + f"as_{m.group(1)}({e[a:m.start()]})" + e[m.end():]


def tr_struct_literals(e: str, cls: str) -> str:
    while True:
        m = re.search(r"\b(Self|RoundConfig|WhirConfig|FoldingFactor)\s*\{", e)
        if not m:
            return e
        o = e.index("{", m.start())
        c = match_close(e, o)
        fields = []
        for f in split_top(e[o + 1:c], ","):
            f = f.strip()
            if not f:
                continue
            if re.match(r"^\w+$", f):
                fields.append(f"{f}={f}")
            else:
                k, v = f.split(":", 1)
                fields.append(f"{k.strip()}={v.strip()}")
        name = cls if m.group(1) == "Self" else m.group(1)
        e = e[:m.start()] + f"{name}._make(" + ", ".join(fields) + ")" + e[c + 1:]


def tr_expr(e: str, cls: str) -> str:
    e = " ".join(e.split())
    e = re.sub(r"\s+\.", ".", e)
    e = e.replace("PF::<EF>::", "").replace("PF::<EF>::", "")
    e = e.replace("EF::bits()", "EF_BITS")
    e = re.sub(r"matches!\(\s*self\s*,\s*Self::(\w+)\s*\)", r'(self.kind == "\1")', e)
    e = re.sub(r"matches!\(\s*([\w.]+)\s*,\s*SecurityAssumption::(\w+)\s*\)", r'(\1.kind == "\2")', e)
    e = e.replace("usize::MAX", "U(2**64 - 1)").replace("f64::from(", "as_f64(")
    e = re.sub(r"Vec::with_capacity\([^)]*\)", "[]", e)
    e = re.sub(r"\bSelf::(\w+)\(", cls + r".\1(", e)
    e = tr_struct_literals(e, cls)
    e = re.sub(r"\b(\d+)usize\b", r"U(\1)", e)
    e = re.sub(r"\b(\d+)_f64\b", r"F64(\1)", e)
    e = re.sub(r"(?<![\w.)])(\d+\.\d*)(?![\w.])", r"F64(\1)", e)          # 2.0, 1., 0.5
    e = re.sub(r"(?<![\w.)])(\d+\.)(?=[a-z])", r"F64(\1)", e)              # 3.max(..) does not occur; kept for safety
    e = tr_casts(e)
    e = e.replace(".push(", ".append(").replace(".last().unwrap()", "[-1]")
    e = re.sub(r"([\w.]+)\.len\(\)", r"len(\1)", e)
    e = re.sub(r"([\w.]+)\.is_empty\(\)", r"(len(\1) == 0)", e)
    e = e.replace("||", " or ").replace("&&", " and ")
    e = re.sub(r"!(?=[\w(])", "not ", e)
    return e


class Emitter:
    def __init__(self, cls: str):
        self.cls, self.lines = cls, []

    def emit(self, depth: int, text: str):
        self.lines.append("    " * depth + text)

    # a block's statements; `sink` says what happens to the trailing expression: "return", "name =", or None
    def block(self, s: str, depth: int, sink):
        i, n = 0, len(s)
        emitted = False
        while True:
            while i < n and s[i].isspace():
                i += 1
            if i >= n:
                break
            rest = s[i:]
            m = re.match(r"let\s+(mut\s+)?", rest)
            if m:
                j = i + m.end()
                eq = self._find_top(s, j, "=")
                pat = s[j:eq].strip()
                pat = re.sub(r":\s*[\w:<>\[\]; _]+$", "", pat).strip()  # type annotation
                pat = pat.replace("mut ", "")
                k = eq + 1
                while s[k].isspace():
                    k += 1
                if re.match(r"(if|match)\b", s[k:]):
                    k = self.construct(s, k, depth, pat + " =")
                    while s[k].isspace():
                        k += 1
                    assert s[k] == ";", s[k:k + 40]
                    i = k + 1
                else:
                    end = self._find_top(s, k, ";")
                    self.emit(depth, f"{pat} = {tr_expr(s[k:end], self.cls)}")
                    i = end + 1
                emitted = True
                continue
            if re.match(r"(if|match|for)\b", rest):
                # statement position: the construct takes the sink only if nothing follows it in this block
                k = self._construct_end(s, i)
                tail = s[k:].strip()
                i = self.construct(s, i, depth, sink if tail == "" else None)
                if tail.startswith(";"):
                    i = s.index(";", i) + 1
                emitted = True
                continue
            if re.match(r"return\b", rest):
                end = self._find_top(s, i, ";")
                self.emit(depth, "return " + tr_expr(s[i + 6:end], self.cls))
                i = end + 1
                emitted = True
                continue
            if re.match(r"break\s*;", rest):
                self.emit(depth, "break")
                i = s.index(";", i) + 1
                emitted = True
                continue
            # expression statement or trailing expression
            try:
                end = self._find_top(s, i, ";")
                text, trailing = s[i:end], False
                i = end + 1
            except ValueError:
                text, trailing = s[i:], True
                i = n
            t = text.strip()
            if t.startswith(("assert!", "debug_assert")) or "check_validity" in t or t.startswith("panic!"):
                if t.startswith("panic!"):
                    self.emit(depth, "raise AssertionError('panic')")
                    emitted = True
                continue
            e = tr_expr(t, self.cls)
            e = re.sub(r"^(\w+) >>= (.*)$", r"\1 = \1 >> \2", e)
            e = re.sub(r"^(\w+) -= (.*)$", r"\1 = \1 - (\2)", e)
            e = re.sub(r"^(\w+) \+= (.*)$", r"\1 = \1 + (\2)", e)
            if trailing and sink:
                self.emit(depth, f"{sink} {e}")
            else:
                self.emit(depth, e)
            emitted = True
        if not emitted:
            self.emit(depth, "pass")

    @staticmethod
    def _find_top(s: str, i: int, ch: str) -> int:
        depth = 0
        while i < len(s):
            c = s[i]
            if c == ch and depth == 0:
                if ch == "=" and (s[i + 1] in "=>" or s[i - 1] in "=!<>+-"):
                    i += 1
                    continue
                return i
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            i += 1
        raise ValueError(f"no top-level {ch!r}")

    def _construct_end(self, s: str, i: int) -> int:
        """index just after an if / match / for construct starting at i (including else chains)"""
        o = self._find_top(s, i, "{")
        c = match_close(s, o)
        if re.match(r"if\b", s[i:]):
            m = re.match(r"\s*else\s*", s[c + 1:])
            if m:
                k = c + 1 + m.end()
                if s[k] == "{":
                    return match_close(s, k) + 1
                return self._construct_end(s, k)
        return c + 1

    def construct(self, s: str, i: int, depth: int, sink) -> int:
        o = self._find_top(s, i, "{")
        c = match_close(s, o)
        head = s[i:o].strip()
        if head.startswith("if"):
            self.emit(depth, f"if {tr_expr(head[2:], self.cls)}:")
            self.block(s[o + 1:c], depth + 1, sink)
            m = re.match(r"\s*else\s*", s[c + 1:])
            if not m:
                return c + 1
            k = c + 1 + m.end()
            if s[k] == "{":
                c2 = match_close(s, k)
                self.emit(depth, "else:")
                self.block(s[k + 1:c2], depth + 1, sink)
                return c2 + 1
            self.emit(depth, "else:")
            return self.construct(s, k, depth + 1, sink)
        if head.startswith("for"):
            m = re.match(r"for\s+(\w+)\s+in\s+(.+?)\.\.(=?)(.+)$", head, flags=re.S)
            var, lo, incl, hi = m.group(1), tr_expr(m.group(2), self.cls), m.group(3), tr_expr(m.group(4), self.cls)
            self.emit(depth, f"for {var} in (U(_i) for _i in range({lo}, ({hi}){' + 1' if incl else ''})):")
            self.block(s[o + 1:c], depth + 1, None)
            return c + 1
        assert head.startswith("match"), head
        subj = tr_expr(head[5:], self.cls)
        body, k, first = s[o + 1:c], 0, True
        while True:
            while k < len(body) and (body[k].isspace() or body[k] == ","):
                k += 1
            if k >= len(body):
                break
            arrow = body.index("=>", k)
            pat = body[k:arrow].strip()
            k = arrow + 2
            while body[k].isspace():
                k += 1
            variant = re.match(r"(?:Self|SecurityAssumption)::(\w+)$", pat)
            assert variant, pat
            self.emit(depth, f"{'if' if first else 'elif'} {subj}.kind == \"{variant.group(1)}\":")
            first = False
            if body[k] == "{":
                c2 = match_close(body, k)
                self.block(body[k + 1:c2], depth + 1, sink)
                k = c2 + 1
            else:
                try:
                    end = self._find_top(body, k, ",")
                except ValueError:
                    end = len(body)
                self.block(body[k:end], depth + 1, sink)
                k = end + 1
        self.emit(depth, "else:")
        self.emit(depth + 1, "raise AssertionError('unmatched')")
        return c + 1


def translate_impl(src: str, cls: str, header: str, wanted: set[str] | None) -> list[str]:
    """every `fn` of the impl block starting at `header` as methods of a Python class"""
    i = src.index(header)
    o = src.index("{", i + len(header) - 1) if not header.rstrip().endswith("{") else i + len(header.rstrip()) - 1
    c = match_close(src, o)
    body = src[o + 1:c]
    out = [f"class {cls}(_Struct):"]
    for m in re.finditer(r"\bfn\s+(\w+)\s*\(", body):
        name = m.group(1)
        if wanted is not None and name not in wanted:
            continue
        po = body.index("(", m.start())
        pc = match_close(body, po)
        params = []
        for p in split_top(body[po + 1:pc], ","):
            p = p.strip()
            if not p:
                continue
            params.append("self" if p in ("&self", "self", "&mut self") else p.split(":")[0].strip())
        bo = body.index("{", pc)
        bc = match_close(body, bo)
        em = Emitter(cls)
        if not params or params[0] != "self":
            em.emit(1, "@staticmethod")
        em.emit(1, f"def {name}({', '.join(params)}):")
        em.block(body[bo + 1:bc], 2, "return")
        out += em.lines + [""]
    return out


PRELUDE = '''
class _Struct:
    @classmethod
    def _make(cls, **kw):
        o = cls.__new__(cls)
        o.__dict__.update(kw)
        return o
'''


def translate(path: str = REF) -> str:
    src = strip_comments(open(path).read())
    lines = [PRELUDE]
    lines += translate_impl(src, "FoldingFactor", "impl FoldingFactor {", {"new", "constant", "at_round", "compute_number_of_rounds", "total_number"})
    lines += translate_impl(src, "RoundConfig", "pub struct RoundConfig<EF: Field> {", set())
    lines += ["    pass", ""]
    lines += translate_impl(src, "WhirConfig", "PF<EF>: TwoAdicField,\n{", {
        "compute_optimal_log_c_for_rate", "new", "rbr_soundness_fold_sumcheck", "folding_pow_bits",
        "rbr_soundness_queries_combination", "n_rounds", "rs_reduction_factor", "log_inv_rate_at", "merkle_tree_height",
        "n_vars_of_final_polynomial", "final_round_config", "starting_domain_size"})
    lines += translate_impl(src, "SecurityAssumption", "impl SecurityAssumption {", None)
    return "\n".join(lines)


def two_adic_generator(bits) -> int:
    """TWO_ADIC_GENERATORS[bits], read from koala-bear/src/koala_bear.rs:50-54 (canonical values); Montgomery form"""
    src = open(os.path.join(os.path.dirname(REF), "..", "..", "backend", "koala-bear", "src", "koala_bear.rs")).read()
    m = re.search(r"TWO_ADIC_GENERATORS: Self::ArrayLike = &KoalaBear::new_array\(\[(.*?)\]\)", src, flags=re.S)
    table = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
    assert len(table) == 25
    return (table[int(bits)] << 32) % P


def load(path: str = REF):
    code = translate(path)
    ns = {"F64": F64, "U": U, "as_f64": as_f64, "as_usize": as_usize, "EF_BITS": U((P ** 5).bit_length()),
          "TWO_ADICITY": U(24), "two_adic_generator": two_adic_generator}
    exec(compile(code, "<config.rs translated>", "exec"), ns)
    return ns, code


CASES = [(nv, rate) for rate in (1, 2, 3, 4) for nv in range(12, 29) if nv + rate - 7 <= 24]


def run_case(ns, nv: int, rate: int) -> dict:
    sa = ns["SecurityAssumption"]._make(kind="JohnsonBound")
    b = ns["_Struct"]._make(starting_log_inv_rate=U(rate), max_num_variables_to_send_coeffs=U(8),
                            rs_domain_initial_reduction_factor=U(5), folding_factor=ns["FoldingFactor"].new(U(7), U(5)),
                            soundness_type=sa, security_level=U(124), pow_bits=U(16))
    cfg = ns["WhirConfig"].new(b, U(nv))
    rounds = [{k: int(v) for k, v in r.__dict__.items()} for r in cfg.round_parameters]
    out = {"num_variables": nv, "starting_log_inv_rate": rate,
           "commitment_ood_samples": int(cfg.commitment_ood_samples),
           "starting_folding_pow_bits": int(cfg.starting_folding_pow_bits),
           "final_queries": int(cfg.final_queries), "final_query_pow_bits": int(cfg.final_query_pow_bits),
           "final_log_inv_rate": int(cfg.final_log_inv_rate), "final_sumcheck_rounds": int(cfg.final_sumcheck_rounds),
           "rounds": rounds}
    if rounds:
        out["final_round_config"] = {k: int(v) for k, v in cfg.final_round_config().__dict__.items()}
        out["merkle_tree_heights"] = [int(cfg.merkle_tree_height(U(r))) for r in range(len(rounds) + 1)]
    return out


def generate(path: str = REF) -> dict:
    ns, code = load(path)
    return {"note": "generated by tools/gen_whir_config_golden.py from the source text of crates/whir/src/config.rs "
                    "(default_whir_config parameters, lean_prover/src/lib.rs:22-49); folded_domain_gen in Montgomery form",
            "cases": [run_case(ns, nv, rate) for nv, rate in CASES],
            "translated_python": code.split("\n")}


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("needs /root/reference")
    d = generate()
    with open(OUT, "w") as f:
        json.dump(d, f, indent=0)
    print(f"wrote {OUT}: {len(d['cases'])} cases")

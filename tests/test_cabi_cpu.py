"""CPU tier: the C-ABI library loads, exports every declared symbol, and fails loudly without a GPU."""
import ctypes as C
import subprocess
import sys
import os

import numpy as np
import pytest

import leanmultisig_b200 as L
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    lib = L.lib()
    names = L.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback():
    lib = L.lib()
    if lib.lm_device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.lm_init(0, 20, C.byref(h))
    assert rc == -3 and b"no CPU path" in lib.lm_last_error()
    with pytest.raises(L.LmError):
        L.Context(0, 20)


def test_product_does_not_import_oracle():
    # the product package must never reach into oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "leanmultisig_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and '"oracle/' not in txt, f


def test_generated_poseidon_tables_are_current():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_poseidon1_consts.py"), "--check"]) == 0


@pytest.fixture(scope="module")
def hostcheck():
    """The kernel's arithmetic headers compiled for the host (tests/hostcheck): same code the GPU runs."""
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    so = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    deps = [src] + [os.path.join(ROOT, "leanmultisig_b200", "csrc", f) for f in ("kb.cuh", "poseidon1.cuh", "poseidon1_tables.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    for f in ("hc_kb_mul", "hc_kb_add", "hc_kb_sub", "hc_r2"):
        getattr(lib, f).restype = C.c_uint32
    return lib


def test_kernel_arithmetic_on_host_matches_oracle(hostcheck, rng):
    x = O.random_field(rng, (512, 16))
    x[0] = O.P - 1
    x[1] = 0
    y = x.copy()
    hostcheck.hc_poseidon1_permute(y.ctypes.data_as(O.u32p), C.c_uint64(len(y)))
    assert np.array_equal(y, O.poseidon1_permute(x, dense=True))
    y = x.copy()
    hostcheck.hc_poseidon1_compress8(y.ctypes.data_as(O.u32p), C.c_uint64(len(y)))
    assert np.array_equal(y[:, :8], O.poseidon1_compress(x)[:, :8])
    a, b = O.random_field(rng, (300, 5)), O.random_field(rng, (300, 5))
    out = np.empty_like(a)
    hostcheck.hc_ef_mul(a.ctypes.data_as(O.u32p), b.ctypes.data_as(O.u32p), out.ctypes.data_as(O.u32p), C.c_uint64(300))
    for i in range(300):
        assert np.array_equal(out[i], O.ef_mul(a[i], b[i]))
    for u, v in [(0, 0), (O.P - 1, O.P - 1), (1, O.P - 1), (12345, 678910)]:
        assert hostcheck.hc_kb_mul(u, v) == O.kb_mul(u, v)
        assert hostcheck.hc_kb_add(u, v) == (u + v) % O.P
        assert hostcheck.hc_kb_sub(u, v) == (u - v) % O.P
    assert hostcheck.hc_r2() == pow(2, 64, O.P)


def test_tensor_core_poseidon_formulation_model_matches_oracle(rng):
    """csrc/poseidon1_umma.cuh on the CPU: the B-matrix image the Merkle kernels load (u8 limbs of 4 C R^-1 2^(8i), G, MI | V and the
    triangle blocks pre-shifted mod p), the row layout, the no-carry recombination and the 8 / 8 / 4 block structure of the
    partial rounds, with every tcgen05.mma replaced by the integer dot products it stands for — against the reference's KAT
    (poseidon1_koalabear_16.rs:1066-1093) and the oracle permutation on random and extreme states."""
    from leanmultisig_b200._lib import lib

    kat_in = O.to_monty(np.arange(16))
    st = kat_in.copy()
    assert lib().lm_host_poseidon1_umma_model(st.ctypes.data_as(O.u32p)) == 0
    assert np.array_equal(st, O.poseidon1_permute(kat_in[None, :])[0])
    x = O.random_field(rng, (300, 16))
    x[0] = O.P - 1
    x[1] = 0
    x[2, ::2] = O.P - 1
    exp = O.poseidon1_permute(x)
    for i in range(len(x)):
        st = x[i].copy()
        assert lib().lm_host_poseidon1_umma_model(st.ctypes.data_as(O.u32p)) == 0
        assert np.array_equal(st, exp[i]), i


def test_tensor_core_poseidon_image_bounds():
    """The exactness argument of csrc/poseidon1_umma.cuh on the ACTUAL constants: every s32 accumulator column of every product
    stays below 2^24 for all-255 input bytes (so the u8 x u8 -> s32 MMAs never wrap and the byte recombination has room), and
    the carry-free low word `init + T0 + 2^8 T1` of the recombination stays below 2^32 with the largest additive constant.
    Columns that accumulate several products (D = G x' plus the triangle blocks) are checked on their totals."""
    from leanmultisig_b200._lib import lib

    n = int(lib().lm_host_poseidon1_umma_image(None, 0))
    img = np.zeros(n, dtype=np.uint8)
    assert int(lib().lm_host_poseidon1_umma_image(img.ctypes.data_as(C.c_void_p), n)) == n

    def col_sums(base, n_cols, k_bytes):  # K-major canonical layout: 8-column x 16-byte core matrices, LBO 128, SBO = k_bytes / 16 * 128
        kchunks = k_bytes // 16
        out = np.zeros(n_cols, dtype=np.int64)
        for c in range(n_cols):
            for kc in range(kchunks):
                off = base + (c // 8) * (kchunks * 128) + kc * 128 + (c % 8) * 16
                out[c] += int(img[off:off + 16].astype(np.int64).sum())
        return out

    B_MDS, B_G, B_MV = 0, 8192, 8192 + 6144
    B_T0 = B_MV + 10240
    B_T1 = B_T0 + 1536
    assert n == B_T1 + 512
    mds = 255 * col_sums(B_MDS, 64, 128)
    g = 255 * col_sums(B_G, 96, 64)
    g[32:80] += 255 * col_sums(B_T0, 48, 32)   # z_0..7 accumulate into D_8..19
    g[64:80] += 255 * col_sums(B_T1, 16, 32)   # z_8..15 into D_16..19
    mv = 255 * col_sums(B_MV, 64, 160)
    assert g[84:].max() == 0 and mv[60:].max() == 0                      # padding columns are empty
    for name, t, init in (("mds", mds, O.P - 1), ("g", g[:84], O.P - 1), ("mv", mv[:60], 0)):
        assert int(t.max()) < 1 << 24, name
        t = t.reshape(-1, 4)
        low = init + t[:, 0] + (t[:, 1] << 8)
        assert int(low.max()) < 1 << 32, (name, int(low.max()))
        # the reduced lane is < p + 2^16: the whole recombined value stays below 2^48
        full = init + t[:, 0] + (t[:, 1] << 8) + (t[:, 2] << 16) + (t[:, 3] << 24)
        assert int(full.max()) < 1 << 48, name
    # no output column is identically zero (every lane depends on its inputs)
    assert (mds > 0).all() and (g[:84] > 0).all() and (mv[:60] > 0).all()


@pytest.mark.parametrize("k,hi_vars,lo_vars", [(1, 2, 3), (3, 4, 5), (8, 3, 4), (12, 2, 6)])
def test_tensor_core_statement_weights_model_matches_oracle(rng, k, hi_vars, lo_vars):
    """csrc/sumcheck.cu weights_gemm_kernel on the CPU: the statement weights sum_k scalar_k eq(point_k, x) as an integer GEMM of u8
    limbs against pre-shifted row matrices (same image builder, row layout and recombination as the kernel, every MMA replaced by
    its dot products) equals the oracle's weights_add_eq (open.rs:518-584), on top of existing weights."""
    from leanmultisig_b200._lib import check, lib

    m = hi_vars + lo_vars
    pts, scs = O.random_field(rng, (k, m, 5)), O.random_field(rng, (k, 5))
    w = O.random_field(rng, (1 << m, 5))
    exp = w.copy()
    for i in range(k):
        O.weights_add_eq(exp, 0, pts[i], scs[i])
    hi = np.ascontiguousarray(np.stack([O.eq_table(pts[i][:hi_vars], scs[i]) for i in range(k)]))
    lo = np.ascontiguousarray(np.stack([O.eq_table(pts[i][hi_vars:]) for i in range(k)]))
    assert hi.shape == (k, 1 << hi_vars, 5) and lo.shape == (k, 1 << lo_vars, 5)
    check(lib().lm_host_eq_gemm_model(w.ctypes.data_as(O.u32p), hi.ctypes.data_as(O.u32p), lo.ctypes.data_as(O.u32p), k, hi_vars, lo_vars))
    assert np.array_equal(w, exp)


def test_native_prover_state_matches_the_python_transcript():
    """lm_fs (C++ ProverState, csrc/spine.cu) against the Python mirror and the oracle's challenger on one script of
    absorb / squeeze operations: identical samples, transcript and sponge state.  Host code only (no device)."""
    import numpy as np

    import leanmultisig_b200 as lm
    import oracle as O
    from oracle import whir as W

    rng = np.random.default_rng(5)
    script = [("scalars", O.random_field(rng, 13)), ("sample_vec", 3), ("poly", O.random_field(rng, (4, 5)), None),
              ("sample", None), ("poly", O.random_field(rng, (3, 5)), O.random_field(rng, 5)), ("sample", None),
              ("duplex", None), ("sample_in_range", 9, 11), ("observe", O.random_field(rng, 8)), ("sample_vec", 4),
              ("scalars", O.random_field(rng, 40)), ("sample", None)]
    outs = []
    for ps in (lm.ProverState(), lm.NativeProverState(), W.ProverState()):
        got = []
        for op in script:
            if op[0] == "scalars":
                ps.add_extension_scalars(op[1])
            elif op[0] == "observe":
                ps.observe_scalars(op[1])
            elif op[0] == "poly":
                ps.add_sumcheck_polynomial(op[1], op[2])
            elif op[0] == "duplex":
                ps.duplex()
            elif op[0] == "sample":
                got.append([int(x) for x in ps.sample()])
            elif op[0] == "sample_vec":
                got.append([[int(x) for x in v] for v in ps.sample_vec(op[1])])
            elif op[0] == "sample_in_range":
                got.append(ps.sample_in_range(op[1], op[2]))
        outs.append((got, list(ps.transcript)))
    assert outs[0] == outs[1] == outs[2]


def test_verify_openings_rejects_bad_arguments_without_a_device():
    """lm_verify_openings validates its arguments before touching the device (include/leanmultisig_b200.h): null handles, a row
    width that hash_slice cannot take (sponge.rs:7-25 needs a multiple of 8, at least 16) and a fold shape that does not match."""
    lib = L.lib()
    u8p = C.POINTER(C.c_uint8)
    root = np.zeros(8, dtype=np.uint32)
    idx = np.zeros(2, dtype=np.uint64)
    rows = np.zeros((2, 16), dtype=np.uint32)
    paths = np.zeros((2, 3, 8), dtype=np.uint32)
    ok = np.zeros(2, dtype=np.uint8)
    p32 = lambda a: a.ctypes.data_as(O.u32p)  # noqa: E731
    args = lambda ctx, width, fold_vars, evals: (  # noqa: E731
        ctx, p32(root), 3, idx.ctypes.data_as(C.POINTER(C.c_uint64)), 2, p32(rows), width, 1, p32(paths), None, fold_vars,
        ok.ctypes.data_as(u8p), evals)
    assert lib.lm_verify_openings(*args(None, 16, 0, None)) != 0 and b"null" in lib.lm_last_error()
    fake = C.c_void_p(1)  # never dereferenced: the checks below come first
    assert lib.lm_verify_openings(*args(fake, 12, 0, None)) != 0 and b"width" in lib.lm_last_error()
    assert lib.lm_verify_openings(*args(fake, 20, 0, None)) != 0 and b"width" in lib.lm_last_error()
    ev = np.zeros((2, 5), dtype=np.uint32)
    assert lib.lm_verify_openings(*args(fake, 16, 3, p32(ev))) != 0 and b"leaf" in lib.lm_last_error()


def test_every_declared_symbol_is_accounted_for_in_the_integration_notes():
    """INTEGRATION.md either shows the reference-side binding of an entry point or lists it among those a Rust integration does
    not bind (section 2.7): nothing in the header is undocumented at the boundary."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert [s for s in L.declared_symbols() if s not in txt] == []


def test_public_header_is_plain_c():
    """the boundary is a C ABI: include/leanmultisig_b200.h must compile as C11 (what cgo / bindgen / ctypes-style FFI consume)
    and as C++17, with nothing but <stdint.h> / <stddef.h> types in the signatures"""
    hdr = os.path.join(ROOT, "include", "leanmultisig_b200.h")
    assert subprocess.call(["gcc", "-fsyntax-only", "-x", "c", "-std=c11", "-Wall", "-Werror", hdr]) == 0
    assert subprocess.call(["g++", "-fsyntax-only", "-x", "c++", "-std=c++17", hdr]) == 0
    import re

    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)  # declarations only: comments may mention the caller's plumbing
    assert "torch" not in code and "at::" not in code and "std::" not in code

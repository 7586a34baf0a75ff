"""ORACLE — TEST INFRASTRUCTURE ONLY.

Logup of the lean_vm tables restated from the reference:
  table assembly            crates/sub_protocols/src/logup.rs:27-211   (prove_generic_logup, natural row order: the
                            reference's chunk-bit-reversed storage is a SIMD layout internal to its GKR)
  verify_generic_logup      crates/sub_protocols/src/logup.rs:323-493
  verify_gkr_quotient       crates/sub_protocols/src/quotient_gkr/mod.rs:147-190
  helpers                   crates/utils/src/multilinear.rs:67-84, crates/backend/poly/src/mle/mle_custom.rs:4-19
Table metadata: crates/lean_vm/src/tables/{execution/mod.rs:27-58, extension_op/mod.rs:91-124, poseidon_16/mod.rs:140-182}.
Values are canonical Python ints / EF 5-tuples internally, Montgomery uint32 at the array interfaces.
"""
from __future__ import annotations

import numpy as np

import oracle as O
from oracle import whir as W

P = O.P
MEMORY_DOMAINSEP, PRECOMPILE_DOMAINSEP, BYTECODE_DOMAINSEP = 0, 1, 2
N_RUNTIME_COLUMNS, N_INSTRUCTION_COLUMNS, COL_PC = 8, 12, 0
N_VARS_TO_SEND_GKR_COEFFS = 5

# name -> (enum order, pull?, selector column, bus data columns, lookups [(index column, value columns)])
TABLES = {
    "execution": (0, False, 20, (19, 21, 22, 23), ((2, (5,)), (3, (6,)), (4, (7,)))),
    "extension_op": (1, True, 29, (30, 6, 7, 13), ((6, tuple(range(14, 19))), (7, tuple(range(19, 24))), (13, tuple(range(24, 29))))),
    "poseidon16": (2, True, 0, (110, 109, 1, 2),
                   ((6, tuple(range(9, 13))), (7, tuple(range(13, 17))), (1, tuple(range(17, 25))), (2, tuple(range(93, 109))))),
}


def sort_tables_by_height(log_heights: dict) -> list:
    return sorted(log_heights.items(), key=lambda kv: (-kv[1], TABLES[kv[0]][0]))


def compute_total_active_len(log_memory, log_bytecode, tables_sorted) -> int:
    max_h = 1 << tables_sorted[0][1]
    log_cycles = dict(tables_sorted)["execution"]
    tot = (1 << log_memory) + max(1 << log_bytecode, max_h) + (1 << log_cycles)
    for name, h in tables_sorted:
        tot += (sum(len(v) for _, v in TABLES[name][4]) + 1) << h
    return tot


def _fp_rows(alphas_c, domainsep, data_cols):
    """rows of finger_print: alphas.last * domainsep + sum_i alphas[i] * data_i   (canonical uint64 [n, 5])"""
    n = data_cols[0].shape[0]
    acc = np.zeros((n, 5), dtype=np.uint64)
    p = np.uint64(P)
    for a, d in zip(alphas_c, data_cols):
        for k in range(5):
            acc[:, k] = (acc[:, k] + d * np.uint64(a[k]) % p) % p
    for k in range(5):
        acc[:, k] = (acc[:, k] + np.uint64(alphas_c[-1][k] * domainsep % P)) % p
    return acc


def build_table(c, alphas_eq_poly, memory, memory_acc, bytecode_multilinear, bytecode_acc, traces):
    """-> (numerators [total] Montgomery, denominators [total, 5] Montgomery).  traces: {name: (columns list, log_n_rows)}"""
    p = np.uint64(P)
    cc = np.array(W.fm(c), dtype=np.uint64)
    al = [W.fm(a) for a in np.asarray(alphas_eq_poly).reshape(-1, 5)]
    can = lambda a: O.from_monty(a).astype(np.uint64)
    nums, dens = [], []

    def den(sign, domainsep, cols):
        fp = _fp_rows(al, domainsep, cols)
        return (cc[None, :] + fp) % p if sign > 0 else (cc[None, :] + (p - fp)) % p

    log_memory = memory.size.bit_length() - 1
    stride = 1 << (N_INSTRUCTION_COLUMNS - 1).bit_length()
    log_bytecode = (bytecode_multilinear.size // stride).bit_length() - 1
    sorted_t = sort_tables_by_height({k: v[1] for k, v in traces.items()})
    max_h = 1 << sorted_t[0][1]
    nums.append((p - can(memory_acc)) % p)
    dens.append(den(-1, MEMORY_DOMAINSEP, [can(memory), np.arange(memory.size, dtype=np.uint64)]))
    bc = can(bytecode_multilinear).reshape(-1, stride)
    nums.append((p - can(bytecode_acc)) % p)
    dens.append(den(-1, BYTECODE_DOMAINSEP, [bc[:, k] for k in range(N_INSTRUCTION_COLUMNS)] + [np.arange(1 << log_bytecode, dtype=np.uint64)]))
    if (1 << log_bytecode) < max_h:
        pad = max_h - (1 << log_bytecode)
        nums.append(np.zeros(pad, dtype=np.uint64))
        one = np.zeros((pad, 5), dtype=np.uint64)
        one[:, 0] = 1
        dens.append(one)
    for name, h in sorted_t:
        cols = traces[name][0]
        n = 1 << h
        _, pull, selector, bus_data, lookups = TABLES[name]
        if name == "execution":
            nums.append(np.ones(n, dtype=np.uint64))
            dens.append(den(-1, BYTECODE_DOMAINSEP, [can(cols[N_RUNTIME_COLUMNS + k]) for k in range(N_INSTRUCTION_COLUMNS)] + [can(cols[COL_PC])]))
        sel = can(cols[selector])
        nums.append((p - sel) % p if pull else sel)
        dens.append(den(+1, PRECOMPILE_DOMAINSEP, [can(cols[k]) for k in bus_data]))
        for index, values in lookups:
            for i, v in enumerate(values):
                nums.append(np.ones(n, dtype=np.uint64))
                dens.append(den(-1, MEMORY_DOMAINSEP, [can(cols[v]), (can(cols[index]) + np.uint64(i)) % p]))
    nums, dens = np.concatenate(nums), np.concatenate(dens)
    assert nums.size == compute_total_active_len(log_memory, log_bytecode, sorted_t)
    return O.to_monty(nums), O.to_monty(dens)


# ------------------------------------------------------------------------------------------ verifier
def mle_of_01234567_etc(point):
    if not point:
        return W.ZERO
    e = mle_of_01234567_etc(point[1:])
    x = point[0]
    return W.add(W.mul(W.sub(W.ONE, x), e), W.mul(x, W.add(e, ((1 << (len(point) - 1)) % P, 0, 0, 0, 0))))


def mle_of_zeros_then_ones(n_zeros, point):
    n_values = 1 << len(point)
    assert n_zeros <= n_values
    if n_zeros == 0:
        return W.ONE
    if n_zeros == n_values:
        return W.ZERO
    half = n_values // 2
    if n_zeros < half:
        return W.add(W.mul(W.sub(W.ONE, point[0]), mle_of_zeros_then_ones(n_zeros, point[1:])), point[0])
    return W.mul(point[0], mle_of_zeros_then_ones(n_zeros - half, point[1:]))


def finger_print(domainsep, data, alphas_c):
    assert len(alphas_c) > len(data)
    acc = W.scal(alphas_c[-1], domainsep)
    for a, d in zip(alphas_c, data):
        acc = W.add(acc, W.mul(a, d))
    return acc


def verify_gkr_quotient(vs: W.VerifierState, n_vars: int):
    """mod.rs:147-190 -> (quotient, point, claim_num, claim_den)"""
    assert n_vars > N_VARS_TO_SEND_GKR_COEFFS
    send = 1 << N_VARS_TO_SEND_GKR_COEFFS
    last_nums = [W.fm(x) for x in vs.next_extension_scalars_vec(send)]
    last_dens = [W.fm(x) for x in vs.next_extension_scalars_vec(send)]
    quotient = W.ZERO
    for a, b in zip(last_nums, last_dens):
        quotient = W.add(quotient, W.mul(a, W.inv(b)))
    point = [W.fm(x) for x in vs.sample_vec(N_VARS_TO_SEND_GKR_COEFFS)]
    cn, cd = W.mle_eval_small(last_nums, point), W.mle_eval_small(last_dens, point)
    for k in range(N_VARS_TO_SEND_GKR_COEFFS, n_vars):
        vs.duplex()
        alpha = W.fm(vs.sample())
        expected = W.add(cn, W.mul(alpha, cd))
        chals, value = W.sumcheck_verify(vs, k, 3, expected, point[::-1])
        q = chals[::-1]
        nl, nr, dl, dr = (W.fm(x) for x in vs.next_extension_scalars_vec(4))
        ce = W.add(W.mul(alpha, W.mul(dl, dr)), W.add(W.mul(nl, dr), W.mul(nr, dl)))
        if value != W.mul(W.eq_outside(point, q), ce):
            raise W.ProofError("InvalidProof: gkr layer %d" % k)
        beta = W.fm(vs.sample())
        omb = W.sub(W.ONE, beta)
        cn, cd = W.add(W.mul(omb, nl), W.mul(beta, nr)), W.add(W.mul(omb, dl), W.mul(beta, dr))
        point = q + [beta]
    return quotient, point, cn, cd


def verify_generic_logup(vs, c, alphas, alphas_eq_poly, log_memory, bytecode_multilinear, table_log_n_rows: dict):
    """logup.rs:323-493 -> dict of the statements the verifier derives"""
    c = W.fm(c)
    alphas_c = [W.fm(a) for a in np.asarray(alphas).reshape(-1, 5)]
    al = [W.fm(a) for a in np.asarray(alphas_eq_poly).reshape(-1, 5)]
    sorted_t = sort_tables_by_height(table_log_n_rows)
    stride = 1 << (N_INSTRUCTION_COLUMNS - 1).bit_length()
    log_stride = stride.bit_length() - 1
    log_bytecode = (bytecode_multilinear.size // stride).bit_length() - 1
    total_n_vars = (compute_total_active_len(log_memory, log_bytecode, sorted_t) - 1).bit_length()
    quotient, point, num_value, den_value = verify_gkr_quotient(vs, total_n_vars)
    if quotient != W.ZERO:
        raise W.ProofError("InvalidProof: logup sum")
    got_num, got_den = W.ZERO, W.ZERO

    def from_end(k):
        return point[len(point) - k:]

    def pref_at(offset, log_height):
        n_missing = total_n_vars - log_height
        bits = [(((offset >> log_height) >> i) & 1, 0, 0, 0, 0) for i in range(n_missing - 1, -1, -1)]
        return W.eq_outside(bits, point[:n_missing])

    nxt = lambda: W.fm(vs.next_extension_scalars_vec(1)[0])
    out = {}
    mem_pt = from_end(log_memory)
    pref = pref_at(0, log_memory)
    value_memory_acc = nxt()
    got_num = W.sub(got_num, W.mul(pref, value_memory_acc))
    value_memory = nxt()
    got_den = W.add(got_den, W.mul(pref, W.sub(c, finger_print(MEMORY_DOMAINSEP, [value_memory, mle_of_01234567_etc(mem_pt)], al))))
    offset = 1 << log_memory
    log_bc_padded = max(log_bytecode, sorted_t[0][1])
    bc_pt = from_end(log_bytecode)
    pref, pref_padded = pref_at(offset, log_bytecode), pref_at(offset, log_bc_padded)
    value_bytecode_acc = nxt()
    got_num = W.sub(got_num, W.mul(pref, value_bytecode_acc))
    bc_full_pt = bc_pt + alphas_c[len(alphas_c) - log_stride:]
    bytecode_value = W.fm(O.mle_eval(bytecode_multilinear, W._pts(bc_full_pt)))
    corr = bytecode_value
    for x in alphas_c[: len(alphas_c) - log_stride]:
        corr = W.mul(corr, W.sub(W.ONE, x))
    inner = W.add(W.add(corr, W.mul(mle_of_01234567_etc(bc_pt), al[N_INSTRUCTION_COLUMNS])), W.scal(al[-1], BYTECODE_DOMAINSEP))
    got_den = W.add(got_den, W.mul(pref, W.sub(c, inner)))
    got_den = W.add(got_den, W.mul(pref_padded, mle_of_zeros_then_ones(1 << log_bytecode, from_end(log_bc_padded))))
    offset += 1 << log_bc_padded
    out.update(value_memory=value_memory, value_memory_acc=value_memory_acc, value_bytecode_acc=value_bytecode_acc,
               gkr_point=point, columns_values={}, bus_numerators_values={}, bus_denominators_values={})
    for name, h in sorted_t:
        values = {}
        _, _, _, _, lookups = TABLES[name]
        if name == "execution":
            eval_pc = nxt()
            values[COL_PC] = eval_pc
            instr = [W.fm(x) for x in vs.next_extension_scalars_vec(N_INSTRUCTION_COLUMNS)]
            for k, v in enumerate(instr):
                values[N_RUNTIME_COLUMNS + k] = v
            pref = pref_at(offset, h)
            got_num = W.add(got_num, pref)
            got_den = W.add(got_den, W.mul(pref, W.sub(c, finger_print(BYTECODE_DOMAINSEP, instr + [eval_pc], al))))
            offset += 1 << h
        sel = nxt()
        pref = pref_at(offset, h)
        got_num = W.add(got_num, W.mul(pref, sel))
        data = nxt()
        got_den = W.add(got_den, W.mul(pref, data))
        out["bus_numerators_values"][name], out["bus_denominators_values"][name] = sel, data
        offset += 1 << h
        for index, vals in lookups:
            index_eval = nxt()
            values[index] = index_eval
            for i, vcol in enumerate(vals):
                v = nxt()
                values[vcol] = v
                pref = pref_at(offset, h)
                got_num = W.add(got_num, pref)
                got_den = W.add(got_den, W.mul(pref, W.sub(c, finger_print(MEMORY_DOMAINSEP, [v, W.add(index_eval, (i, 0, 0, 0, 0))], al))))
                offset += 1 << h
        out["columns_values"][name] = values
    got_den = W.add(got_den, mle_of_zeros_then_ones(offset, point))
    if got_num != num_value:
        raise W.ProofError("InvalidProof: logup numerators")
    if got_den != den_value:
        raise W.ProofError("InvalidProof: logup denominators")
    return out


# ------------------------------------------------------------------------------------------ CPU GKR prover
def prove_gkr_quotient_cpu(ps: W.ProverState, nums, dens):
    """prove_gkr_quotient / prove_gkr_layer (quotient_gkr/mod.rs:31-141) on the oracle kernels, natural order.
    nums: base-field active prefix, dens: [active, 5].  -> (quotient, point, claim_num, claim_den) canonical"""
    active = nums.size
    n_vars = (active - 1).bit_length()
    n = 1 << n_vars
    pn = np.zeros(n, dtype=np.uint32)
    pn[:active] = nums
    pd = np.zeros((n, 5), dtype=np.uint32)
    pd[:, 0] = int(O.to_monty(1))
    pd[:active] = dens
    layers = [(pn, pd)]
    for _ in range(n_vars - N_VARS_TO_SEND_GKR_COEFFS):
        layers.append(O.gkr_layer_up(*layers[-1]))
    tn, td = layers.pop()
    ps.add_extension_scalars(tn.reshape(-1))
    ps.add_extension_scalars(td.reshape(-1))
    top_n, top_d = [W.fm(v) for v in tn], [W.fm(v) for v in td]
    quotient = W.ZERO
    for a, b in zip(top_n, top_d):
        quotient = W.add(quotient, W.mul(a, W.inv(b)))
    point = [W.fm(x) for x in ps.sample_vec(N_VARS_TO_SEND_GKR_COEFFS)]
    cn, cd = W.mle_eval_small(top_n, point), W.mle_eval_small(top_d, point)
    for lay_n, lay_d in reversed(layers):
        ps.duplex()
        alpha_m = ps.sample()
        alpha = W.fm(alpha_m)
        s, mmf = W.add(cn, W.mul(alpha, cd)), W.ONE
        emb = (lambda x: O.embed(x)) if lay_n.ndim == 1 else (lambda x: x)
        cols = [emb(lay_n[0::2]), emb(lay_n[1::2]), np.ascontiguousarray(lay_d[0::2]), np.ascontiguousarray(lay_d[1::2])]
        remaining, q = list(point), []
        for _ in range(len(point)):
            c0, c2 = O.gkr_round(*cols, W._pts(remaining[:-1]), alpha_m)
            a = remaining[-1]
            c0c, c2c = W.mul(W.fm(c0), mmf), W.mul(W.fm(c2), mmf)
            h1 = W.mul(W.sub(s, W.mul(W.sub(W.ONE, a), c0c)), W.inv(a))
            bare = [c0c, W.sub(W.sub(h1, c0c), c2c), c2c]
            ps.add_sumcheck_polynomial(np.stack([W.tm(x) for x in bare]), W.tm(a))
            r_m = ps.sample()
            r = W.fm(r_m)
            eq_eval = W.add(W.mul(W.sub(W.ONE, a), W.sub(W.ONE, r)), W.mul(a, r))
            s = W.mul(eq_eval, W.peval(bare, r))
            mmf = W.mul(mmf, eq_eval)
            cols = [O.fold_lsb(cc, r_m) for cc in cols]
            q.append(r)
            remaining.pop()
        q.reverse()
        inner = np.stack([cc[0] for cc in cols])
        ps.add_extension_scalars(inner.reshape(-1))
        beta = W.fm(ps.sample())
        nl, nr, dl, dr = (W.fm(v) for v in inner)
        omb = W.sub(W.ONE, beta)
        cn, cd = W.add(W.mul(omb, nl), W.mul(beta, nr)), W.add(W.mul(omb, dl), W.mul(beta, dr))
        point = q + [beta]
    return quotient, point, cn, cd

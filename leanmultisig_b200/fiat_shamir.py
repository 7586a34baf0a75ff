"""Prover-side Fiat-Shamir transcript, mirroring the reference's `FSProver` trait and `ProverState`
(crates/backend/fiat-shamir/src/traits.rs, prover.rs:28-178, challenger.rs:8-76, utils.rs).

In a drop-in integration this object stays the reference's Rust `ProverState`; the WHIR / sumcheck drivers of this
package only need the trait surface below.  Two pieces run in the CUDA library:
  * `pow_grinding`'s witness search (data parallel, ~2^bits permutations)  -> lm_pow_grind (device)
  * the duplex permutation itself (sequential, one per 8 words)            -> lm_host_poseidon1_permute (host code
    compiled from the same arithmetic header as the kernels)
Everything is Montgomery-form uint32, exactly the reference's in-memory representation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import field as F
from ._lib import check, lib, u32p

RATE, WIDTH, CAPACITY = 8, 16, 8


def _permute(state: np.ndarray) -> None:
    check(lib().lm_host_poseidon1_permute(state.ctypes.data_as(u32p)))


class Challenger:
    """challenger.rs:8-76: observe overwrites the rate half and permutes; sample reads the rate half once."""

    def __init__(self):
        self.state = np.zeros(WIDTH, dtype=np.uint32)
        self.rate_fresh = False

    def observe(self, value) -> None:
        self.state[CAPACITY:] = value
        _permute(self.state)
        self.rate_fresh = True

    def observe_many(self, scalars) -> None:
        s = np.ascontiguousarray(scalars, dtype=np.uint32).reshape(-1)
        for i in range(0, s.size, RATE):
            chunk = s[i:i + RATE]
            if chunk.size < RATE:
                chunk = np.concatenate([chunk, np.zeros(RATE - chunk.size, dtype=np.uint32)])
            self.observe(chunk)

    def duplex(self) -> None:
        self.observe(np.zeros(RATE, dtype=np.uint32))

    def sample(self) -> np.ndarray:
        if not self.rate_fresh:
            raise RuntimeError("stale rate. insert a duplex() before.")
        self.rate_fresh = False
        return self.state[CAPACITY:].copy()

    def sample_many(self, n: int) -> np.ndarray:
        out = []
        for i in range(n):
            if i:
                self.duplex()
            out.append(self.sample())
        return np.concatenate(out) if out else np.zeros(0, dtype=np.uint32)

    def sample_in_range(self, bits: int, n_samples: int) -> list[int]:
        fes = self.sample_many(-(-n_samples // RATE))[:n_samples]
        canon = (fes.astype(np.uint64) * np.uint64(F._RINV)) % np.uint64(F.P)
        return [int(x) & ((1 << bits) - 1) for x in canon]


class ProverState:
    """prover.rs:28-178.  `ctx` (a whir.Context) is only needed when pow_grinding is called with bits > 0."""

    def __init__(self, ctx=None):
        self.ctx = ctx
        self.challenger = Challenger()
        self.transcript: list[int] = []
        self.merkle_paths: list[list] = []

    # -- absorb ------------------------------------------------------------------------------------------
    def add_base_scalars(self, scalars) -> None:
        s = np.ascontiguousarray(scalars, dtype=np.uint32).reshape(-1)
        self.challenger.observe_many(s)
        self.transcript.extend(int(x) for x in s)

    def add_extension_scalars(self, scalars) -> None:
        self.add_base_scalars(scalars)

    def observe_scalars(self, scalars) -> None:
        self.challenger.observe_many(scalars)

    def duplex(self) -> None:
        self.challenger.duplex()

    def add_sumcheck_polynomial(self, coeffs, eq_alpha=None) -> None:
        """prover.rs:105-128: everything is absorbed, the constant coefficient is not sent"""
        c = np.ascontiguousarray(coeffs, dtype=np.uint32).reshape(-1, 5)
        if eq_alpha is None:
            self.challenger.observe_many(c.reshape(-1))
        else:
            from .air import expand_bare_to_full

            full = expand_bare_to_full(c, eq_alpha)
            self.challenger.observe_many(np.concatenate([F.to_monty(x) for x in full]))
        self.transcript.extend(int(x) for x in c[1:].reshape(-1))

    def hint_merkle_paths(self, paths) -> None:
        self.merkle_paths.append(list(paths))

    def into_proof(self):
        """ProverState::into_proof (fiat-shamir/src/prover.rs:48-53): transcript + the query batches pruned as
        hint_merkle_paths_base does (prover.rs:116-118, merkle_pruning.rs:18-84) -> wire.Proof (postcard / lz4 ready)"""
        from .wire import Proof

        return Proof.from_prover_state(self)

    # -- squeeze -----------------------------------------------------------------------------------------
    def sample_vec(self, n: int) -> list[np.ndarray]:
        fes = self.challenger.sample_many(-(-(n * 5) // RATE))[: n * 5]
        return [fes[5 * i:5 * i + 5].copy() for i in range(n)]

    def sample(self) -> np.ndarray:
        return self.sample_vec(1)[0]

    def sample_in_range(self, bits: int, n_samples: int) -> list[int]:
        return self.challenger.sample_in_range(bits, n_samples)

    # -- proof of work -------------------------------------------------------------------------------------
    def pow_grinding(self, bits: int) -> None:
        if bits == 0:
            return
        if self.ctx is None:
            raise RuntimeError("pow_grinding needs a device context (ProverState(ctx))")
        w = C.c_uint64()
        st = np.ascontiguousarray(self.challenger.state)
        check(lib().lm_pow_grind(self.ctx.handle, st.ctypes.data_as(u32p), bits, C.byref(w)))
        wm = np.array([w.value * F._R % F.P], dtype=np.uint32)
        self.challenger.observe_many(wm)
        lane = int(self.challenger.state[CAPACITY]) * F._RINV % F.P
        if lane & ((1 << bits) - 1):
            raise RuntimeError("device PoW witness rejected by the host sponge")
        self.transcript.append(int(wm[0]))


class NativeProverState:
    """The same trait surface on the C++ transcript of the library (lm_fs, csrc/spine.cu).  Use it with the native
    drivers (`GkrQuotientProver.prove_native`, `prove_batched_air_sumcheck_native`), which run a whole protocol without
    returning to Python between rounds; every method of `ProverState` is available, so the WHIR prover and the Logup
    host code work on it unchanged."""

    def __init__(self, ctx=None):
        self.ctx = ctx
        h = C.c_void_p()
        check(lib().lm_fs_new(ctx.handle if ctx is not None else None, C.byref(h)))
        self.handle = h
        self.merkle_paths: list[list] = []

    def _words(self, a):
        s = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1)
        return s, s.ctypes.data_as(u32p)

    def add_base_scalars(self, scalars) -> None:
        s, p = self._words(scalars)
        check(lib().lm_fs_add_scalars(self.handle, p, s.size))

    add_extension_scalars = add_base_scalars

    def observe_scalars(self, scalars) -> None:
        s, p = self._words(scalars)
        check(lib().lm_fs_observe(self.handle, p, s.size))

    def duplex(self) -> None:
        check(lib().lm_fs_duplex(self.handle))

    def add_sumcheck_polynomial(self, coeffs, eq_alpha=None) -> None:
        c, p = self._words(coeffs)
        if eq_alpha is None:
            check(lib().lm_fs_add_sumcheck_polynomial(self.handle, p, c.size // 5, None))
        else:
            a, pa = self._words(eq_alpha)
            check(lib().lm_fs_add_sumcheck_polynomial(self.handle, p, c.size // 5, pa))

    def hint_merkle_paths(self, paths) -> None:
        self.merkle_paths.append(list(paths))

    def into_proof(self):
        """ProverState::into_proof (fiat-shamir/src/prover.rs:48-53): transcript + the query batches pruned as
        hint_merkle_paths_base does (prover.rs:116-118, merkle_pruning.rs:18-84) -> wire.Proof (postcard / lz4 ready)"""
        from .wire import Proof

        return Proof.from_prover_state(self)

    def sample_vec(self, n: int) -> list[np.ndarray]:
        out = np.empty((max(n, 1), 5), dtype=np.uint32)
        check(lib().lm_fs_sample(self.handle, n, out.ctypes.data_as(u32p)))
        return [out[i].copy() for i in range(n)]

    def sample(self) -> np.ndarray:
        return self.sample_vec(1)[0]

    def sample_in_range(self, bits: int, n_samples: int) -> list[int]:
        out = np.empty(max(n_samples, 1), dtype=np.uint64)
        check(lib().lm_fs_sample_in_range(self.handle, bits, n_samples, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return [int(x) for x in out[:n_samples]]

    def pow_grinding(self, bits: int) -> None:
        check(lib().lm_fs_pow_grinding(self.handle, bits))

    @property
    def transcript(self) -> list[int]:
        n = C.c_uint64()
        check(lib().lm_fs_transcript_len(self.handle, C.byref(n)))
        out = np.empty(max(n.value, 1), dtype=np.uint32)
        check(lib().lm_fs_transcript(self.handle, out.ctypes.data_as(u32p)))
        return [int(x) for x in out[: n.value]]

    @property
    def state(self) -> np.ndarray:
        st = np.empty(16, dtype=np.uint32)
        check(lib().lm_fs_state(self.handle, st.ctypes.data_as(u32p), None))
        return st

    def state_and_freshness(self):
        """(sponge state, rate_fresh): what `set_state` takes to resume this transcript elsewhere"""
        st, fresh = np.empty(16, dtype=np.uint32), C.c_int()
        check(lib().lm_fs_state(self.handle, st.ctypes.data_as(u32p), C.byref(fresh)))
        return st, bool(fresh.value)

    def set_state(self, state, rate_fresh: bool) -> None:
        s, p = self._words(state)
        assert s.size == 16
        check(lib().lm_fs_set_state(self.handle, p, 1 if rate_fresh else 0))

    def free(self):
        if self.handle:
            check(lib().lm_fs_free(self.handle))
            self.handle = None

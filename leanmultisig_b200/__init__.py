"""leanmultisig_b200 — B200-native proving hot path for leanEthereum/leanMultisig.

The product is the CUDA library behind ``include/leanmultisig_b200.h`` (``csrc/``).  This package is the thin
Python host side used by the tests and the benchmark: it mirrors the reference's call sites
(``WhirConfig::commit``, ``MerkleData::open``, ``MleRef::evaluate`` ...) one to one on top of the C ABI and
holds no arithmetic of its own.
"""
from ._lib import LIB_PATH, LmError, build, declared_symbols, lib  # noqa: F401
from .air import (AirSumcheckSession, fill_trace_poseidon_16, prove_batched_air_sumcheck,  # noqa: F401
                  prove_batched_air_sumcheck_native)
from .fiat_shamir import NativeProverState, ProverState  # noqa: F401
from .logup import GkrQuotientProver, finger_print  # noqa: F401
from .whir import Context, DeviceBuffer, ProductSumcheck, SparseStatement, Tree, WhirProver, Witness  # noqa: F401
from .whir_config import WhirConfig  # noqa: F401
from . import stacked_pcs  # noqa: F401,E402
from . import verify  # noqa: F401,E402

__all__ = ["Context", "DeviceBuffer", "ProductSumcheck", "Tree", "LmError", "build", "lib", "declared_symbols", "LIB_PATH"]

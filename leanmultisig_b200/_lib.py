"""ctypes loader for the C-ABI shared library (include/leanmultisig_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc (csrc/Makefile), and if that is
impossible, or no CUDA device is visible when a context is created, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libleanmultisig_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "leanmultisig_b200.h")

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class LmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"leanmultisig_b200 error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile every CUDA source for sm_100a into lib/libleanmultisig_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function the public header declares (used by the CPU-side export test)."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing and could not be built; the CUDA extension is required")
        L = C.CDLL(LIB_PATH)
        L.lm_last_error.restype = C.c_char_p
        L.lm_kernel_launches.restype = C.c_uint64
        L.lm_kernel_launches.argtypes = []
        vp, sz, u32, u64, i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
        sig = {
            "lm_device_count": [],
            "lm_init": [i, u32, C.POINTER(vp)],
            "lm_destroy": [vp],
            "lm_set_stream": [vp, vp],
            "lm_sync": [vp],
            "lm_host_register": [vp, sz],
            "lm_host_unregister": [vp],
            "lm_commit": [vp, vp, u32, u32, u64, u32, u32, C.POINTER(vp), u32p],
            "lm_commit_dev": [vp, vp, u32, u32, u64, u32, u32, i, C.POINTER(vp), u32p],
            "lm_commit_stacked": [vp, vp, u32, u32, u32, u32, C.POINTER(vp), u32p],
            "lm_access_counts": [vp, C.POINTER(vp), u64p, u32p, u32, u64, u32p],
            "lm_open": [vp, u64p, u32, u32p, u32p],
            "lm_verify_openings": [vp, u32p, u32, u64p, u32, u32p, u32, u32, u32p, u32p, u32, C.POINTER(C.c_uint8), u32p],
            "lm_tree_shape": [vp, u64p, u32p, u32p, u32p],
            "lm_tree_eval": [vp, u32p, u32p],
            "lm_tree_read_codeword": [vp, u32p],
            "lm_tree_read_layers": [vp, u32p],
            "lm_tree_free": [vp],
            "lm_mle_eval": [vp, vp, u32, u32, u64, u32p, u32p],
            "lm_sc_new_from_tree": [vp, C.POINTER(vp)],
            "lm_sc_new": [vp, vp, u32, u32, u64, C.POINTER(vp)],
            "lm_sc_add_eq": [vp, u64, u32p, u32, u32p],
            "lm_sc_add_next": [vp, u64, u32p, u32, u32p],
            "lm_sc_add_base_eq": [vp, u32p, u32, u32p],
            "lm_sc_round": [vp, u32p, u32p],
            "lm_sc_fold": [vp, u32p],
            "lm_sc_fold_round": [vp, u32p, u32p, u32p],
            "lm_sc_num_vars": [vp, u32p, u32p],
            "lm_sc_read": [vp, u32p, u32p],
            "lm_sc_eval_poly": [vp, u32p, u32p],
            "lm_sc_commit_poly": [vp, u32, u32, C.POINTER(vp), u32p],
            "lm_sc_export_dev": [vp, vp, vp],
            "lm_sc_new_from_dev": [vp, vp, vp, u32, C.POINTER(vp)],
            "lm_sc_free": [vp],
            "lm_air_new": [vp, u32, C.POINTER(vp), u32, u32, u32p, u32p, u32, u32p, u32, u32p, C.POINTER(vp)],
            "lm_air_new_shard": [vp, u32, C.POINTER(vp), u32, u32, u32p, u32p, u32, u32p, u32, u32p, u32p, u32p, C.POINTER(vp)],
            "lm_air_new_folded": [vp, u32, u32p, u32, u32, u32p, u32p, u32, u32p, u32, u32p, C.POINTER(vp)],
            "lm_air_info": [vp, u32p, u32p, u32p],
            "lm_air_round": [vp, u32p],
            "lm_air_fold": [vp, u32p],
            "lm_air_final": [vp, u32p],
            "lm_air_free": [vp],
            "lm_finger_print": [vp, u32p, u64, u32, u32p, u32p, u32p],
            "lm_gkr_new": [vp, u32p, u32p, u64, C.POINTER(vp)],
            "lm_gkr_new_dev": [vp, vp, vp, u64, C.POINTER(vp)],
            "lm_air_new_dev": [vp, u32, vp, u32, u32, u32p, u32p, u32, u32p, u32, u32p, C.POINTER(vp)],
            "lm_gkr_new_shard": [vp, u32p, u32p, u64, u32, u32, C.POINTER(vp)],
            "lm_gkr_layer_begin_shard": [vp, u32, u32p, u32p, u32p],
            "lm_gkr_num_vars": [vp, u32p],
            "lm_gkr_top_vars": [vp, u32p],
            "lm_gkr_top": [vp, u32p, u32p],
            "lm_gkr_layer_begin": [vp, u32, u32p, u32p],
            "lm_gkr_round": [vp, u32p, u32p],
            "lm_gkr_fold": [vp, u32p],
            "lm_gkr_layer_end": [vp, u32p],
            "lm_gkr_free": [vp],
            "lm_dev_alloc": [vp, sz, C.POINTER(vp)],
            "lm_dev_free": [vp, vp],
            "lm_dev_upload": [vp, vp, vp, sz],
            "lm_dev_download": [vp, vp, vp, sz],
            "lm_dev_poseidon1": [vp, vp, u64, i],
            "lm_dev_reorder_and_dft": [vp, vp, u32, u32, u32, u32, u32, vp],
            "lm_dev_reorder_and_dft_scatter": [vp, vp, u32, u32, u32, u32, vp, u64p, u32, u32],
            "lm_dev_reorder_and_dft_scatter_cols": [vp, vp, u32, u32, u32, u32, vp, u64p, u32, u32, u32, u32],
            "lm_dev_dft_layers_mapped_cols": [vp, vp, u64, u32, u32, u64, u64, u64, u64, u64, u64],
            "lm_dev_merkle_absorb": [vp, vp, u64, u32, u32, u32, u32, u32, vp],
            "lm_dev_ipc_export": [vp, vp, C.c_char_p],
            "lm_dev_ipc_open": [vp, C.c_char_p, C.POINTER(vp)],
            "lm_dev_ipc_close": [vp, vp],
            "lm_dev_dft": [vp, vp, u64, u64],
            "lm_dev_dft_layers_mapped": [vp, vp, u64, u32, u32, u64, u64, u64, u64],
            "lm_dev_dft_layers_mapped_out": [vp, vp, vp, u64, u32, u32, u64, u64, u64, u64, u64, u64],
            "lm_dev_merkle_tree": [vp, vp, u64, u32, u32, u32, vp],
            "lm_dev_merkle_leaves": [vp, vp, u64, u32, u32, u32, vp],
            "lm_dev_merkle_levels": [vp, vp, u64],
            "lm_dev_mle_eval": [vp, vp, u32, u32, u64, vp, vp],
            "lm_dev_fold_msb": [vp, vp, u64, u32, u32p, vp],
            "lm_dev_eq_table": [vp, u32p, u32, u32p, vp],
            "lm_pow_grind": [vp, u32p, u32, u64p],
            "lm_logup_new": [vp, u64, u32p, u32p, u32, C.POINTER(vp)],
            "lm_logup_section": [vp, u64, u32, vp, C.c_int32, u32, vp, u32],
            "lm_logup_col_eval": [vp, vp, u64, u32, u32p, u32p],
            "lm_logup_col_eval_batch": [vp, C.POINTER(vp), u64p, u32, u32, u32p, u32p],
            "lm_logup_read": [vp, vp, vp],
            "lm_logup_finish": [vp, C.POINTER(vp)],
            "lm_logup_free": [vp],
            "lm_poseidon16_fill_trace": [vp, C.POINTER(vp), u64],
            "lm_dev_poseidon16_fill_trace": [vp, vp, u64],
            "lm_host_poseidon1_permute": [u32p],
            "lm_host_poseidon1_umma_model": [u32p],
            "lm_host_poseidon1_umma_image": [vp, u64],
            "lm_host_eq_gemm_model": [u32p, u32p, u32p, u32, u32, u32],
            "lm_fs_new": [vp, C.POINTER(vp)],
            "lm_fs_free": [vp],
            "lm_fs_add_scalars": [vp, u32p, u64],
            "lm_fs_observe": [vp, u32p, u64],
            "lm_fs_duplex": [vp],
            "lm_fs_add_sumcheck_polynomial": [vp, u32p, u32, u32p],
            "lm_fs_sample": [vp, u32, u32p],
            "lm_fs_sample_in_range": [vp, u32, u32, u64p],
            "lm_fs_pow_grinding": [vp, u32],
            "lm_fs_transcript_len": [vp, u64p],
            "lm_fs_transcript": [vp, u32p],
            "lm_fs_state": [vp, u32p, C.POINTER(C.c_int)],
            "lm_gkr_prove": [vp, vp, u32p, u32p, u32p, u32p],
            "lm_sc_add_eq_batch": [vp, u64, u32p, u32, u32p, u32],
            "lm_sc_add_strided_eq": [vp, u64, u32, u64, u32p, u32, u32p],
            "lm_open_fold": [vp, u64p, u32, u32p, u32, u32p, u32p, u32p],
            "lm_whir_stir_update": [vp, u64p, u32, u32, u32, u32p, u32p, u32p, u32, u32p, u32p],
            "lm_whir_sumcheck_rounds": [vp, vp, u32, u32, u32p, u32p],
            "lm_gkr_prove_hostloop": [vp, vp, u32p, u32p, u32p, u32p],
            "lm_fs_set_state": [vp, u32p, i],
            "lm_air_prove_batched": [C.POINTER(vp), u32, u32p, u32p, u32p, vp, u32p, u32p],
        }
        u8p = C.POINTER(C.c_uint8)
        sig["lm_lz4_compress_prepend_size"] = [u8p, u64, u8p, u64, u64p]
        sig["lm_lz4_decompress_size_prepended"] = [u8p, u64, u8p, u64, u64p]
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.lm_lz4_compress_bound.argtypes = [u64]
        L.lm_lz4_compress_bound.restype = C.c_uint64
        L.lm_host_poseidon1_umma_image.restype = C.c_uint64
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise LmError(rc, lib().lm_last_error().decode())

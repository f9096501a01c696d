"""ctypes loader for the C-ABI engine (fetching_b200/lib/libpfetch.so; include/pfetch.h).

The product path fails loudly when the CUDA library is missing: there is no CPU fallback and
nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFETCH_LIB") or os.path.join(PKG, "lib", "libpfetch.so")  # PFETCH_LIB: experimental variants

PF_OK = 0
PF_ERR_INVALID, PF_ERR_CUDA, PF_ERR_SHAPE, PF_ERR_NOMEM, PF_ERR_UNSUPPORTED, PF_ERR_NCCL, PF_ERR_NO_DEVICE = (
    -1, -2, -3, -4, -5, -6, -7)
STATUS = {0: "PF_OK", -1: "PF_ERR_INVALID", -2: "PF_ERR_CUDA", -3: "PF_ERR_SHAPE", -4: "PF_ERR_NOMEM",
          -5: "PF_ERR_UNSUPPORTED", -6: "PF_ERR_NCCL", -7: "PF_ERR_NO_DEVICE"}

MODEL_NORMAL, MODEL_BOLTZMANN_SCALAR, MODEL_BOLTZMANN_DENSE, MODEL_MIXTURE, MODEL_BLASSO = 1, 2, 3, 4, 5
F64, F32 = 0, 1
OPT_ASSUME_IDENTITY_COV = 1
OPT_NODE_SPLIT = 2
OPT_PEER_EXCHANGE = 3
OPT_MATVEC_TENSOR = 4
OPT_INLINE_THETA = 5
OPT_RESIDENT = 6


class ModelDesc(C.Structure):
    """struct pf_model_desc (include/pfetch.h)."""
    _fields_ = [
        ("model", C.c_int32), ("dtype", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32),
        ("n_rows", C.c_uint64), ("dim", C.c_uint32), ("components", C.c_uint32),
        ("item_rows", C.c_uint64), ("n_items", C.c_uint64), ("param", C.c_double),
        ("data", C.POINTER(C.c_double)), ("response", C.POINTER(C.c_double)),
    ]


class ChainConfig(C.Structure):
    """struct pf_chain_config (include/pfetch.h)."""
    _fields_ = [
        ("frontier", C.c_uint32), ("scheduling_policy", C.c_uint32), ("chain_length", C.c_uint64),
        ("correlation", C.c_double), ("reassign_ratio", C.c_double), ("dimension", C.c_uint32),
        ("threshold", C.c_int32), ("verbose", C.c_int32), ("print_tsv", C.c_int32), ("seed", C.c_uint32),
        ("deferred_proposals", C.c_uint32), ("batch_items", C.c_uint64),
        ("update_style", C.c_uint32), ("reserved2", C.c_uint32), ("tau", C.c_double), ("initial_theta", C.POINTER(C.c_double)),
    ]


class ChainStats(C.Structure):
    """struct pf_chain_stats (include/pfetch.h)."""
    _fields_ = [
        ("levels", C.c_uint64), ("rows", C.c_uint64), ("engine_calls", C.c_uint64), ("nodes_scored", C.c_uint64),
        ("max_frontier", C.c_uint64), ("allocated_nodes", C.c_uint64), ("recycled_nodes", C.c_uint64),
        ("seconds", C.c_double), ("eval_seconds", C.c_double), ("items_scored", C.c_uint64),
        ("abandoned", C.c_uint64), ("accepted", C.c_uint64),
        ("min_margin", C.c_double), ("min_margin_rel", C.c_double),
    ]


_dp = C.POINTER(C.c_double)
ROW_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_int, C.c_double, _dp, C.c_uint32)
JOB_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_char_p, C.c_uint64, C.c_uint64,
                     C.c_uint64, C.c_double, _dp, C.c_uint32)
EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, _dp, _dp, _dp)
RANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, _dp, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), _dp, _dp)

# every symbol include/pfetch.h declares: (restype, argtypes)
SYMBOLS = {
    "pf_engine_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "pf_engine_destroy": (None, [C.c_void_p]),
    "pf_engine_data_size": (C.c_uint64, [C.c_void_p]),
    "pf_engine_theta_len": (C.c_uint32, [C.c_void_p]),
    "pf_engine_max_frontier": (C.c_uint32, [C.c_void_p]),
    "pf_engine_model": (C.c_int, [C.c_void_p]),
    "pf_engine_dim": (C.c_uint32, [C.c_void_p]),
    "pf_engine_components": (C.c_uint32, [C.c_void_p]),
    "pf_engine_group": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pf_chain_run": (C.c_int, [C.c_void_p, C.POINTER(ChainConfig), ROW_FN, JOB_FN, C.c_void_p,
                               C.POINTER(ChainStats)]),
    "pf_chain_run_with_evaluator": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(ChainConfig),
                                              EVAL_FN, RANGE_FN, C.c_void_p, ROW_FN, JOB_FN, C.c_void_p,
                                              C.POINTER(ChainStats)]),
    "pf_eval_frontier": (C.c_int, [C.c_void_p, C.c_uint32, _dp, _dp, _dp]),
    "pf_eval_frontier_repeat": (C.c_int, [C.c_void_p, C.c_uint32, _dp, _dp, _dp, C.c_uint32, _dp]),
    "pf_eval_frontier_range": (C.c_int, [C.c_void_p, C.c_uint32, _dp, C.c_uint64, C.c_uint64,
                                         C.POINTER(C.c_uint64), _dp, _dp]),
    "pf_eval_frontier_device": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pf_blasso_log_prior": (C.c_double, [_dp, C.c_uint32, C.c_double]),
    "pf_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "pf_engine_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pf_engine_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pf_engine_peer_attach": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pf_engine_peer_attach_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "pf_engine_peer_status": (C.c_int, [C.c_void_p]),
    "pf_engine_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "pf_abi_version": (C.c_int, []),
    "pf_last_error": (C.c_char_p, []),
    "pf_engine_launch_count": (C.c_uint64, [C.c_void_p]),
    "pf_engine_resident_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pf_device_sm_count": (C.c_int, [C.c_int]),
    "pf_debug_math": (C.c_int, [C.c_int, _dp, _dp, C.c_size_t, C.c_int]),
    "pf_debug_fp64_probe": (C.c_int, [C.c_int, _dp]),
    "pf_debug_dmma_probe": (C.c_int, [C.c_int, _dp]),
    "pf_debug_issue_probe": (C.c_int, [C.c_int, _dp]),
    "pf_debug_rf_probe": (C.c_int, [C.c_int, _dp]),
    "pf_debug_dfma_dmma_probe": (C.c_int, [C.c_int, _dp]),
    "pf_debug_launch_probe": (C.c_int, [C.c_int, _dp]),
    "pf_engine_theta_buffer": (_dp, [C.c_void_p, C.POINTER(C.c_uint64)]),
}

_lib = None


class PfetchError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen libpfetch.so and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PfetchError(
                "fetching_b200/lib/libpfetch.so is missing: build it with `python -m fetching_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != PF_OK:
        msg = load().pf_last_error()
        raise PfetchError("%s: %s (%s)" % (what, STATUS.get(status, status), msg.decode() if msg else ""))

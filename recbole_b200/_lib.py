"""ctypes binding of librecbole_b200.so (the C ABI declared in include/recbole_b200.h).

There is NO fallback: if the shared library has not been built (python recbole_b200/build.py, or
__graft_entry__.build()), importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librecbole_b200.so")

ABI_VERSION = 1

# enums of include/recbole_b200.h
OPT_SGD, OPT_ADAM, OPT_ADAM_LAZY = 0, 1, 2
SCORER_FP32, SCORER_TC = 0, 1
M_RECALL, M_MRR, M_NDCG, M_HIT, M_PRECISION, M_MAP = range(6)
NUM_METRICS = 6
METRIC_ORDER = ("recall", "mrr", "ndcg", "hit", "precision", "map")
EINVAL, EWORKSPACE, ERANGE = 10001, 10002, 10003
STAGES = ("keys", "sort_user", "sort_item", "user_side", "user_fixup", "item_side", "item_fixup", "loss", "fullsort",
          "topk_merge", "metrics", "sampler", "gather_dot", "tc_convert", "tc_score", "tc_refine", "fm_fwd",
          "fm_update", "misc", "plan", "barrier", "owner", "fetch", "barrier_b")
MAX_PEERS = 8


class RB2Optim(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32), ("step", ctypes.c_int32),
        ("lr", ctypes.c_float), ("weight_decay", ctypes.c_float),
        ("beta1", ctypes.c_float), ("beta2", ctypes.c_float),
        ("one_minus_beta1", ctypes.c_float), ("one_minus_beta2", ctypes.c_float),
        ("eps", ctypes.c_float), ("step_size", ctypes.c_float), ("bc2_sqrt", ctypes.c_float),
        ("lazy_step_size", ctypes.c_void_p), ("lazy_bc2_sqrt", ctypes.c_void_p),
    ]


class RB2ScorerState(ctypes.Structure):
    """include/recbole_b200.h: rb2_scorer_state (caller-owned knobs + adaptive statistics + last-call counters)."""
    _fields_ = [("variant", ctypes.c_int32), ("kprime", ctypes.c_int32), ("ce_scorer", ctypes.c_int32),
                ("fail_ema", ctypes.c_float), ("calls", ctypes.c_int32), ("last_fallback_rows", ctypes.c_int32),
                ("last_pass2_rows", ctypes.c_int32), ("reserved", ctypes.c_int32), ("trace", ctypes.c_void_p)]


class RB2FmFloat(ctypes.Structure):
    """include/recbole_b200.h: rb2_fm_float (the FLOAT fields of a context-aware model)."""
    _fields_ = [("values", ctypes.c_void_p), ("n_float", ctypes.c_int32), ("Ef", ctypes.c_void_p),
                ("mEf", ctypes.c_void_p), ("vEf", ctypes.c_void_p), ("Wf", ctypes.c_void_p), ("mWf", ctypes.c_void_p),
                ("vWf", ctypes.c_void_p)]


class RB2FmSeq(ctypes.Structure):
    """include/recbole_b200.h: rb2_fm_seq (the TOKEN_SEQ fields of a context-aware model)."""
    _fields_ = [("n_seq", ctypes.c_int32), ("n_token_cols", ctypes.c_int32), ("seq_start", ctypes.c_void_p),
                ("col_seq", ctypes.c_void_p), ("seq_row_base", ctypes.c_int64), ("pooled", ctypes.c_void_p),
                ("coef", ctypes.c_void_p)]


class RB2Peers(ctypes.Structure):
    _fields_ = [
        ("world", ctypes.c_int32), ("me", ctypes.c_int32), ("item_block", ctypes.c_int64),
        ("item_p", ctypes.c_void_p * MAX_PEERS), ("grad_slots", ctypes.c_void_p * MAX_PEERS),
        ("stamps", ctypes.c_void_p * MAX_PEERS), ("flags", ctypes.c_void_p * MAX_PEERS),
        ("loss_slots", ctypes.c_void_p * MAX_PEERS), ("seq", ctypes.c_int64),
    ]


_p, _i64, _i32, _sz, _u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_size_t, ctypes.c_uint64

# name -> (restype, argtypes); every symbol include/recbole_b200.h declares
SIGNATURES = {
    "rb2_abi_version": (ctypes.c_int, []),
    "rb2_last_error": (ctypes.c_char_p, []),
    "rb2_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "rb2_profile_read": (ctypes.c_int, [_p, _p, _p]),
    "rb2_bpr_workspace_bytes": (_sz, [_i64, _i32]),
    "rb2_bpr_train_step": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _i64,
                                          ctypes.POINTER(RB2Optim), _p, _p, _p, _sz, _p]),
    "rb2_bpr_train_step_sharded": (ctypes.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _i64, _i64,
                                                  ctypes.POINTER(RB2Optim), _p, _p, _p, _p, _p, _p, _sz, _p]),
    "rb2_bpr_train_step_sharded_ev": (ctypes.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _i64, _i64,
                                                     ctypes.POINTER(RB2Optim), _p, _p, _p, _p, _p, _p, _sz, _p, _p]),
    "rb2_bpr_p2p_workspace_bytes": (_sz, [_i64, _i32]),
    "rb2_bpr_train_step_p2p": (ctypes.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i32, _p, _i64, _p, _p, _i64, _i64,
                                              ctypes.POINTER(RB2Optim), ctypes.POINTER(RB2Peers), _p, _p, _p, _p, _sz,
                                              _p, _i32, _p, _p, _p, _p]),
    "rb2_ipc_export": (ctypes.c_int, [_p, _p, ctypes.POINTER(_i64)]),
    "rb2_ipc_open": (ctypes.c_int, [_p, _i64, ctypes.POINTER(_p)]),
    "rb2_ipc_close_all": (ctypes.c_int, []),
    "rb2_item_plan_workspace_bytes": (_sz, [_i64]),
    "rb2_item_plan": (ctypes.c_int, [_p, _p, _i64, _i64, _p, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "rb2_dense_rows_update": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _p, _p, ctypes.POINTER(RB2Optim), _p]),
    "rb2_sparse_rows_update_workspace_bytes": (_sz, [_i64, _i32]),
    "rb2_sparse_rows_update": (ctypes.c_int, [_p, _p, _p, _p, _i64, _i32, _p, _p, _i64, ctypes.POINTER(RB2Optim),
                                              _p, _sz, _p]),
    "rb2_bpr_loss": (ctypes.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _i64, _p, _p, _sz, _p]),
    "rb2_sort_positions_workspace_bytes": (_sz, [_i64]),
    "rb2_sort_positions": (ctypes.c_int, [_p, _i64, _i32, _p, _p, _p, _sz, _p]),
    "rb2_adam_lazy_flush": (ctypes.c_int, [_p, _p, _p, _p, _i64, _i32, ctypes.POINTER(RB2Optim), _p]),
    "rb2_fm_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rb2_fm_train_step": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p, _p, _i32, _p, _i64,
                                         ctypes.POINTER(RB2Optim), _p, _p, _p, _sz, _p, ctypes.POINTER(RB2FmFloat),
                                         ctypes.POINTER(RB2FmSeq)]),
    "rb2_fm_lazy_flush": (ctypes.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, ctypes.POINTER(RB2Optim), _p]),
    "rb2_fm_predict": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _i32, _i64, _p, _p, _sz, _p,
                                      ctypes.POINTER(RB2FmFloat), ctypes.POINTER(RB2FmSeq)]),
    "rb2_fm_loss": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _i32, _p, _i64, _p, _p, _sz, _p,
                                   ctypes.POINTER(RB2FmFloat), ctypes.POINTER(RB2FmSeq)]),
    "rb2_fm_grad_step": (ctypes.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _i32, _p, _i64, _i64, _p, _p, _sz, _p]),
    "rb2_scalar_rows_update_workspace_bytes": (_sz, [_i64]),
    "rb2_scalar_rows_update": (ctypes.c_int, [_p, _p, _p, _i64, _p, _p, _i64, ctypes.POINTER(RB2Optim), _p, _sz, _p]),
    "rb2_scalar_step": (ctypes.c_int, [_p, _p, ctypes.POINTER(RB2Optim), _p]),
    "rb2_gather_dot": (ctypes.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _i64, _p, _p]),
    "rb2_fullsort_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "rb2_fullsort_topk": (ctypes.c_int, [_p, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _i32, _i32, _p, _p, _p, _sz,
                                         _p]),
    "rb2_fullsort_scores": (ctypes.c_int, [_p, _p, _i64, _i64, _p, _i64, _i32, _p, _p]),
    "rb2_fullsort_topk_s": (ctypes.c_int, [_p, _p, _i64, _p, _i64, _i64, _i32, _p, _p, _i32, _i32, _p, _p, _p, _sz,
                                           _p, ctypes.POINTER(RB2ScorerState)]),
    "rb2_ce_head_s": (ctypes.c_int, [_p, _i64, _p, _i64, _i32, _p, _i32, _p, _p, _p, _p, _p, _sz, _p,
                                     ctypes.POINTER(RB2ScorerState)]),
    "rb2_ce_head_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "rb2_ce_head": (ctypes.c_int, [_p, _i64, _p, _i64, _i32, _p, _i32, _p, _p, _p, _p, _p, _sz, _p]),
    "rb2_ce_head_set_scorer": (ctypes.c_int, [_i32]),
    "rb2_ce_head_backward_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "rb2_ce_head_backward": (ctypes.c_int, [_p, _i64, _p, _i64, _i32, _p, _p, ctypes.c_float, _p, _p, _p, _sz, _p]),
    "rb2_dense_step": (ctypes.c_int, [_p, _p, _p, _p, _i64, ctypes.POINTER(RB2Optim), _p]),
    "rb2_fullsort_tc_last_fallback_rows": (_i32, []),
    "rb2_fullsort_tc_last_pass2_rows": (_i32, []),
    "rb2_fullsort_tc_set_kprime": (ctypes.c_int, [_i32]),
    "rb2_fullsort_tc_set_variant": (ctypes.c_int, [_i32]),
    "rb2_fullsort_tc_set_trace": (ctypes.c_int, [_p]),
    "rb2_topk_merge": (ctypes.c_int, [_p, _p, _i32, _i64, _i32, _p, _p, _p]),
    "rb2_topk_metrics_workspace_bytes": (_sz, [_i64, _i32]),
    "rb2_topk_metrics": (ctypes.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "rb2_neg_sample_workspace_bytes": (_sz, [_i64, _i32]),
    "rb2_neg_sample_ref": (ctypes.c_int, [_p, _i64, _i32, _p, _i64, ctypes.POINTER(_i64), _p, _p, _i64, _p, _p,
                                          _sz, _p]),
    "rb2_neg_sample_hash": (ctypes.c_int, [_p, _i64, _i32, _i64, _p, _p, _i64, _u64, _u64, _p, _p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "recbole_b200: %s is missing. Build it with `python recbole_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.rb2_abi_version() != ABI_VERSION:
        raise ImportError("recbole_b200: ABI version mismatch (library %d, binding %d)"
                          % (lib.rb2_abi_version(), ABI_VERSION))
    return lib


lib = _load()


class RB2Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = lib.rb2_last_error().decode("utf-8", "replace")
        raise RB2Error("recbole_b200 C ABI call failed (code %d): %s" % (rc, msg))

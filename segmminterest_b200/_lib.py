"""ctypes binding of libmmi_b200.so (include/mmi_b200.h).  Fails loudly when the library
is missing: there is no CPU / eager fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMI_LIB_PATH") or os.path.join(HERE, "libmmi_b200.so")   # override: A/B builds of the kernels (tools/)

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
IMPL_SIMT, IMPL_TC = 0, 1

c_p = C.c_void_p
i64 = C.c_int64


class Dropout(C.Structure):
    """mmi_dropout: one dropout site (csrc/dropout.cuh); thr8 = 0 is off"""
    _fields_ = [("key", C.c_uint32), ("thr8", C.c_uint32), ("scale", C.c_float)]


class GemmArgs(C.Structure):
    _fields_ = [("layout", C.c_int), ("impl", C.c_int), ("in_dtype", C.c_int), ("out_dtype", C.c_int),
                ("M", i64), ("N", i64), ("K", i64),
                ("A", c_p), ("lda", i64), ("B", c_p), ("ldb", i64), ("C", c_p), ("ldc", i64),
                ("bias", c_p), ("act", C.c_int),
                ("preact", c_p), ("ld_preact", i64),
                ("mul_gelu_grad", c_p), ("ld_mul", i64),
                ("add", c_p), ("ld_add", i64), ("add_mod", i64), ("add_dtype", C.c_int),
                ("accumulate", C.c_int), ("split_k", C.c_int), ("save_act_grad", C.c_int), ("mul_is_grad", C.c_int),
                ("drop", Dropout), ("mul_scale", C.c_float),
                ("split_ws", c_p), ("split_ws_bytes", i64)]


class AttnBlock(C.Structure):
    _fields_ = [("q", c_p), ("ldq", i64), ("k", c_p), ("ldk", i64), ("v", c_p), ("ldv", i64),
                ("mask_k", c_p), ("Lk", C.c_int),
                ("dq", c_p), ("lddq", i64), ("dk", c_p), ("lddk", i64), ("dv", c_p), ("lddv", i64),
                ("dbq", c_p), ("dbk", c_p), ("dbv", c_p)]


class AttnArgs(C.Structure):
    _fields_ = [("dtype", C.c_int), ("impl", C.c_int),
                ("B", C.c_int), ("H", C.c_int), ("dh", C.c_int), ("Lq", C.c_int),
                ("mask_q", c_p), ("nblk", C.c_int), ("blk", AttnBlock * 2),
                ("out", c_p), ("ldo", i64), ("lse", c_p),
                ("dout", c_p), ("lddo", i64), ("delta", c_p),
                ("drop", Dropout),
                ("dq_acc", c_p * 2), ("dq_count", c_p * 2)]


class LossArgs(C.Structure):
    _fields_ = [("logits", c_p), ("gt", c_p), ("B", C.c_int), ("L", C.c_int), ("exposure_prob", c_p),
                ("bias_weight", c_p), ("bias_bias", c_p),
                ("inv_bsz", C.c_float), ("w_focal", C.c_float), ("w_bpr", C.c_float), ("bpr_scale", C.c_float),
                ("use_focal", C.c_int), ("use_bpr", C.c_int), ("rewrite_gt", C.c_int),
                ("logits_out", c_p), ("scalars", c_p), ("dlogits", c_p), ("dbias_weight", c_p), ("dbias_bias", c_p),
                ("use_huber", C.c_int), ("use_hazard", C.c_int), ("use_surviveCE", C.c_int), ("use_interestCE", C.c_int),
                ("use_interestKL", C.c_int), ("mask_loss", C.c_int), ("ce_after_focal", C.c_int), ("kl_after_focal", C.c_int),
                ("w_huber", C.c_float), ("w_hazard", C.c_float), ("w_surviveCE", C.c_float), ("w_interestCE", C.c_float),
                ("w_interestKL", C.c_float)]


_SIGS = {
    "mmi_version": (C.c_int, []),
    "mmi_last_error": (C.c_char_p, []),
    "mmi_has_tc": (C.c_int, []),
    "mmi_gather_l1norm_fwd": (C.c_int, [c_p, C.c_int, i64, C.c_int, c_p, i64, c_p, C.c_int, c_p, C.c_int, c_p]),
    "mmi_gemm": (C.c_int, [C.POINTER(GemmArgs), c_p]),
    "mmi_gemm_split_workspace": (i64, [C.c_int, i64, i64, i64]),
    "mmi_adaptive_pool_fwd": (C.c_int, [c_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p]),
    "mmi_adaptive_pool_bwd": (C.c_int, [c_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p]),
    "mmi_colsum_acc": (C.c_int, [c_p, C.c_int, i64, C.c_int, i64, c_p, c_p, i64, c_p]),
    "mmi_layernorm_fwd": (C.c_int, [c_p, C.c_int, i64, C.c_int, c_p, c_p, C.c_float, c_p, c_p, c_p]),
    "mmi_layernorm_fwd_drop": (C.c_int, [c_p, C.c_int, i64, C.c_int, c_p, c_p, C.c_float, c_p, c_p, C.POINTER(Dropout), c_p]),
    "mmi_layernorm_bwd_drop": (C.c_int, [c_p, c_p, C.c_int, i64, C.c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                         C.POINTER(Dropout), C.POINTER(Dropout), c_p, c_p]),
    "mmi_dropout_mask": (C.c_int, [C.POINTER(Dropout), i64, i64, C.c_int, C.c_uint32, c_p, c_p]),
    "mmi_layernorm_bwd_workspace": (i64, [C.c_int]),
    "mmi_layernorm_bwd": (C.c_int, [c_p, c_p, C.c_int, i64, C.c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmi_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), c_p]),
    "mmi_attn_bwd_dq": (C.c_int, [C.POINTER(AttnArgs), c_p]),
    "mmi_attn_bwd_dkv": (C.c_int, [C.POINTER(AttnArgs), C.c_int, c_p]),
    "mmi_attn_bwd_fused": (C.c_int, [C.POINTER(AttnArgs), C.c_int, c_p]),
    "mmi_attn_bwd_all": (C.c_int, [C.POINTER(AttnArgs), c_p]),
    "mmi_head_fwd": (C.c_int, [c_p, C.c_int, i64, C.c_int, c_p, c_p, c_p, c_p, c_p]),
    "mmi_head_bwd_workspace": (i64, [C.c_int]),
    "mmi_head_bwd": (C.c_int, [c_p, C.c_int, i64, C.c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "mmi_focal_loss_fwd_bwd": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, C.c_float, C.c_float, C.c_int, c_p, c_p, c_p]),
    "mmi_loss_fwd_bwd": (C.c_int, [C.POINTER(LossArgs), c_p]),
    "mmi_id_embed_fwd": (C.c_int, [c_p, i64, C.c_int, c_p, C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, c_p, c_p, C.c_int, c_p]),
    "mmi_id_embed_bwd": (C.c_int, [c_p, C.c_int, c_p, i64, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, c_p, c_p]),
    "mmi_id_rows_bwd": (C.c_int, [c_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p, c_p, c_p, c_p]),
    "mmi_scatter_rows_add": (C.c_int, [c_p, c_p, i64, C.c_int, i64, c_p, c_p]),
    "mmi_rowdot_fwd": (C.c_int, [c_p, i64, c_p, i64, C.c_int, i64, C.c_int, c_p, c_p, c_p, c_p]),
    "mmi_rowdot_bwd": (C.c_int, [c_p, c_p, c_p, i64, c_p, i64, C.c_int, i64, C.c_int, c_p, c_p, c_p, c_p]),
    "mmi_clip_adamw_workspace": (i64, [i64]),
    "mmi_eval_metrics_workspace": (i64, [C.c_int, C.c_int]),
    "mmi_eval_metrics": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, C.c_int, c_p, c_p, c_p, c_p]),
    "mmi_clip_adamw": (C.c_int, [c_p, c_p, c_p, c_p, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_int, c_p, c_p, c_p, c_p]),
    "mmi_cast_bf16": (C.c_int, [c_p, c_p, i64, i64, C.c_int, c_p]),
}

EXPORTS = tuple(_SIGS)

_lib = None


class MMIError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MMIError(f"{LIB_PATH} is missing: run `python -m segmminterest_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mmi_last_error().decode()
        raise MMIError(f"{what} failed (rc={rc}): {msg}")

"""ctypes binding of the C-ABI CUDA library (include/vqa_b200.h).

There is deliberately no fallback: if ``csrc/libvqa_b200.so`` is missing and cannot be built, or a
kernel returns a non-zero status, a ``RuntimeError`` is raised.  Nothing in this package routes
through PyTorch eager ops or the CPU oracle for the hot path.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p

from . import build as _build

K_MAJOR, MN_MAJOR = 0, 1
F32, BF16 = 0, 1

_lock = threading.Lock()
_lib = None

_PROTOTYPES = {
    "vqa_b200_abi_version": (c_int, []),
    "vqa_b200_last_error": (c_char_p, []),
    "vqa_b200_gemm": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_int, c_int64,
                              c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              c_void_p, c_int64, c_void_p, c_void_p]),
    "vqa_b200_mfb_fused": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                   c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_dropout_mask": (c_int, [c_void_p, c_int, c_int, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_pack_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                   c_int64, c_int64, c_void_p]),
    "vqa_b200_split3_bf16": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                     c_int64, c_int, c_int, c_void_p]),
    "vqa_b200_attn_logits_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         c_int, c_void_p]),
    "vqa_b200_attn_logits_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                         c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p]),
    "vqa_b200_softmax_pool_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_int, c_int, c_void_p]),
    "vqa_b200_softmax_pool_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vqa_b200_mfb_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p, c_void_p, c_void_p, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_norm_bwd_prep": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64,
                                       c_void_p, c_int, c_int, c_int, c_void_p]),
    "vqa_b200_inv_norm": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "vqa_b200_scale_rows": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_int64, c_int, c_int,
                                    c_void_p]),
    "vqa_b200_group_dot": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_int, c_int,
                                   c_int, c_void_p]),
    "vqa_b200_colsum": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "vqa_b200_relu_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_int, c_int64,
                                  c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "vqa_b200_gemm_batched": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int64, c_int64,
                                      c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                      c_void_p, c_int, c_float, c_uint32, c_void_p, c_int, c_void_p]),
    "vqa_b200_act_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_uint32,
                                 c_void_p, c_void_p]),
    "vqa_b200_act_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_void_p, c_int, c_int64,
                                 c_void_p, c_int, c_int, c_int, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_row_softmax_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vqa_b200_row_softmax_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vqa_b200_gate_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vqa_b200_lstm_supported": (c_int, [c_int, c_int]),
    "vqa_b200_lstm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                  c_uint32, c_void_p, c_void_p]),
    "vqa_b200_lstm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int, c_int,
                                  c_int, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_lstm_cell_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int,
                                       c_int64, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_lstm_cell_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_int, c_int, c_int64, c_float, c_uint32, c_void_p, c_void_p]),
    "vqa_b200_adam_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double,
                                   c_double, c_double, c_int64, c_void_p]),
    "vqa_b200_adam_step_dev": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                                       c_double, c_double, c_double, c_void_p, c_void_p]),
    "vqa_b200_logsoftmax_argmax": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int,
                                           c_void_p]),
    "vqa_b200_kldiv_logsoftmax_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                              c_int, c_void_p]),
    "vqa_b200_kldiv_logsoftmax_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_int64, c_int, c_int, c_void_p]),
    "vqa_b200_gate_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
}

# Pointer arguments (0-based positions) that an entry point WRITES: the `mutates_args` of the torch custom op that wraps
# it (ops._register_custom_ops: one `torch.ops.vqa_b200.<name>` per kernel entry point, schema derived from the prototype
# above -- pointers are `Tensor?`, the trailing stream is supplied by the op from the caller's current stream).
MUTATED_ARGS = {
    "vqa_b200_gemm": (6, 20),
    "vqa_b200_mfb_fused": (8, 11, 12, 19),
    "vqa_b200_dropout_mask": (0,),
    "vqa_b200_pack_bf16": (1,),
    "vqa_b200_split3_bf16": (3,),
    "vqa_b200_attn_logits_fwd": (5,),
    "vqa_b200_attn_logits_bwd": (5, 11, 12, 13),
    "vqa_b200_softmax_pool_fwd": (3, 4),
    "vqa_b200_softmax_pool_bwd": (5, 6, 7),
    "vqa_b200_mfb_bwd": (12, 14, 15, 22),
    "vqa_b200_norm_bwd_prep": (6, 8),
    "vqa_b200_inv_norm": (1,),
    "vqa_b200_scale_rows": (5,),
    "vqa_b200_group_dot": (6,),
    "vqa_b200_colsum": (3,),
    "vqa_b200_relu_bwd": (6, 11),
    "vqa_b200_gemm_batched": (8,),
    "vqa_b200_act_fwd": (3,),
    "vqa_b200_act_bwd": (6, 9),
    "vqa_b200_row_softmax_fwd": (1,),
    "vqa_b200_row_softmax_bwd": (2,),
    "vqa_b200_gate_fwd": (2,),
    "vqa_b200_gate_bwd": (3, 4),
    "vqa_b200_lstm_fwd": (0, 2, 3, 4),
    "vqa_b200_lstm_bwd": (7,),
    "vqa_b200_lstm_cell_fwd": (0, 2, 3, 5),
    "vqa_b200_lstm_cell_bwd": (5, 6, 7),
    "vqa_b200_logsoftmax_argmax": (2, 4, 5),
    "vqa_b200_kldiv_logsoftmax_fwd": (4, 5, 6),
    "vqa_b200_kldiv_logsoftmax_bwd": (7,),
}

# exported by -DVQA_B200_DEBUG builds only (include/vqa_b200.h, last section): bound when present, never required
_DEBUG_PROTOTYPES = {
    "vqa_b200_debug_set_mn_desc": (None, [c_uint32, c_uint32, c_uint32]),
    "vqa_b200_debug_set_counters": (None, [c_void_p]),
    "vqa_b200_debug_set_lstm": (None, [c_void_p, c_int]),
}

EXPORTED_SYMBOLS = tuple(sorted(_PROTOTYPES))


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if needed) the CUDA library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if not _build.up_to_date():
            if os.path.isfile(path) and _build._nvcc() is None:
                pass          # prebuilt library shipped without a toolchain on this box: use it as is
            else:
                path = _build.build()
        if not os.path.isfile(path):
            raise RuntimeError("vqa_b200: %s is missing and could not be built; there is no fallback path" % path)
        lib = ctypes.CDLL(path)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)     # AttributeError here == header / library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        for name, (res, args) in _DEBUG_PROTOTYPES.items():
            fn = getattr(lib, name, None)
            if fn is not None:
                fn.restype = res
                fn.argtypes = args
        if lib.vqa_b200_abi_version() != 2:
            raise RuntimeError("vqa_b200: ABI version mismatch")
        _lib = lib
        return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().vqa_b200_last_error()
        raise RuntimeError("vqa_b200 %s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))

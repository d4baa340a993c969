"""ctypes binding of libkws_b200.so (the C ABI declared in include/kws_b200.h).

There is no fallback: if the shared library is missing or a call fails, this
module raises.  Build it with ``python enhance-cb-whisper_b200/build.py`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# KWS_B200_LIB: explicit path of another flavour of the same ABI (the -DKWS_DEBUG_HOOKS development build)
LIB_PATH = os.environ.get("KWS_B200_LIB") or os.path.join(HERE, "libkws_b200.so")
ABI_VERSION = 8

# constants of include/kws_b200.h
F16, BF16 = 0, 1
MLP_OUT_NORM_F16, MLP_OUT_RAW_F32, MLP_OUT_RAW_16 = 0, 1, 2
PAIRS_ALL, PAIRS_DIAG, PAIRS_PER_KEYWORD = 0, 1, 2
STEM_OUT_NCHW_F32, STEM_OUT_NHWC_BF16, STEM_OUT_POOL_NHWC_BF16 = 0, 1, 2

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/kws_b200.h one to one
SIGNATURES = {
    "kws_abi_version": (_i, []),
    "kws_last_error": (C.c_char_p, []),
    "kws_sm_count": (_i, []),
    "kws_pack_stem_weights": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "kws_stem_weight_bytes": (_sz, [_i]),
    "kws_pack_stem_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "kws_stem_fused_weight_bytes": (_sz, [_i]),
    "kws_fold_temporal_weights": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp]),
    "kws_cast_f32_to_16": (_i, [_vp, _vp, _sz, _i, _vp]),
    "kws_normalize_rows": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_int32), _i, _vp, _f, _vp, _vp]),
    "kws_cast_rows16": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_int32), _i, _i, _vp, _vp]),
    "kws_mlp": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp]),
    "kws_mlp_fused": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_int32), _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp]),
    "kws_mlp_fused_supported": (_i, [_i, _i, _i]),
    "kws_temporal": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp]),
    "kws_sim": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "kws_stem": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "kws_stem_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "kws_sim_stem": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "kws_sim_stem_supported": (_i, [_i, _i, _i, _i]),
    "kws_sim_stem_range": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "kws_sim_stem_ragged": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "kws_sim_stem_pool": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "kws_sim_stem_pool_workspace_bytes": (_sz, [_i, C.c_longlong, _i, _i]),
    "kws_resize_bilinear": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "kws_interp_rows": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_int32), _i, _i, _f, _vp, _vp]),
    "kws_sim_operand": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "kws_resize_row_weights": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "kws_maxpool_nhwc": (_i, [_vp, C.c_longlong, _i, _i, _i, _vp, _vp]),
    "kws_scores": (_i, [_vp, _vp, _sz, _f, _vp, _vp, _vp]),
    "kws_topk": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "kws_topk_workspace_bytes": (_sz, [_i, _i, _i]),
}


class KWSError(RuntimeError):
    """A C-ABI call returned non-zero (message from kws_last_error())."""


_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise KWSError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no fallback path. "
                "Run `python enhance-cb-whisper_b200/build.py`."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        v = lib.kws_abi_version()
        if v != ABI_VERSION:
            raise KWSError(f"libkws_b200.so ABI version {v} != expected {ABI_VERSION}")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().kws_last_error()
        raise KWSError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")

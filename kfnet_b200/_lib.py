"""ctypes binding of libkfnet_b200.so (include/kfnet_b200.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkfnet_b200.so")

KFN_MODE_FP32, KFN_MODE_BF16, KFN_MODE_BF16X3, KFN_MODE_FP16X2 = 0, 1, 2, 3
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> (restype, argtypes)  -- one entry per KFN_API declaration in include/kfnet_b200.h
SIGNATURES = {
    "kfn_version": (C.c_char_p, []),
    "kfn_last_error": (C.c_char_p, []),
    "kfn_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "kfn_destroy": (_i, [_vp]),
    "kfn_load_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    "kfn_finalize_weights": (_i, [_vp]),
    "kfn_workspace_bytes": (_i64, [_vp, _i, _i]),
    "kfn_scoordnet_forward": (_i, [_vp, _vp, _i, _vp, _vp, _i64, _vp]),
    "kfn_oflow_feat_forward": (_i, [_vp, _vp, _i, _vp, _vp, _i64, _vp]),
    "kfn_diff_volume": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "kfn_oflownet_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kfn_kalman_scan": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kfn_flow_warp_kalman": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kfn_bilinear_sampler": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "kfn_apply_transform": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "kfn_kalman_update": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "kfn_sequence_extra_bytes": (_i64, [_vp, _i, _i]),
    "kfn_eval_sequences": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "kfn_eval_sequences_host": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kfn_eval_stream_host": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "kfn_conv_layer": (_i, [_vp, C.c_char_p, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "kfn_train_begin": (_i, [_vp, _i, _i, _i]),
    "kfn_train_step": (_i, [_vp, _vp, _vp, _vp, C.c_float, C.c_float, _i, C.c_float, _i, _vp, _vp]),
    "kfn_train_get": (_i, [_vp, C.c_char_p, _i, _vp, _i64]),
    "kfn_train_commit": (_i, [_vp]),
    "kfn_train_set_step": (_i, [_vp, _i]),
    "kfn_launch_count": (_i64, [_vp, _i]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(kfnet_b200 has no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class KfnError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise KfnError(f"kfn status {rc}: {load().kfn_last_error().decode()}")

"""ctypes binding of libnerftiny.so (include/nerftiny.h).

Loading never falls back to a CPU implementation: if the shared library is missing the import of the
product path raises, and `nt_create` fails on a box without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NT_LIB_PATH") or os.path.join(HERE, "libnerftiny.so")  # NT_LIB_PATH: A/B builds (tools/)

NT_OK = 0
NT_ERR_RANGE = -3
PREC_FP32 = 0
PREC_TC32 = 1
PREC_BF16 = 2
PREC_FP16 = 3
PREC_MIXED = 4
ANY_STEP_ZERO_DEVICE = -2
OPT_DETACH_T_FINE = 1
OPT_MLP_TC_VERSION = 2
OPT_LAST_DELTA = 3
OPT_DW_OVERLAP_CTAS = 4
N_PARAMS = 593924
N_LAYERS = 12

vp, i64, i32, f32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t


class LayerDesc(C.Structure):
    _fields_ = [("out_features", C.c_int), ("in_features", C.c_int), ("weight_offset", C.c_int64),
                ("bias_offset", C.c_int64)]


# name -> (restype, argtypes); mirrors include/nerftiny.h one to one
SIGNATURES = {
    "nt_version": (i32, []),
    "nt_last_error": (C.c_char_p, []),
    "nt_create": (i32, [C.POINTER(vp), i32, i32, i32]),
    "nt_destroy": (None, [vp]),
    "nt_param_count": (i64, []),
    "nt_layer_table": (i32, [C.POINTER(LayerDesc)]),
    "nt_launch_count": (i64, [vp]),
    "nt_set_option": (i32, [vp, i32, i32]),
    "nt_raygen": (i32, [vp, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    "nt_encode": (i32, [vp, i64, vp, vp, vp, vp, vp]),
    "nt_network_forward": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, sz, vp]),
    "nt_sample_coarse": (i32, [vp, i64, vp, vp, i32, vp, vp]),
    "nt_shard_globals_local": (i32, [vp, i64, vp, vp, i32, vp, vp]),
    "nt_shard_globals_resolve": (i32, [vp, vp, vp]),
    "nt_mlp_workspace_bytes": (sz, [vp, i32, i64, i32, i32]),
    "nt_packed_weight_bytes": (sz, [vp, i32]),
    "nt_pack_weights": (i32, [vp, i32, vp, vp, vp]),
    "nt_mlp_forward": (i32, [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, sz, i32, vp]),
    "nt_mlp_backward": (i32, [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "nt_mlp_forward_debug": (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    "nt_gemm_bf16_debug": (i32, [vp, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, i32, vp, i32, vp]),
    "nt_composite_coarse": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
    "nt_get_density": (i32, [vp, i64, i32, vp, vp, vp, vp]),
    "nt_color_cum": (i32, [vp, i64, i32, vp, vp, vp, vp]),
    "nt_composite_fine": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]),
    "nt_composite_coarse_backward": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "nt_composite_fine_backward": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "nt_sample_pdf": (i32, [vp, i64, vp, vp, vp, vp, vp, vp]),
    "nt_sample_pdf_backward": (i32, [vp, i64, vp, vp, vp, vp, vp, vp]),
    "nt_check_status": (i32, [vp, vp]),
    "nt_ray_loss": (i32, [vp, i64, vp, vp, vp, vp, vp, vp, vp]),
    "nt_render_workspace_bytes": (sz, [vp, i32, i64, i32]),
    "nt_render_forward": (i32, [vp, i32, i64, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, sz, i32, vp]),
    "nt_render_backward": (i32, [vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "nt_adam_step_allreduce": (i32, [vp, i64, vp, C.POINTER(vp), i32, vp, vp, f32, f32, f32, f32, i64, f32, vp, vp]),
    "nt_adam_step": (i32, [vp, i64, vp, vp, vp, vp, f32, f32, f32, f32, i64, f32, vp]),
}

_lib = None


class NerfTinyError(RuntimeError):
    pass


class ResampleRangeError(NerfTinyError):
    """The reference calls exit(0) here (nerf.py:251-253); the drop-in raises instead."""


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfTinyError(
            f"{LIB_PATH} not found: build it with `python -m nerf_tiny_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == NT_OK:
        return
    msg = load().nt_last_error().decode("utf-8", "replace")
    if rc == NT_ERR_RANGE:
        raise ResampleRangeError(msg)
    raise NerfTinyError(f"libnerftiny error {rc}: {msg}")


def layer_table():
    arr = (LayerDesc * N_LAYERS)()
    check(load().nt_layer_table(arr))
    return [(d.out_features, d.in_features, d.weight_offset, d.bias_offset) for d in arr]

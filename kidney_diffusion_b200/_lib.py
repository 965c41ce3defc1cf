"""ctypes binding of libkidney_b200.so (the C ABI declared in include/kidney_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or the device is not a B200 the product path
raises.  ``build.build_library()`` (nvcc, sm_100a) produces the library in-tree.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_long, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkidney_b200.so")


class KdError(RuntimeError):
    pass


class KdConvDesc(Structure):
    _fields_ = [
        ("mode", c_int), ("B", c_int), ("H", c_int), ("W", c_int), ("Ca", c_int), ("Cb", c_int), ("Cout", c_int),
        ("ksize", c_int), ("act", c_int), ("out_mode", c_int), ("out_f32", c_int), ("addend_f32", c_int),
    ]


class KdConvFusion(Structure):
    _fields_ = [("stats", c_void_p), ("logit_w", c_void_p), ("logit_parts", c_void_p), ("pre_coef", c_void_p), ("splitk_ws", c_void_p),
                ("splitk_ws_bytes", c_size_t)]


_P = c_void_p
_F = c_float
_I = c_int
_L = c_long

# name -> (restype, argtypes); mirrors include/kidney_b200.h one to one
SIGNATURES = {
    "kd_version": (c_int, []),
    "kd_last_error": (c_char_p, []),
    "kd_check_device": (c_int, []),
    "kd_set_conv_impl": (c_int, [_I]),
    "kd_conv_gemm": (c_int, [POINTER(KdConvDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "kd_conv_stats_layout": (c_int, [POINTER(KdConvDesc), POINTER(c_int)]),
    "kd_conv_splitk_workspace_bytes": (c_size_t, [POINTER(KdConvDesc)]),
    "kd_conv_gemm_fused": (c_int, [POINTER(KdConvDesc), _P, _P, _P, _P, _P, _P, _P, POINTER(KdConvFusion), _P]),
    "kd_linear_small": (c_int, [_P, _I, _I, _L, _P, _P, _P, _I, _L, _I, _I, _P]),
    "kd_sinu_emb": (c_int, [_P, _P, _I, _I, _P, _P]),
    "kd_gn_stats": (c_int, [_P, _I, _L, _I, _I, _I, _I, _P, _I, _P]),
    "kd_gn_finalize": (c_int, [_P, _I, _F, _P, _I, _F, _I, _I, c_double, _F, _P, _P]),
    "kd_gn_apply": (c_int, [_P, _P, _I, _L, _I, _I, _I, _I, _F, _P, _P, _P, _P, _L, _I, _I, _P]),
    "kd_gn_reduce_finalize": (c_int, [_P, _I, _I, _I, _I, _F, _P, _I, _I, _I, _I, _F, _I, _I, _I, c_double, _F, _P, _P, _P, _P, _P, _P, _L, _P, _P]),
    "kd_rowdot": (c_int, [_P, _P, _P, _P, _I, _L, _I, _P]),
    "kd_gca_pool": (c_int, [_P, _P, _I, _I, _L, _I, _I, _P, _P, _P]),
    "kd_gca_finalize": (c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "kd_gca_gate": (c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "kd_gate_residual": (c_int, [_P, _P, _P, _P, _P, _I, _L, _I, _P]),
    "kd_elementwise_blocks": (c_int, [_L, _I]),
    "kd_oct_stats": (c_int, [_P, _I, _L, _I, _P, _I, _P]),
    "kd_oct_reduce": (c_int, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "kd_oct_reduce_splits": (c_int, [_I, _I, _I]),
    "kd_gn_finalize_oct": (c_int, [_P, _I, _I, _F, _P, _I, _I, _F, _I, _I, _I, c_double, _F, _P, _P, _P, _P, _L, _P, _P]),
    "kd_layernorm_h16": (c_int, [_P, _P, _P, _P, _P, _L, _I, _F, _P]),
    "kd_layernorm_f32": (c_int, [_P, _P, _P, _P, _L, _I, _F, _P]),
    "kd_kv_assemble": (c_int, [_P, _L, _I, _P, _I, _P, _P, _I, _I, _P]),
    "kd_attn_mqa": (c_int, [_P, _L, _P, _P, _I, _I, _I, _I, _F, _P]),
    "kd_attn_vt_elems": (c_int, [_I, _I]),
    "kd_attn_mqa_tc": (c_int, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "kd_attn_cross": (c_int, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "kd_attn_small_f32": (c_int, [_P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "kd_axpby": (c_int, [_P, _P, _F, _F, _P, _L, _P]),
    "kd_init_conv_kp": (c_int, [_I, _I]),
    "kd_init_conv": (c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "kd_im2col_nchw": (c_int, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "kd_final_conv_pack_elems": (c_long, [_I]),
    "kd_final_conv_pack": (c_int, [_P, _I, _I, _I, _P, _P]),
    "kd_final_conv": (c_int, [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "kd_dynthresh_workspace_bytes": (c_size_t, [_I]),
    "kd_dynthresh": (c_int, [_P, _P, _I, _L, _I, _F, _F, _L, _L, _F, _P, c_size_t, _P, _P]),
    "kd_ddpm_step": (c_int, [_P, _P, _P, _P, _P, _P, _I, _L, _I, _F, _F, _F, _F, _F, _F, _P, _F, _F, _F, _P]),
    "kd_inpaint_blend": (c_int, [_P, _P, _P, _P, _F, _F, _I, _I, _L, _P]),
    "kd_finalize_image": (c_int, [_P, _P, _P, _I, _I, _L, _P]),
    "kd_q_sample": (c_int, [_P, _P, _F, _F, _P, _L, _P]),
    "kd_randn": (c_int, [_P, _L, c_uint64, c_uint64, _P]),
    "kd_border_pack": (c_int, [_P, _P, _P, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _P]),
    "kd_count_saturated": (c_int, [_P, _L, _P, _P]),
    "kd_dwconv3x3": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "kd_linattn_blocks": (c_int, [_I]),
    "kd_linattn_workspace_bytes": (c_size_t, [_I, _I, _I]),
    "kd_linattn_context": (c_int, [_P, _L, _I, _I, _I, _I, _I, _P, _I, _P, c_size_t, _P, _P]),
    "kd_linattn_apply": (c_int, [_P, _L, _I, _P, _P, _I, _I, _I, _F, _I, _P]),
    "kd_conv_f32": (c_int, [_P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "kd_gn_chunks_f32": (c_int, [_L]),
    "kd_gn_stats_f32": (c_int, [_P, _I, _P, _I, _F, _I, _L, _I, _P, _P]),
    "kd_gn_finalize_f32": (c_int, [_P, _I, _L, _I, _I, _F, _P, _P]),
    "kd_gn_apply_f32": (c_int, [_P, _P, _I, _L, _I, _I, _I, _I, _F, _P, _P, _P, _P, _L, _I, _I, _P]),
    "kd_rowdot_f32": (c_int, [_P, _P, _P, _P, _L, _I, _P]),
    "kd_softmax_pool_f32": (c_int, [_P, _P, _I, _L, _I, _P, _P, _P, _P]),
    "kd_gate_residual_f32": (c_int, [_P, _P, _P, _P, _I, _L, _I, _P]),
    "kd_attn_f32": (c_int, [_P, _L, _P, _L, _L, _I, _P, _L, _L, _I, _P, _I, _I, _I, _I, _F, _P]),
    "kd_dwconv3x3_f32": (c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "kd_linattn_f32": (c_int, [_P, _L, _P, _P, _L, _L, _I, _I, _I, _I, _F, _I, _P, _P, _P]),
    "kd_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "kd_peer_free": (c_int, [_P]),
    "kd_peer_export": (c_int, [_P, _P]),
    "kd_peer_open": (c_int, [_P, POINTER(c_void_p)]),
    "kd_peer_close": (c_int, [_P]),
    "kd_strip_push": (c_int, [_P, _L, _L, _I, _I, _I, _P, _P, c_uint32, _P]),
    "kd_flag_wait": (c_int, [_P, c_uint32, c_double, _P, _P]),
    "kd_cond_gather": (c_int, [_P, _I, _P, _I, _I, _I, _I, _I, _F, _I, _I, _P]),
    "kd_canvas_fill": (c_int, [_P, _I, _P, _I, _P, _I, _I, _I, _P]),
    "kd_patch_paste": (c_int, [_P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol; raises KdError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KdError(
            f"{LIB_PATH} not found: build it with `python -m kidney_diffusion_b200.build` "
            "(hand-written CUDA for sm_100a; there is no CPU / PyTorch fallback path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            raise KdError(f"symbol {name} missing from {LIB_PATH}")
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().kd_last_error().decode("utf-8", "replace")
        raise KdError(f"{what} failed with status {rc}: {msg}")


_device_ok = False


def require_b200() -> None:
    """Fail loudly unless the current CUDA device is an sm_100 part."""
    global _device_ok
    if _device_ok:
        return
    check(load().kd_check_device(), "kd_check_device")
    _device_ok = True

"""ctypes binding of libmvster_b200.so (the C ABI declared in include/mvster_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the
caller gets an exception.  Pointers are raw device addresses (``tensor.data_ptr()``); the
stream is torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# MVSTER_LIB_PATH: an alternative build of the same library (A/B of compile-time switches); default = the in-tree build
LIB_PATH = Path(os.environ.get("MVSTER_LIB_PATH") or Path(__file__).resolve().parent / "lib" / "libmvster_b200.so")

_p = C.c_void_p
_i = C.c_int
_f = C.c_float

# name -> (restype, argtypes); one entry per function declared in include/mvster_b200.h
SIGNATURES = {
    "mvster_version": (_i, []),
    "mvster_last_error": (C.c_char_p, []),
    "mvster_launch_count": (C.c_uint64, []),
    "mvster_hypo_init_inverse_f32": (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    "mvster_hypo_schedule_inverse_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "mvster_hypo_init_linear_f32": (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    "mvster_hypo_schedule_linear_f32": (_i, [_p, _p, _i, _f, _p, _i, _i, _i, _i, _p]),
    "mvster_reg3d_num_layers": (_i, [_i]),
    "mvster_reg3d_layer_info": (_i, [_i, _i, _i, C.POINTER(C.c_int64)]),
    "mvster_reg3d_blob_floats": (C.c_size_t, [_i, _i]),
    "mvster_reg3d_workspace_floats": (C.c_size_t, [_i, _i, _i, _i]),
    "mvster_reg3d_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_pose_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "mvster_et_fuse_f32": (_i, [_p, C.POINTER(_p), _i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "mvster_et_fuse_bf16": (_i, [_p, C.POINTER(_p), _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "mvster_cast_bf16": (_i, [_p, _p, C.c_longlong, _p]),
    "mvster_et_normalize_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "mvster_et_last_kernel": (C.c_char_p, []),
    "mvster_et_fuse_bwd_f32": (_i, [_p, C.POINTER(_p), _i, _p, _p, _p, _p, _p, _p, C.POINTER(_p), _i, _i, _i, _i, _i, _i, _i, _i, _f, _p]),
    "mvster_conv3d_ndhwc_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_reg2d_blob_floats": (C.c_size_t, [_i]),
    "mvster_reg2d_workspace_floats": (C.c_size_t, [_i, _i, _i, _i]),
    "mvster_reg2d_layer_info": (_i, [_i, _i, C.POINTER(C.c_int64)]),
    "mvster_reg2d_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mvster_reg2d_tc_blob_floats": (C.c_size_t, []),
    "mvster_reg2d_tc_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_conv3d_tc2_supported": (_i, [_i, _i, _i, _i, _i]),
    "mvster_conv3d_tc2_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_reg2d_tc3_blob_bytes": (C.c_size_t, [_i]),
    "mvster_reg2d_tc3_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "mvster_reg2d_tc3_ex_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_deconv_tc3_supported": (_i, [_i, _i, _i]),
    "mvster_deconv_tc3_packed_bytes": (C.c_size_t, [_i, _i, _i]),
    "mvster_deconv_tc3_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_set_sm_budget": (None, [_i]),
    "mvster_tc3_set_overflow_flag": (None, [_p]),
    "mvster_tc3_overflow_flag": (_p, []),
    "mvster_conv_tc3_supported": (_i, [_i, _i, _i, _i, _i]),
    "mvster_conv_tc3_plan": (_i, [_i, _i, _i, _i, C.POINTER(_i), _i]),
    "mvster_conv_tc3_packed_bytes": (C.c_size_t, [_i, _i, _i, _i, _i]),
    "mvster_conv_tc3_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_conv_tc3_scaled_f32": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_deconv_tc3_scaled_f32": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_reg2d_bf16": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_conv_tc3_pb16": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_deconv_tc3_pb16": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_pointwise_tc3_blocks_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, C.c_longlong, _p]),
    "mvster_pointwise_tc3_blocks_ex_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, C.c_longlong, _i, _p]),
    "mvster_conv2d_nhwc_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_conv_first_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "mvster_fpn_merge_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "mvster_pointwise_tc2_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "mvster_fpn_out4_gather_f32": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _p]),
    "mvster_sinkhorn_f32": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "mvster_geo_consistency_f32": (_i, [_p, _p, C.POINTER(C.c_double), _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _f, _p]),
    "mvster_head_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "mvster_head_ex_f32": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "mvster_upsample_bilinear_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "mvster_nchw_to_nhwc_f32": (_i, [_p, _p, _i, _i, _i, _i, _p]),
}

_lib = None


class MvsterLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load (once) and type the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MvsterLibraryError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
            "`python -m mvster_b200.build` (or __graft_entry__.build()); there is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mvster_last_error().decode(errors="replace")
        raise MvsterLibraryError(f"{what} failed with code {rc}: {msg}")


def launch_count() -> int:
    return int(load().mvster_launch_count())

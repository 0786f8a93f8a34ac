"""ctypes binding of libm4d.so (include/m4d.h) - the only way the Python host reaches the GPU.

There is no CPU fallback: importing this module without the built library, or calling an op without a CUDA
device, raises.  Tensors are torch CUDA tensors used purely as device buffers; every call is enqueued on torch's
current stream so the whole frame can be captured in a CUDA graph.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# M4D_LIB: another build of the same library (A/B timing of kernel variants on the GPU box); default = the in-tree build
LIB_PATH = os.environ.get("M4D_LIB") or os.path.join(_HERE, "libm4d.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m m4depth_b200._build` (nvcc, sm_100a). "
        "m4depth_b200 has no CPU or PyTorch fallback path.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_i64 = C.c_int64

SIGNATURES = {
    "m4d_abi_version": (C.c_int, []),
    "m4d_last_error_string": (C.c_char_p, []),
    "m4d_launch_count": (C.c_uint64, []),
    "m4d_backproject_fwd": (_i, [_p, _p, C.POINTER(C.c_int32), _p, _p, _p]),
    "m4d_backproject_bwd": (_i, [_p, _p, _p, C.POINTER(C.c_int32), _p, _p, _p]),
    "m4d_dense_image_warp": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "m4d_get_rot_mat": (_i, [_p, _i, _i, _p, _p]),
    "m4d_prev_d2para": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _p, _p]),
    "m4d_parallax2depth": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _p, _p]),
    "m4d_depth2parallax": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _i, _p, _p]),
    "m4d_pscv_fused_fwd": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i,
                                _p, _i, _p, _i, _p, _i, _f, _p, _p]),
    "m4d_pscv_fused_fwd_ex": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i,
                                   _p, _i, _p, _i, _p, _i, _f, _p, _i, _p]),
    "m4d_debug_div_check": (_i, [_p, _p, _i, _p, _p, _p, _p]),
    "m4d_pscv_fused_bwd": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p, _p]),
    "m4d_sncv_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "m4d_sncv_fwd_ex": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "m4d_group_l2norm": (_i, [_p, _i, _i, _i, _p, _p]),
    "m4d_domain_norm": (_i, [_p, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p]),
    "m4d_rgb_conv_dn": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _f, _p, _p, _p]),
    "m4d_rgb_conv_dn_hostw": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _f, _p, _p, _p]),
    "m4d_rgb_conv_stats_hostw": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _p]),
    "m4d_domain_norm_apply": (_i, [_p, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p]),
    "m4d_conv3x3_nhwc": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p]),
    "m4d_conv3x3_tc_packed_floats": (C.c_int64, [_i, _i]),
    "m4d_conv3x3_tc_pack": (_i, [_p, _i, _i, _p, _p]),
    "m4d_conv3x3_tc_fwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _f, _p, _i, _p]),
    "m4d_conv3x3_tc_packed_floats_s": (C.c_int64, [_i, _i, _i]),
    "m4d_conv3x3_tc_pack_s": (_i, [_p, _i, _i, _i, _p, _p]),
    "m4d_conv3x3_tc_fwd_s": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _i, _p]),
    "m4d_conv3x3_tc_fwd_ex": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p]),
    "m4d_debug_conv_profile": (_i, [_p]),
    "m4d_conv3x3_tc_packed_floats_p": (C.c_int64, [_i, _i, _i, _i]),
    "m4d_conv3x3_tc_pack_p": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "m4d_conv3x3_tc_fwd_p": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p]),
    "m4d_resize_bilinear_legacy": (_i, [_p, _i, _i, _i, _i, _i, _i, _f, _p, _i, _p]),
    "m4d_resize_nearest": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "m4d_level_prologue": (_i, [_p, _p, _p, _i, _i, _p, _p, _i, _p, _p, _p, _i, _i, _i,
                                _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "m4d_level_epilogue": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p]),
    "m4d_camera_pyramid": (_i, [_p, _p, _i, _i, _p, _p, _p]),
    "m4d_fill": (_i, [_p, _i64, _f, _p]),
    "m4d_pad_shift": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "m4d_depth_metrics": (_i, [_p, _p, _i64, _f, _p, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

ABI_VERSION = lib.m4d_abi_version()

INTERP_GATHER, INTERP_BP, INTERP_BP_FMA = 0, 1, 2
INTERP_FLAG_GENERIC = 0x100     # OR into interp: shape-generic PSCV kernel instead of the specialised one
INTERP_FLAG_TILE = 0x200        # OR into interp: CTA-tile K=9 kernel instead of the warp-autonomous one
INTERP_FLAG_WARP = 0x400        # OR into interp: warp-autonomous LDG-gather kernel instead of the shared-memory staged one
INTERP_VARIANT_STAGED_8 = 0x1000      # OR into interp (tuning / tests): staged kernel, 16x8 tiles, phases in sequence (pscv9s_kernel)
INTERP_VARIANT_STAGED_4 = 0x2000      # staged kernel, 16x4 tiles
INTERP_VARIANT_STAGED_C16 = 0x4000    # c = 16, one group (level 1): the staged kernel with 32x8 tiles and 64-byte pixel rows


class M4DError(RuntimeError):
    pass


def last_error():
    return lib.m4d_last_error_string().decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        raise M4DError(f"libm4d error {rc}: {last_error()}")


def ptr(t):
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise M4DError("libm4d needs CUDA tensors: there is no CPU path")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib.m4d_launch_count())


def f32c(t, name="tensor"):
    """Check a tensor is a dense float32 CUDA tensor (no silent copies on the hot path)."""
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise M4DError(f"{name} must be a contiguous float32 CUDA tensor (got {t.dtype}, cuda={t.is_cuda}, "
                       f"contiguous={t.is_contiguous()})")
    return t

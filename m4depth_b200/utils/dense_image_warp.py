"""Mirror of the reference's ``utils/dense_image_warp.py`` on torch CUDA tensors, backed by libm4d.

* ``back_project(inputs, coords)``   = TF op ``BackProject`` (cuda_backproject/backproject_op.cc:32-35,53-92)
* ``dense_image_warp(image, flow)``  = utils/dense_image_warp.py:195-268, BackProject branch (:246-253): query =
  grid + flow, clipped to the image, bilinear sample.  (Values equal the python-gather branch to fp32 rounding.)
"""
import ctypes as C

import torch

from .. import _lib as L


def back_project(inputs, coords, return_index_grids=False):
    """inputs [B,H,W,F,C], coords [B,H,W,S,F,2] (x,y) -> [B,H,W,S,F,C]; zero outside the image / for NaN coords.

    With ``return_index_grids`` also returns the int32 tap grid [B,H,W,S,F,4] = x0,x1,y0,y1 (-1 outside).
    """
    L.f32c(inputs, "inputs"), L.f32c(coords, "coords")
    if inputs.dim() != 5 or coords.dim() != 6 or coords.shape[-1] != 2:
        raise L.M4DError("back_project: inputs must be [B,H,W,F,C] and coords [B,H,W,S,F,2]")
    B, H, W, Fd, Cc = inputs.shape
    S = coords.shape[3]
    if tuple(coords.shape[:3]) != (B, H, W) or coords.shape[4] != Fd:
        raise L.M4DError("back_project: coords / inputs shape mismatch")
    out = torch.empty((B, H, W, S, Fd, Cc), dtype=torch.float32, device=inputs.device)
    idx = torch.empty((B, H, W, S, Fd, 4), dtype=torch.int32, device=inputs.device) if return_index_grids else None
    dim = (C.c_int32 * 6)(B, H, W, S, Fd, Cc)
    L.check(L.lib.m4d_backproject_fwd(L.ptr(inputs), L.ptr(coords), dim, L.ptr(out), L.ptr(idx), L.stream()))
    return (out, idx) if return_index_grids else out


def back_project_grad(inputs, coords, grad):
    """Gradient op of ``back_project`` (TF op BackProjectGrad, registered at utils/dense_image_warp.py:46-52):
    (inputs [B,H,W,F,C], coords [B,H,W,S,F,2], grad [B,H,W,S,F,C]) -> (inputs_grad [B,H,W,F,C], coords_grad [B,H,W,S,F,2])."""
    L.f32c(inputs, "inputs"), L.f32c(coords, "coords"), L.f32c(grad, "grad")
    if inputs.dim() != 5 or coords.dim() != 6 or coords.shape[-1] != 2 or grad.dim() != 6:
        raise L.M4DError("back_project_grad: inputs must be [B,H,W,F,C], coords [B,H,W,S,F,2], grad [B,H,W,S,F,C]")
    B, H, W, Fd, Cc = inputs.shape
    S = coords.shape[3]
    if tuple(coords.shape[:3]) != (B, H, W) or coords.shape[4] != Fd or tuple(grad.shape) != (B, H, W, S, Fd, Cc):
        raise L.M4DError("back_project_grad: shape mismatch")
    inputs_grad = torch.empty_like(inputs)
    coords_grad = torch.empty_like(coords)
    dim = (C.c_int32 * 6)(B, H, W, S, Fd, Cc)
    L.check(L.lib.m4d_backproject_bwd(L.ptr(grad), L.ptr(inputs), L.ptr(coords), dim, L.ptr(inputs_grad), L.ptr(coords_grad),
                                      L.stream()))
    return inputs_grad, coords_grad


def dense_image_warp(image, flow, name='dense_image_warp'):
    """image [b,h,w,c], flow [b,h,w,2] (row, col) -> [b,h,w,c]; pixel (y,x) samples image at (y,x) + flow."""
    L.f32c(image, "image"), L.f32c(flow, "flow")
    b, h, w, c = image.shape
    if tuple(flow.shape) != (b, h, w, 2):
        raise L.M4DError("dense_image_warp: flow must be [b,h,w,2]")
    out = torch.empty_like(image)
    L.check(L.lib.m4d_dense_image_warp(L.ptr(image), L.ptr(flow), b, h, w, c, L.ptr(out), L.stream()))
    return out

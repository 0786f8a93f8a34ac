"""Mirror of the reference's ``utils/depth_operations.py`` (function names, argument order, return structure) on
torch CUDA tensors, NHWC fp32, backed by libm4d.  ``camera`` is ``{"f": [b,2], "c": [b,2]}`` like the reference.
"""
import torch

from .. import _lib as L


def _pose(rot, trans, camera):
    rot, trans = L.f32c(rot, "rot"), L.f32c(trans, "trans")
    f, c = L.f32c(camera["f"], "camera['f']"), L.f32c(camera["c"], "camera['c']")
    if rot.dim() != 2 or rot.shape[1] not in (3, 4):
        raise ValueError('Rotation must be expressed as a small angle (x,y,z) or a quaternion (w,x,y,z)')
    return rot, trans, f, c


def get_rot_mat(rot):
    """[b,4] quaternion (w,x,y,z) or [b,3] small angles -> [b,3,3]  (depth_operations.py:18-53)."""
    rot = L.f32c(rot, "rot")
    if rot.dim() != 2 or rot.shape[1] not in (3, 4):
        raise ValueError('Rotation must be expressed as a small angle (x,y,z) or a quaternion (w,x,y,z)')
    out = torch.empty((rot.shape[0], 3, 3), dtype=torch.float32, device=rot.device)
    L.check(L.lib.m4d_get_rot_mat(L.ptr(rot), rot.shape[0], rot.shape[1], L.ptr(out), L.stream()))
    return out


def get_coords_2d(map, camera):
    """-> (coords2d [b,h,w,3,1], mesh [b,h,w,2]) (:56-68).  Host-composed helper (not on the fused hot path: the
    kernels evaluate these per pixel in registers); kept for API parity."""
    b, h, w = map.shape[0:3]
    dev = map.device
    ys = torch.arange(h, dtype=torch.float32, device=dev) + 0.5
    xs = torch.arange(w, dtype=torch.float32, device=dev) + 0.5
    mesh = torch.stack((xs.view(1, 1, w).expand(1, h, w), ys.view(1, h, 1).expand(1, h, w)), dim=-1)
    mesh = mesh - camera["c"].view(b, 1, 1, 2)
    n = mesh / camera["f"].view(b, 1, 1, 2)
    coords = torch.cat((n, torch.ones(b, h, w, 1, dtype=torch.float32, device=dev)), dim=-1)
    return coords.unsqueeze(-1), mesh


def _geo(fn, x, rot, trans, camera):
    x = L.f32c(x, "map")
    rot, trans, f, c = _pose(rot, trans, camera)
    b, h, w = x.shape[0:3]
    out = torch.empty_like(x)
    L.check(fn(L.ptr(x), L.ptr(rot), rot.shape[1], L.ptr(trans), L.ptr(f), L.ptr(c), b, h, w, L.ptr(out), L.stream()))
    return out


def parallax2depth(disp, rot, trans, camera):
    """[b,h,w,1] parallax -> depth (:140-166)."""
    return _geo(L.lib.m4d_parallax2depth, disp, rot, trans, camera)


def depth2parallax(depth, rot, trans, camera):
    """[b,h,w,1] depth -> parallax (:168-194)."""
    return _geo(L.lib.m4d_depth2parallax, depth, rot, trans, camera)


def prev_d2para(prev_d, rot, trans, camera):
    """previous-frame depth -> parallax seen from the current frame, rotation ignored (:196-215)."""
    return _geo(L.lib.m4d_prev_d2para, prev_d, rot, trans, camera)


def tile_in_batch(map, nbre_copies):
    """[b,...] -> [n*b,...] (:217-221).  The fused PSCV never materialises this; API parity only."""
    return map.unsqueeze(0).expand(nbre_copies, *map.shape).reshape(-1, *map.shape[1:])


def get_parallax_sweeping_cv(c1, c2, disp_prev_t, disp, rot, trans, camera, search_range, nbre_cuts=1,
                             interp=L.INTERP_GATHER, return_index_grids=False):
    """PSCV (:223-281) -> (cv [b,h,w,cuts*(2r+1)], prev_disp [b,h,w,2r+1]) in one fused kernel.

    ``interp`` picks the bilinear convention (include/m4d.h): the reference's python-gather path (default, what
    TF-CPU runs), the BackProject path, or the BackProject path with nvcc's FMA contraction.
    With ``return_index_grids`` a third result holds the BackProject tap grids int32 [b,h,w,2r+1,4] = x0,x1,y0,y1.
    """
    for t, n in ((c1, "c1"), (c2, "c2"), (disp_prev_t, "disp_prev_t"), (disp, "disp")):
        L.f32c(t, n)
    rot, trans, f, c = _pose(rot, trans, camera)
    b, h, w, ch = c1.shape
    K = 2 * search_range + 1
    cv = torch.empty((b, h, w, nbre_cuts * K), dtype=torch.float32, device=c1.device)
    pd = torch.empty((b, h, w, K), dtype=torch.float32, device=c1.device)
    idx = torch.empty((b, h, w, K, 4), dtype=torch.int32, device=c1.device) if return_index_grids else None
    L.check(L.lib.m4d_pscv_fused_fwd_ex(
        L.ptr(c1), L.ptr(c2), L.ptr(disp_prev_t), L.ptr(disp), L.ptr(rot), rot.shape[1], L.ptr(trans), L.ptr(f), L.ptr(c),
        b, h, w, ch, nbre_cuts, search_range, L.ptr(cv), nbre_cuts * K, L.ptr(pd), K, None, 0, 1.0, L.ptr(idx),
        interp, L.stream()))
    return (cv, pd, idx) if return_index_grids else (cv, pd)


def get_parallax_sweeping_cv_grad(c1, c2, disp_prev_t, disp, rot, trans, camera, search_range, d_cv, d_prev_disp=None, nbre_cuts=1):
    """Gradient of ``get_parallax_sweeping_cv`` (gather convention) for given output gradients:
    -> (d_c1, d_c2, d_disp_prev_t, d_disp).  What TensorFlow's autodiff produces for :223-281 in train_step."""
    for t, n in ((c1, "c1"), (c2, "c2"), (disp_prev_t, "disp_prev_t"), (disp, "disp"), (d_cv, "d_cv")):
        L.f32c(t, n)
    if d_prev_disp is not None:
        L.f32c(d_prev_disp, "d_prev_disp")
    rot, trans, f, c = _pose(rot, trans, camera)
    b, h, w, ch = c1.shape
    K = 2 * search_range + 1
    if tuple(d_cv.shape) != (b, h, w, nbre_cuts * K) or (d_prev_disp is not None and tuple(d_prev_disp.shape) != (b, h, w, K)):
        raise L.M4DError("get_parallax_sweeping_cv_grad: gradient shapes must match the forward outputs")
    d_c1, d_c2 = torch.empty_like(c1), torch.empty_like(c2)
    d_pt, d_pl = torch.empty_like(disp_prev_t), torch.empty_like(disp)
    L.check(L.lib.m4d_pscv_fused_bwd(
        L.ptr(c1), L.ptr(c2), L.ptr(disp_prev_t), L.ptr(disp), L.ptr(rot), rot.shape[1], L.ptr(trans), L.ptr(f), L.ptr(c),
        b, h, w, ch, nbre_cuts, search_range, L.ptr(d_cv), nbre_cuts * K, L.ptr(d_prev_disp), K,
        L.ptr(d_c1), L.ptr(d_c2), L.ptr(d_pt), L.ptr(d_pl), L.stream()))
    return d_c1, d_c2, d_pt, d_pl


def cost_volume(c1, c2, search_range, name="cost_volume", dilation_rate=1, nbre_cuts=1):
    """SNCV (:283-313): [b,h,w,(2r+1)^2*cuts], channel = (dy*(2r+1)+dx)*cuts + cut, leaky_relu(0.1) applied."""
    L.f32c(c1, "c1"), L.f32c(c2, "c2")
    if dilation_rate != 1:
        raise L.M4DError("cost_volume: only dilation_rate=1 is built (the network never uses another value)")
    b, h, w, ch = c1.shape
    n = 2 * search_range + 1
    out = torch.empty((b, h, w, n * n * nbre_cuts), dtype=torch.float32, device=c1.device)
    L.check(L.lib.m4d_sncv_fwd(L.ptr(c1), L.ptr(c2), b, h, w, ch, nbre_cuts, search_range, L.ptr(out),
                               n * n * nbre_cuts, L.stream()))
    return out

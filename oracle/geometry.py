"""Oracle: camera geometry of the parallax path (test infrastructure, see oracle/__init__.py).

Restates ``/root/reference/utils/depth_operations.py``:
  get_rot_mat 18-53 · get_coords_2d 56-68 · parallax2depth 140-166 · depth2parallax 168-194 ·
  prev_d2para 196-215 · the flow construction of get_parallax_sweeping_cv 239-265.

All tensors are torch CPU float32, NHWC.  Every arithmetic step is a separate torch op so it is
rounded to fp32 exactly once, like the reference's TF graph (no FMA contraction).  The one place
TF does not define the order - the 3-term sum inside the batched ``rot_mat @ coords2d`` matmul
(:154, :182, :246) - is fixed here as ``(R0*nx + R1*ny) + R2*1`` and the CUDA kernels follow the
same order, so index grids are bit-exact between oracle and kernels.
"""
import torch

from ._ieee import sqrt as ieee_sqrt

F32 = torch.float32


def get_rot_mat(rot):
    """rot [b,4] quaternion (w,x,y,z) or [b,3] small angles -> [b,3,3]  (depth_operations.py:18-53)."""
    rot = rot.to(F32)
    b, c = rot.shape
    if c == 3:
        one = torch.ones(b, dtype=F32)
        m = torch.stack((one, -rot[:, 2], rot[:, 1],
                         rot[:, 2], one, -rot[:, 0],
                         -rot[:, 1], rot[:, 0], one), dim=-1)
        return m.reshape(b, 3, 3)
    if c != 4:
        raise ValueError('Rotation must be expressed as a small angle (x,y,z) or a quaternion (w,x,y,z)')
    w, x, y, z = rot.unbind(-1)
    tx = 2.0 * x
    ty = 2.0 * y
    tz = 2.0 * z
    twx = tx * w
    twy = ty * w
    twz = tz * w
    txx = tx * x
    txy = ty * x
    txz = tz * x
    tyy = ty * y
    tyz = tz * y
    tzz = tz * z
    m = torch.stack((1.0 - (tyy + tzz), txy - twz, txz + twy,
                     txy + twz, 1.0 - (txx + tzz), tyz - twx,
                     txz - twy, tyz + twx, 1.0 - (txx + tyy)), dim=-1)
    return m.reshape(b, 3, 3)


def get_coords_2d(fmap, camera):
    """-> (coords [b,h,w,3] = ((x+.5-cx)/fx, (y+.5-cy)/fy, 1), mesh [b,h,w,2])  (:56-68).

    The reference returns coords with a trailing singleton axis; callers here add it if needed.
    """
    b, h, w = fmap.shape[0:3]
    ys = torch.arange(0, h, dtype=F32) + 0.5
    xs = torch.arange(0, w, dtype=F32) + 0.5
    gx = xs.view(1, 1, w).expand(1, h, w)
    gy = ys.view(1, h, 1).expand(1, h, w)
    mesh = torch.stack((gx, gy), dim=-1) - camera["c"].to(F32).view(b, 1, 1, 2)
    n = mesh / camera["f"].to(F32).view(b, 1, 1, 2)
    coords = torch.cat((n, torch.ones(b, h, w, 1, dtype=F32)), dim=-1)
    return coords, mesh


def _rotate(R, coords):
    """r = R @ n with the documented summation order; R [b,3,3], coords [b,h,w,3] -> 3x [b,h,w]."""
    nx, ny = coords[..., 0], coords[..., 1]
    out = []
    for i in range(3):
        r0 = R[:, i, 0].view(-1, 1, 1)
        r1 = R[:, i, 1].view(-1, 1, 1)
        r2 = R[:, i, 2].view(-1, 1, 1)
        out.append((r0 * nx + r1 * ny) + r2 * 1.0)
    return out


def _epipolar_terms(fmap, rot, trans, camera):
    """Shared pieces of :140-194 and :239-259.

    Returns dict of [b,h,w] tensors: alpha, px, py (projected coords), dx, dy (delta), s (sqrt),
    sx, sy (start coords = (mesh/f)*f), stz (scaled t_z, [b,1,1]).
    """
    b = fmap.shape[0]
    coords, _ = get_coords_2d(fmap, camera)
    R = get_rot_mat(rot)
    f = camera["f"].to(F32)
    fx = f[:, 0].view(b, 1, 1)
    fy = f[:, 1].view(b, 1, 1)
    t = trans.to(F32)
    rx, ry, rz = _rotate(R, coords)
    alpha = rz
    px = (rx * fx) / alpha
    py = (ry * fy) / alpha
    stx = (t[:, 0] * f[:, 0]).view(b, 1, 1)
    sty = (t[:, 1] * f[:, 1]).view(b, 1, 1)
    stz = (t[:, 2] * 1.0).view(b, 1, 1)
    dx = stx - stz * px
    dy = sty - stz * py
    s = ieee_sqrt(dx * dx + dy * dy)
    sx = coords[..., 0] * fx
    sy = coords[..., 1] * fy
    return dict(alpha=alpha, px=px, py=py, dx=dx, dy=dy, s=s, sx=sx, sy=sy, stz=stz,
                stx=stx, sty=sty)


def parallax2depth(disp, rot, trans, camera):
    """disp [b,h,w,1] -> depth [b,h,w,1] = (s/disp - tz)/alpha  (:140-166)."""
    g = _epipolar_terms(disp, rot, trans, camera)
    depth = (g["s"] / disp[..., 0] - g["stz"]) / g["alpha"]
    return depth.unsqueeze(-1)


def depth2parallax(depth, rot, trans, camera):
    """depth [b,h,w,1] -> parallax [b,h,w,1] = s/(depth*alpha + tz)  (:168-194)."""
    g = _epipolar_terms(depth, rot, trans, camera)
    disp = g["s"] / (depth[..., 0] * g["alpha"] + g["stz"])
    return disp.unsqueeze(-1)


def prev_d2para(prev_d, rot, trans, camera):
    """Previous-frame depth -> parallax seen from the current frame, rotation ignored (:196-215)."""
    b = prev_d.shape[0]
    coords, _ = get_coords_2d(prev_d, camera)
    f = camera["f"].to(F32)
    t = trans.to(F32)
    fx = f[:, 0].view(b, 1, 1)
    fy = f[:, 1].view(b, 1, 1)
    tz = t[:, 2].view(b, 1, 1)
    sx = coords[..., 0] * fx
    sy = coords[..., 1] * fy
    stx = (t[:, 0] * f[:, 0]).view(b, 1, 1)
    sty = (t[:, 1] * f[:, 1]).view(b, 1, 1)
    den = prev_d[..., 0] - tz
    vx = (stx - tz * sx) / den
    vy = (sty - tz * sy) / den
    # tf.norm(axis) = sqrt(reduce_sum(x*x))
    return ieee_sqrt(vx * vx + vy * vy).unsqueeze(-1)


def pscv_query_points(disp, rot, trans, camera, search_range):
    """Query points of the parallax sweep (:229-265 + dense_image_warp.py:238-244).

    disp [b,h,w,1] (= para_prev_l).  Returns (qy, qx) each [K,b,h,w] with K = 2*search_range+1,
    *before* any clipping: q = grid + flow, flow = reverse(proj + delta - start).
    """
    b, h, w = disp.shape[0:3]
    g = _epipolar_terms(disp, rot, trans, camera)
    K = 2 * search_range + 1
    ks = torch.arange(-search_range, search_range + 1, dtype=F32).view(K, 1, 1, 1)
    rho = torch.clamp(disp[..., 0].unsqueeze(0) + ks, 1e-6, 1e6)
    div = g["s"].unsqueeze(0) / rho
    ex = g["dx"].unsqueeze(0) / div
    ey = g["dy"].unsqueeze(0) / div
    flow_x = (g["px"].unsqueeze(0) + ex) - g["sx"].unsqueeze(0)
    flow_y = (g["py"].unsqueeze(0) + ey) - g["sy"].unsqueeze(0)
    gy = torch.arange(h, dtype=F32).view(1, 1, h, 1)
    gx = torch.arange(w, dtype=F32).view(1, 1, 1, w)
    return gy + flow_y, gx + flow_x

"""Oracle: bilinear warp, both reference code paths (test infrastructure, see oracle/__init__.py).

* ``interpolate_bilinear``  restates ``/root/reference/utils/dense_image_warp.py:61-192``
  (4 gathers + 3 lerps, floor clamped to [0, size-2], ceil = floor+1, alpha clamped to [0,1]).
* ``back_project``          restates ``/root/reference/cuda_backproject/backproject_op_gpu.cu.cc:19-79``
  (floor/ceil taps, 4 weights, zero outside [0,W-1]x[0,H-1] or NaN) for the general
  ``inputs [B,H,W,F,C]``, ``coords [B,H,W,S,F,2]`` (x,y) signature of ``backproject_op.cc:32-35,53-92``.
* ``dense_image_warp``      restates ``dense_image_warp.py:195-268`` and dispatches to either.

A plain-C restatement of the BackProject kernel bodies (forward and gradient) lives in ``oracle/m4d_oracle.c``
(built by ``oracle/Makefile``) and is cross-checked against ``back_project`` / ``back_project_grad`` in
``tests/test_oracle_golden.py``.
"""
import torch

F32 = torch.float32


def interpolate_bilinear(grid, query_points):
    """grid [b,h,w,c], query_points [b,n,2] in (row, col) -> [b,n,c]  (dense_image_warp.py:61-192)."""
    b, h, w, c = grid.shape
    n = query_points.shape[1]
    alphas, floors, ceils = [], [], []
    for dim, size in ((0, h), (1, w)):
        q = query_points[..., dim]
        fl = torch.minimum(torch.maximum(torch.zeros((), dtype=F32), torch.floor(q)),
                           torch.tensor(float(size - 2), dtype=F32))
        ifl = fl.to(torch.int32)
        floors.append(ifl)
        ceils.append(ifl + 1)
        a = torch.clamp(q - fl, 0.0, 1.0)
        alphas.append(a.unsqueeze(-1))
    flat = grid.reshape(b * h * w, c)
    boff = (torch.arange(b, dtype=torch.int64) * (h * w)).view(b, 1)

    def gather(yc, xc):
        lin = boff + yc.to(torch.int64) * w + xc.to(torch.int64)
        return flat[lin.reshape(-1)].reshape(b, n, c)

    tl = gather(floors[0], floors[1])
    tr = gather(floors[0], ceils[1])
    bl = gather(ceils[0], floors[1])
    br = gather(ceils[0], ceils[1])
    top = alphas[1] * (tr - tl) + tl
    bot = alphas[1] * (br - bl) + bl
    return alphas[0] * (bot - top) + top


def back_project_index_grids(coords, H, W):
    """Integer tap grids of the BackProject kernel (backproject_op_gpu.cu.cc:44-51).

    coords [...,2] (x,y).  Returns int32 tensors x0,x1,y0,y1 and bool ``inside``; entries where
    ``inside`` is False are set to -1 (the kernel computes nothing there).
    """
    x = coords[..., 0]
    y = coords[..., 1]
    inside = (x >= 0) & (y >= 0) & (x <= W - 1) & (y <= H - 1)
    xs = torch.where(inside, x, torch.zeros_like(x))
    ys = torch.where(inside, y, torch.zeros_like(y))
    neg = torch.full_like(xs, -1).to(torch.int32)
    x0 = torch.where(inside, torch.floor(xs).to(torch.int32), neg)
    x1 = torch.where(inside, torch.ceil(xs).to(torch.int32), neg)
    y0 = torch.where(inside, torch.floor(ys).to(torch.int32), neg)
    y1 = torch.where(inside, torch.ceil(ys).to(torch.int32), neg)
    return x0, x1, y0, y1, inside


def back_project(inputs, coords):
    """BackProject forward.  inputs [B,H,W,F,C], coords [B,H,W,S,F,2] (x,y) -> [B,H,W,S,F,C].

    out = I00*w00 + I01*w01 + I10*w10 + I11*w11 summed left to right with separately rounded
    products (the reference binary may contract these into FMAs; difference <= 1 ulp of the sum),
    zero where the coordinate is outside the image or NaN (memset at :91 + guard at :47).
    """
    B, H, W, Fd, C = inputs.shape
    S = coords.shape[3]
    x0, x1, y0, y1, inside = back_project_index_grids(coords, H, W)
    x = torch.where(inside, coords[..., 0], torch.zeros((), dtype=F32))
    y = torch.where(inside, coords[..., 1], torch.zeros((), dtype=F32))
    x0c, x1c = x0.clamp(min=0).to(torch.int64), x1.clamp(min=0).to(torch.int64)
    y0c, y1c = y0.clamp(min=0).to(torch.int64), y1.clamp(min=0).to(torch.int64)
    dx = x - x0c.to(F32)
    dy = y - y0c.to(F32)
    w00 = (1 - dy) * (1 - dx)
    w01 = (1 - dy) * dx
    w10 = dy * (1 - dx)
    w11 = dy * dx
    flat = inputs.reshape(B * H * W * Fd, C)
    bidx = torch.arange(B, dtype=torch.int64).view(B, 1, 1, 1, 1)
    fidx = torch.arange(Fd, dtype=torch.int64).view(1, 1, 1, 1, Fd)

    def tap(yy, xx):
        lin = ((bidx * H + yy) * W + xx) * Fd + fidx
        return flat[lin.reshape(-1)].reshape(B, H, W, S, Fd, C)

    out = (tap(y0c, x0c) * w00.unsqueeze(-1) + tap(y0c, x1c) * w01.unsqueeze(-1)
           + tap(y1c, x0c) * w10.unsqueeze(-1) + tap(y1c, x1c) * w11.unsqueeze(-1))
    return torch.where(inside.unsqueeze(-1), out, torch.zeros((), dtype=F32))


def back_project_grad(inputs, coords, grad):
    """BackProjectGrad (backproject_op_gpu.cu.cc:108-196; registered as the gradient of BackProject at
    utils/dense_image_warp.py:46-52).  inputs [B,H,W,F,C], coords [B,H,W,S,F,2] (x,y), grad [B,H,W,S,F,C]
    -> (inputs_grad [B,H,W,F,C], coords_grad [B,H,W,S,F,2]).

    inputs_grad: grad * w_ij scatter-added onto the four taps (:178-181; the reference's atomicAdd order is not fixed,
    here index_add in sample order).  coords_grad (:183-184, summed over channels in order):
        d/dx = sum_c g * ((1-dy)*(I01 - I00) + dy*(I11 - I10)),   d/dy = sum_c g * ((1-dx)*(I10 - I00) + dx*(I11 - I01));
    both zero where the coordinate is outside [0,W-1]x[0,H-1] or NaN (memsets at :209-210 + guard at :130)."""
    B, H, W, Fd, C = inputs.shape
    S = coords.shape[3]
    x0, x1, y0, y1, inside = back_project_index_grids(coords, H, W)
    x = torch.where(inside, coords[..., 0], torch.zeros((), dtype=F32))
    y = torch.where(inside, coords[..., 1], torch.zeros((), dtype=F32))
    x0c, x1c = x0.clamp(min=0).to(torch.int64), x1.clamp(min=0).to(torch.int64)
    y0c, y1c = y0.clamp(min=0).to(torch.int64), y1.clamp(min=0).to(torch.int64)
    dx = (x - x0c.to(F32)).unsqueeze(-1)
    dy = (y - y0c.to(F32)).unsqueeze(-1)
    wx0, wx1, wy0, wy1 = 1 - dx, dx, 1 - dy, dy
    flat = inputs.reshape(B * H * W * Fd, C)
    bidx = torch.arange(B, dtype=torch.int64).view(B, 1, 1, 1, 1)
    fidx = torch.arange(Fd, dtype=torch.int64).view(1, 1, 1, 1, Fd)
    m = inside.unsqueeze(-1).to(F32)
    g = grad * m
    igrad = torch.zeros_like(flat)
    taps = {}
    for name, yy, xx, wgt in (("00", y0c, x0c, wy0 * wx0), ("01", y0c, x1c, wy0 * wx1),
                              ("10", y1c, x0c, wy1 * wx0), ("11", y1c, x1c, wy1 * wx1)):
        lin = (((bidx * H + yy) * W + xx) * Fd + fidx).reshape(-1)
        taps[name] = flat[lin].reshape(B, H, W, S, Fd, C)
        igrad.index_add_(0, lin, (g * wgt).reshape(-1, C))
    gx = (g * (wy0 * (taps["01"] - taps["00"]) + wy1 * (taps["11"] - taps["10"]))).sum(-1)
    gy = (g * (wx0 * (taps["10"] - taps["00"]) + wx1 * (taps["11"] - taps["01"]))).sum(-1)
    return igrad.reshape(B, H, W, Fd, C), torch.stack((gx, gy), dim=-1)


def dense_image_warp(image, flow, use_cuda_backproject=False, back_project_fn=None):
    """image [b,h,w,c], flow [b,h,w,2] (row, col) -> [b,h,w,c]; query = grid + flow (:244).

    ``use_cuda_backproject`` selects the branch at :246-253 (clip to the image, reverse to (x,y),
    BackProject with S=F=1) instead of the python gather path at :255-259.  ``back_project_fn`` replaces the
    restated BackProject by another implementation of the op - the tests pass the reference's own compiled kernel
    (``oracle.ref_binary.back_project``) here.
    """
    b, h, w, c = image.shape
    gy = torch.arange(h, dtype=F32).view(1, h, 1).expand(1, h, w)
    gx = torch.arange(w, dtype=F32).view(1, 1, w).expand(1, h, w)
    q = torch.stack((gy, gx), dim=-1) + flow
    if use_cuda_backproject:
        # tf.clip_by_value = minimum(maximum(x, lo), hi); NaN propagates
        lo = torch.zeros(2, dtype=F32)
        hi = torch.tensor([float(h - 1), float(w - 1)], dtype=F32)
        q = torch.minimum(torch.maximum(q, lo), hi)
        coords = torch.flip(q, dims=[-1]).reshape(b, h, w, 1, 1, 2)
        out = (back_project_fn or back_project)(image.unsqueeze(-2), coords)
        return out.reshape(b, h, w, c)
    return interpolate_bilinear(image, q.reshape(b, h * w, 2)).reshape(b, h, w, c)

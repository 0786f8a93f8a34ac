"""Oracle: the two cost volumes (test infrastructure, see oracle/__init__.py).

* ``get_parallax_sweeping_cv`` restates ``/root/reference/utils/depth_operations.py:223-281`` (PSCV /
  "DSCV") *literally*, including the (2r+1)x ``tile_in_batch`` copies (:217-221, :267-268), so that
  it is also the CPU timing baseline for that function.
* ``cost_volume`` restates :283-313 (SNCV): 49*cuts slice -> multiply -> reduce_mean, leaky 0.1.

fp16 correlate (:276-278).  ``tf.cast(c1, fp16) * tf.cast(c2_w, fp16)`` rounds both operands and
the product to fp16.  ``tf.reduce_mean`` over fp16 does not have a defined summation order in TF
(Eigen CPU, XLA and cuDNN-less GPU paths differ), so two definitions are offered:

* ``half_mode="fp32acc"`` (the contract the CUDA kernel implements): fp16-rounded products are
  summed in fp32 in channel order, divided by n in fp32, rounded once to fp16, cast to fp32.
* ``half_mode="seqfp16"``: sequential fp16 adds (each rounded to fp16) then an fp16 divide - the
  most pessimistic reading of an Eigen half reducer.  Tests bound the gap between the two.
"""
import torch

from ._ieee import sqrt as ieee_sqrt

from .geometry import _epipolar_terms
from .warp import dense_image_warp

F32 = torch.float32
F16 = torch.float16


def tile_in_batch(x, n):
    """[b,...] -> [n*b,...], copy-major (index = copy*b + batch)  (:217-221)."""
    return x.unsqueeze(0).expand(n, *x.shape).reshape(-1, *x.shape[1:])


def _half_group_mean(prod_h, nbre_cuts, half_mode):
    """prod_h [...,c] fp16 products -> [cuts, ...] fp32 group means."""
    c = prod_h.shape[-1]
    gw = c // nbre_cuts
    g = prod_h.reshape(*prod_h.shape[:-1], nbre_cuts, gw)
    if half_mode == "fp32acc":
        acc = torch.zeros(g.shape[:-1], dtype=F32)
        gf = g.to(F32)
        for j in range(gw):
            acc = acc + gf[..., j]
        m = (acc / float(gw)).to(F16)
    elif half_mode == "seqfp16":
        acc = torch.zeros(g.shape[:-1], dtype=F16)
        for j in range(gw):
            acc = (acc.to(F32) + g[..., j].to(F32)).to(F16)
        m = (acc.to(F32) / float(gw)).to(F16)
    else:
        raise ValueError(half_mode)
    return m.movedim(-1, 0).to(F32)


def get_parallax_sweeping_cv(c1, c2, disp_prev_t, disp, rot, trans, camera, search_range,
                             nbre_cuts=1, use_cuda_backproject=True, half_mode="fp32acc", back_project_fn=None):
    """-> (cv [b,h,w,cuts*(2r+1)] cut-major, prev_disp [b,h,w,2r+1])  (:223-281)."""
    b, h, w = c1.shape[0:3]
    n = 2 * search_range + 1
    expl = torch.arange(-search_range, search_range + 1, dtype=F32).view(n, 1, 1, 1, 1)

    d = tile_in_batch(disp, n).reshape(n, b, h, w, 1) + expl
    d = torch.clamp(d, 1e-6, 1e6)                                  # :236

    g = _epipolar_terms(c1, rot, trans, camera)                   # :239-259
    dx = g["dx"].view(1, b, h, w, 1)
    dy = g["dy"].view(1, b, h, w, 1)
    start = torch.stack((g["sx"], g["sy"]), dim=-1).view(1, b, h, w, 2)
    proj = torch.stack((g["px"], g["py"]), dim=-1).view(1, b, h, w, 2)
    s = ieee_sqrt(dx * dx + dy * dy)                              # :261
    div = s / d                                                    # :262
    delta = torch.cat((dx / div, dy / div), dim=-1)                # :263
    flow = (proj + delta) - start                                  # :264
    flow = torch.flip(flow, dims=[-1]).reshape(n * b, h, w, 2)     # :265

    c1_t = tile_in_batch(c1, n)                                    # :267
    comb = tile_in_batch(torch.cat((c2, disp_prev_t), dim=-1), n)  # :268
    comb_w = dense_image_warp(comb, flow, use_cuda_backproject, back_project_fn)    # :270
    c2_w = comb_w[..., :-1]
    prev_disp = comb_w[..., -1]

    prod = c1_t.to(F16) * c2_w.to(F16)                             # :276
    cv = _half_group_mean(prod, nbre_cuts, half_mode)              # [cuts, n*b, h, w]  :277
    cv = cv.reshape(nbre_cuts * n, b, h, w).permute(1, 2, 3, 0).contiguous()       # :278
    prev_disp = prev_disp.reshape(n, b, h, w).permute(1, 2, 3, 0).contiguous()     # :280
    return cv, prev_disp


def cost_volume(c1, c2, search_range, dilation_rate=1, nbre_cuts=1):
    """SNCV: out[..., (dy*(2r+1)+dx)*cuts + k] = leaky_0.1(mean_group(c1 * shift(c2)))  (:283-313)."""
    b, h, w, c = c2.shape
    r = search_range * dilation_rate
    pad = torch.zeros(b, h + 2 * r, w + 2 * r, c, dtype=F32)
    pad[:, r:r + h, r:r + w, :] = c2
    n = 2 * search_range + 1
    gw = c // nbre_cuts
    c1g = c1.reshape(b, h, w, nbre_cuts, gw)
    outs = []
    for y in range(n):
        for x in range(n):
            sl = pad[:, y * dilation_rate:y * dilation_rate + h,
                     x * dilation_rate:x * dilation_rate + w, :].reshape(b, h, w, nbre_cuts, gw)
            outs.append((c1g * sl).mean(dim=-1))                   # [b,h,w,cuts]
    cv = torch.stack(outs, dim=3).reshape(b, h, w, n * n * nbre_cuts)
    return torch.where(cv >= 0, cv, cv * 0.1)

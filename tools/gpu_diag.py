#!/usr/bin/env python
"""Diagnostic (GPU box): mismatch statistics of libm4d vs the oracle, without stopping at the first failure."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import m4depth_b200 as m
from test_gpu_parity import pscv_inputs, dev_cam, cu, LEVEL_SHAPES


def stats(name, got, want):
    got, want = got.detach().cpu(), want.detach().cpu()
    ne = (got != want) & ~(torch.isnan(got) & torch.isnan(want))
    n = int(ne.sum())
    if n == 0:
        print(f"  {name}: EXACT ({got.numel()} values)")
        return
    d = (got.double() - want.double()).abs()
    rel = d / want.double().abs().clamp(min=1e-30)
    idx = torch.nonzero(ne)[:5].tolist()
    print(f"  {name}: {n}/{got.numel()} differ, max abs {float(d.max()):.3e}, max rel {float(rel[ne].max()):.3e}; first {idx}")
    for i in idx[:3]:
        print(f"     got {got[tuple(i)].item():.9g} want {want[tuple(i)].item():.9g}")


for shape in LEVEL_SHAPES[:6]:
    name, b, h, w, c, cuts, kind = shape
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(100 + h * w + c, b, h, w, c, cuts, kind)
    for interp, iname in ((m.INTERP_GATHER, "gather"), (m.INTERP_BP, "bp")):
        print(name, iname)
        want_cv, want_pd = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts,
                                                           use_cuda_backproject=(interp == m.INTERP_BP))
        cv, pd, idx = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4,
                                                       nbre_cuts=cuts, interp=interp, return_index_grids=True)
        qy, qx = oracle.pscv_query_points(pl, rot, trans, cam, 4)
        qxc = torch.minimum(torch.maximum(qx, torch.zeros(())), torch.tensor(float(w - 1)))
        qyc = torch.minimum(torch.maximum(qy, torch.zeros(())), torch.tensor(float(h - 1)))
        x0, x1, y0, y1, _ = oracle.back_project_index_grids(torch.stack((qxc, qyc), dim=-1), h, w)
        want_idx = torch.stack((x0, x1, y0, y1), dim=-1).permute(1, 2, 3, 0, 4)
        stats("idx", idx, want_idx)
        stats("prev_disp", pd, want_pd)
        stats("cv", cv, want_cv)

# geometry pieces on their own
g = torch.Generator().manual_seed(1)
b, h, w = 2, 12, 40
_, _, pt, pl, rot, trans, cam = pscv_inputs(5, b, h, w, 32, 2, "kitti")
depth = torch.exp(torch.rand(b, h, w, 1, generator=g) * 3 + 1)
dc = dev_cam(cam)
print("geometry")
stats("rot_mat", m.utils.get_rot_mat(cu(rot)), oracle.get_rot_mat(rot))
stats("prev_d2para", m.utils.prev_d2para(cu(depth), cu(rot), cu(trans), dc), oracle.prev_d2para(depth, rot, trans, cam))
stats("parallax2depth", m.utils.parallax2depth(cu(pl), cu(rot), cu(trans), dc), oracle.parallax2depth(pl, rot, trans, cam))
stats("depth2parallax", m.utils.depth2parallax(cu(depth), cu(rot), cu(trans), dc), oracle.depth2parallax(depth, rot, trans, cam))

#!/usr/bin/env python
"""Localise which primitive differs between torch-CPU (oracle) and CUDA (IEEE) evaluation of prev_d2para."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle, m4depth_b200 as m
from test_gpu_parity import pscv_inputs, dev_cam, cu

g = torch.Generator().manual_seed(1)
b, h, w = 2, 12, 40
_, _, pt, pl, rot, trans, cam = pscv_inputs(5, b, h, w, 32, 2, "kitti")
depth = torch.exp(torch.rand(b, h, w, 1, generator=g) * 3 + 1)

def chain(dev):
    F32 = torch.float32
    d = depth.to(dev); f = cam["f"].to(dev); c = cam["c"].to(dev); t = trans.to(dev)
    xs = torch.arange(w, dtype=F32, device=dev) + 0.5
    ys = torch.arange(h, dtype=F32, device=dev) + 0.5
    mx = xs.view(1, 1, w).expand(b, h, w) - c[:, 0].view(b, 1, 1)
    my = ys.view(1, h, 1).expand(b, h, w) - c[:, 1].view(b, 1, 1)
    fx = f[:, 0].view(b, 1, 1); fy = f[:, 1].view(b, 1, 1); tz = t[:, 2].view(b, 1, 1)
    nx = mx / fx; ny = my / fy
    sx = nx * fx; sy = ny * fy
    stx = (t[:, 0] * f[:, 0]).view(b, 1, 1); sty = (t[:, 1] * f[:, 1]).view(b, 1, 1)
    den = d[..., 0] - tz
    ax = stx - tz * sx; ay = sty - tz * sy
    vx = ax / den; vy = ay / den
    ss = vx * vx + vy * vy
    out = torch.sqrt(ss)
    return dict(mx=mx, nx=nx, sx=sx, den=den, ax=ax, vx=vx, vy=vy, ss=ss, out=out)

cpu, gpu = chain("cpu"), chain("cuda")
for k in cpu:
    ne = int((cpu[k] != gpu[k].cpu()).sum())
    print(f"{k}: {ne} differ of {cpu[k].numel()}")
mine = m.utils.prev_d2para(cu(depth), cu(rot), cu(trans), dev_cam(cam))[..., 0]
print("mine vs torch-cuda chain:", int((mine != gpu["out"]).sum()), " mine vs cpu:", int((mine.cpu() != cpu["out"]).sum()))
print("oracle vs cpu chain:", int((oracle.prev_d2para(depth, rot, trans, cam)[..., 0] != cpu["out"]).sum()))
# primitives in isolation
a = torch.rand(1 << 20, generator=g) * 10 + 0.01
bb = torch.rand(1 << 20, generator=g) * 10 + 0.01
for name, fn in (("div", lambda x, y: x / y), ("mul", lambda x, y: x * y), ("sqrt", lambda x, y: torch.sqrt(x)),
                 ("div_bcast", lambda x, y: x.view(-1, 1024) / y.view(-1, 1024)[:, :1])):
    r1 = fn(a, bb); r2 = fn(a.cuda(), bb.cuda()).cpu()
    print(name, int((r1 != r2).sum()), "differ")
print(torch.__config__.show().split("\n")[0:12])

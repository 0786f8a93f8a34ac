"""Design aid for the shared-memory staged PSCV (csrc/pscv_smem.cu): how large is the c2 window a pixel tile gathers from?

For each (tile_w x tile_h) tile of a level-2 map: bounding box of the gather-convention taps (x0..x0+1, y0..y0+1 over the
nine hypotheses, dense_image_warp.py:135-149) -> distribution of box width / height / area.  Two data sources:
  micro   the kernel microbench distribution of SURVEY.md 8(d): para_prev_l = exp(U[log .5, log 16]) per pixel
  insitu  para_prev_l of level 2 as the random-weight bench model produces it (oracle on CPU, bench.synth_frames)
CPU only (uses oracle/: this is a tool, not product code).
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def box_stats(para_l, rot, trans, cam, tw, th, label):
    b, h, w, _ = para_l.shape
    qy, qx = oracle.pscv_query_points(para_l, rot, trans, cam, 4)            # [9,b,h,w]
    ok = torch.isfinite(qx) & torch.isfinite(qy)
    x0 = torch.clamp(torch.floor(qx), 0, w - 2)
    y0 = torch.clamp(torch.floor(qy), 0, h - 2)
    big = 1e9
    x0lo = torch.where(ok, x0, torch.full_like(x0, big)).amin(0)
    x0hi = torch.where(ok, x0, torch.full_like(x0, -big)).amax(0)
    y0lo = torch.where(ok, y0, torch.full_like(y0, big)).amin(0)
    y0hi = torch.where(ok, y0, torch.full_like(y0, -big)).amax(0)
    ty, tx = -(-h // th), -(-w // tw)

    def tiles(t, red, fill):
        p = torch.full((b, ty * th, tx * tw), fill)
        p[:, :h, :w] = t
        p = p.reshape(b, ty, th, tx, tw)
        return red(red(p, 4), 2)

    amin = lambda t, d: t.amin(d)
    amax = lambda t, d: t.amax(d)
    ex = tiles(x0hi, amax, -big) - tiles(x0lo, amin, big) + 2
    ey = tiles(y0hi, amax, -big) - tiles(y0lo, amin, big) + 2
    area = (ex * ey).flatten()
    q = lambda t, p: float(torch.quantile(t.flatten().float(), p))
    print(f"{label:8s} tile {tw}x{th}: box w p50/p90/max {q(ex,.5):.0f}/{q(ex,.9):.0f}/{ex.max():.0f}  "
          f"h {q(ey,.5):.0f}/{q(ey,.9):.0f}/{ey.max():.0f}  area px p50/p90/p99/max "
          f"{q(area,.5):.0f}/{q(area,.9):.0f}/{q(area,.99):.0f}/{area.max():.0f}  mean area/tile {float(area.mean()) / (tw * th):.2f}x")
    return area


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--insitu", action="store_true")
    ap.add_argument("--frames", type=int, default=3)
    args = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    b, h, w = 2, 96, 320
    cam = {"f": torch.tensor([[0.580948 * w, 1.924101 * h]] * b), "c": torch.tensor([[0.490788 * w, 0.460944 * h]] * b)}
    rot = torch.cat([torch.ones(b, 1), 0.01 * torch.randn(b, 3, generator=g)], 1)
    rot = rot / rot.norm(dim=1, keepdim=True)
    trans = torch.tensor([0.0, 0.0, 1.0]) + torch.randn(b, 3, generator=g) * torch.tensor([0.05, 0.05, 0.3])
    import math
    pl = torch.exp(torch.rand(b, h, w, 1, generator=g) * (math.log(16) - math.log(.5)) + math.log(.5))
    for tw, th in ((16, 8), (32, 8), (16, 16), (32, 4), (8, 8)):
        box_stats(pl, rot, trans, cam, tw, th, "micro")
    if not args.insitu:
        return
    sys.argv = [sys.argv[0]]
    import bench
    import importlib.util as _u
    _s = _u.spec_from_file_location("_w", os.path.join(ROOT, "m4depth_b200", "weights.py")); _m = _u.module_from_spec(_s); _s.loader.exec_module(_m)
    init_random_weights = _m.init_random_weights
    torch.set_num_threads(os.cpu_count())
    model = oracle.M4Depth(init_random_weights(6, seed=7), nbre_levels=6, pscv_kwargs={"use_cuda_backproject": False})
    cam1 = bench.kitti_camera(1)
    frames = bench.synth_frames(args.frames + 1, 1, 1234)
    lvl2 = model.d_estimator.levels[1]
    lvl2.trace = {}
    with torch.no_grad():
        for t, fr in enumerate(frames):
            s = dict(fr)
            s["new_traj"] = torch.tensor([t == 0])
            model([[s], cam1])
            if t == 0:
                continue
            pl2 = lvl2.trace["para_prev_l"]
            print(f"frame {t}: para_prev_l min/median/p99/max {float(pl2.min()):.3g}/{float(pl2.median()):.3g}/"
                  f"{float(torch.quantile(pl2.flatten(), .99)):.3g}/{float(pl2.max()):.3g}")
            lcam = {"f": cam1["f"] / 4.0, "c": cam1["c"] / 4.0}
            for tw, th in ((16, 8), (32, 8)):
                box_stats(pl2, s["rot"], s["trans"], lcam, tw, th, f"insitu{t}")


if __name__ == "__main__":
    main()

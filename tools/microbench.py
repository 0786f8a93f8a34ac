#!/usr/bin/env python
"""Kernel microbenchmarks on the GPU box (not the bench contract; an iteration tool).

  python tools/microbench.py pscv      fused backproject+PSCV at the level shapes of config 3 (b=8), 3 interp modes
  python tools/microbench.py sncv      SNCV at the same shapes
  python tools/microbench.py conv      every conv layer shape of a config-3 frame (b=8)
L2 is flushed (256 MB write) before every timed launch; times are CUDA-event medians.
"""
import os, sys, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import m4depth_b200 as m
from m4depth_b200.m4depth_network import _Conv2D

PEAK = 6547.5
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=12, flush=os.environ.get("NOFLUSH") != "1"):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


LEVELS = [(192, 640, 16, 1), (96, 320, 32, 2), (48, 160, 64, 2), (24, 80, 96, 4), (12, 40, 128, 4), (6, 20, 192, 8)]


def gn(x, cuts):
    b, h, w, c = x.shape
    g = x.reshape(b, h, w, cuts, c // cuts)
    return (g / g.norm(dim=-1, keepdim=True)).reshape(b, h, w, c).contiguous()


def pscv(b=8):
    g = torch.Generator(device="cuda").manual_seed(0)
    for lvl, (h, w, c, cuts) in enumerate(LEVELS, 1):
        lk = lambda t: torch.where(t >= 0, t, 0.1 * t)
        c1 = gn(lk(torch.randn(b, h, w, c, device="cuda", generator=g)), cuts)
        c2 = gn(lk(torch.randn(b, h, w, c, device="cuda", generator=g)), cuts)
        pl = torch.exp(torch.rand(b, h, w, 1, device="cuda", generator=g) * (math.log(16) - math.log(0.5)) + math.log(0.5))
        pt = torch.exp(torch.rand(b, h, w, 1, device="cuda", generator=g) * (math.log(16) - math.log(0.05)) + math.log(0.05))
        rot = torch.cat([torch.ones(b, 1, device="cuda"), 0.01 * torch.randn(b, 3, device="cuda", generator=g)], 1)
        rot = (rot / rot.norm(dim=1, keepdim=True)).contiguous()
        trans = (torch.tensor([0.0, 0.0, 1.0], device="cuda") + torch.randn(b, 3, device="cuda", generator=g) * torch.tensor([0.05, 0.05, 0.3], device="cuda")).contiguous()
        cam = {"f": torch.tensor([[0.580948 * w, 1.924101 * h]] * b, device="cuda"), "c": torch.tensor([[0.490788 * w, 0.460944 * h]] * b, device="cuda")}
        smooth = os.environ.get("SMOOTH") == "1"
        if smooth:
            pl = torch.nn.functional.avg_pool2d(pl.permute(0, 3, 1, 2), 9, 1, 4).permute(0, 2, 3, 1).contiguous()
        bytes_p9 = 4 * h * w * (2 * c + 2 + 9 * cuts + 9) * b
        bytes_p1 = 4 * h * w * (2 * c + 2 + 9 * cuts + 1) * b
        if os.environ.get("LEVELS") and str(lvl) not in os.environ["LEVELS"].split(","):
            continue
        L = m._lib
        cv = torch.empty(b, h, w, 9 * cuts, device="cuda"); pd = torch.empty(b, h, w, 9, device="cuda"); cl = torch.empty(b, h, w, device="cuda")
        st = L.stream()

        def raw(mode, p9):
            # direct C-ABI call on preallocated outputs: no host-side allocation between the timing events
            L.check(L.lib.m4d_pscv_fused_fwd_ex(L.ptr(c1), L.ptr(c2), L.ptr(pt), L.ptr(pl), L.ptr(rot), 4, L.ptr(trans), L.ptr(cam["f"]),
                                                L.ptr(cam["c"]), b, h, w, c, cuts, 4, L.ptr(cv), 9 * cuts, L.ptr(pd) if p9 else None, 9,
                                                None if p9 else L.ptr(cl), 1, 0.5, None, mode, st))
        if os.environ.get("PARA") == "insitu":      # what the random-weight bench model feeds level 2: parallax 0.6 .. 2.3
            pl = 0.6 + 1.7 * torch.rand(b, h, w, 1, device="cuda", generator=g)
        variants = [(0, "gather"), (0x400, "gather/warp"), (1, "bp"), (2, "bp_fma"), (0x200, "gather/tile"), (0x100, "gather/generic")]
        if os.environ.get("VARIANTS"):
            variants = [v for v in variants if v[1] in os.environ["VARIANTS"].split(",")]
        if os.environ.get("CPS"):
            variants = [((int(c[0]) << 16) | (int(c[1:] or 0) << 12), f"gather/var{c[0]}cps{c[1:]}") for c in os.environ["CPS"].split(",")] + [(0x200, "gather/tile")]
        for mode, name in variants:
            for p9 in (True, False):
                med, mn = timeit(lambda: raw(mode, p9))
                nb = bytes_p9 if p9 else bytes_p1
                print(f"pscv L{lvl} {h}x{w}x{c} b={b} {name:14s} P={9 if p9 else 1}: median {med:8.1f} us  min {mn:8.1f} us  -> {nb / med / 1e3:7.1f} GB/s "
                      f"({nb / med / 1e3 / PEAK * 100:5.1f}% of measured HBM peak, {nb / 1e6:.1f} MB)")


def sncv(b=8):
    g = torch.Generator(device="cuda").manual_seed(0)
    for lvl, (h, w, c, cuts) in enumerate(LEVELS, 1):
        f = gn(torch.randn(b, h, w, c, device="cuda", generator=g), cuts)
        nbytes = 4 * h * w * (c + 49 * cuts) * b
        out = torch.empty(b, h, w, 49 * cuts, device="cuda")
        L = m._lib
        for variant, name in ((0, "auto"), (1, "pixel_dy")):
            med, mn = timeit(lambda: L.check(L.lib.m4d_sncv_fwd_ex(L.ptr(f), L.ptr(f), b, h, w, c, cuts, 3, L.ptr(out), 49 * cuts, variant, L.stream())))
            print(f"sncv L{lvl} {h}x{w}x{c} b={b} {name:8s}: median {med:8.1f} us  min {mn:8.1f} us -> {nbytes / mn / 1e3:7.1f} GB/s ({nbytes / mn / 1e3 / PEAK * 100:5.1f}%)")


def conv(b=8):
    g = torch.Generator(device="cuda").manual_seed(0)
    H, W = 384, 1280
    shapes = []
    cin, h, w = 3, H, W
    for co in (16, 32, 64, 96, 128, 192):
        shapes.append(("enc_s1", h, w, cin, co, 1)); shapes.append(("enc_s2", h, w, co, co, 2))
        cin, h, w = co, -(-h // 2), -(-w // 2)
    for lvl, (h, w, c, cuts) in enumerate(LEVELS, 1):
        ci = 58 * cuts + 6
        for co in (128, 128, 96, 64, 32, 16, 5):
            shapes.append((f"ref_L{lvl}", h, w, ci, co, 1)); ci = co
    tot = 0.0; totf = 0.0
    for name, h, w, ci, co, s in shapes:
        xs = (ci + 3) // 4 * 4
        x = torch.randn(b, h, w, xs, device="cuda", generator=g)[..., :ci]
        cv = _Conv2D(co, s)
        cv.assign(torch.randn(3, 3, ci, co) * 0.05, torch.zeros(co), "cuda")
        if os.environ.get("CONV_ONLY") and not name.startswith(tuple(os.environ["CONV_ONLY"].split(","))):
            continue
        algo = int(os.environ.get("CONV_ALGO", "0"))
        if algo == 2 and cv.packed is None:
            algo = 1
        force = int(os.environ.get("CONV_FORCE", "0"))          # slices | halo stages << 4 (tuning)
        med, mn = timeit(lambda: cv(x, alpha=0.1, algo=algo, slices=force), iters=5)
        fl = 2.0 * 9 * ci * co * (-(-h // s)) * (-(-w // s)) * b
        tot += med; totf += fl
        path = "tc" if (algo != 1 and cv.packed is not None and ci >= cv.tc_min_cin) else "ffma"
        print(f"conv {name:8s} {h:4d}x{w:4d} {ci:3d}->{co:3d} s{s} {path:4s}: {med:9.1f} us  {fl / med / 1e6:7.2f} TFLOP/s")
    print(f"conv total {tot / 1e3:.2f} ms per step of {b} frames, {totf / tot / 1e6:.2f} TFLOP/s average")


if __name__ == "__main__":
    nb = int(os.environ.get("B", "8"))
    for a in sys.argv[1:] or ["pscv", "sncv", "conv"]:
        {"pscv": pscv, "sncv": sncv, "conv": conv}[a](nb)

#!/usr/bin/env python
"""Where a tensor-core conv layer's warp roles spend their clocks (GPU box; library built with the role timers:
    M4D_NVCC_EXTRA=-DM4D_TC_PROFILE python m4depth_b200/_build.py --force).

    python tools/conv_phases.py [b h w cin cout stride]...      default: the level-1 refiner layers of config 3

Prints, per layer, clocks per k-block (averaged over the CTAs) of: the A producer's wait for a free halo stage; the splitter's
wait for the landed halo and its split; the issuer's waits (accumulator set, split halo, weight slabs) and its issue regions;
the epilogue's wait for a finished accumulator set, its TMEM drain and its per-tile finalisation (bias, activation, store)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import m4depth_b200 as m
from m4depth_b200.m4depth_network import _Conv2D

L = m._lib
NAMES = ["A:wait_empty", "I:wait_acc", "I:wait_halo", "I:wait_wgt", "I:issue", "S:wait_landed", "S:split", "E:wait_acc", "E:drain", "E:final",
         "CTA:life", "S:bar", "S:fence", "E:store_wait", "S:load+max", "S:convert"]          # E:* are group 0's (thin layers: it owns every other tile); CTA:life = SM clocks the CTA ran for


def run(b, h, w, cin, cout, stride):
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    buf = torch.zeros(sms * 16, dtype=torch.int64, device="cuda")
    if not L.lib.m4d_debug_conv_profile(buf.data_ptr()):
        sys.exit("libm4d was built without -DM4D_TC_PROFILE")
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = (cin + 3) // 4 * 4
    x = torch.randn(b, h, w, xs, device="cuda", generator=g)[..., :cin]
    conv = _Conv2D(cout, stride)
    conv.assign(torch.randn(3, 3, cin, cout) * 0.05, torch.zeros(cout), "cuda")
    for _ in range(3):
        conv(x, alpha=0.1, algo=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); conv(x, alpha=0.1, algo=2); e1.record(); torch.cuda.synchronize()
    L.lib.m4d_debug_conv_profile(None)
    p = buf.view(sms, 16).double()
    oh, ow = (h, w) if stride == 1 else (h // 2, w // 2)
    tiles = -(-ow // 8) * -(-oh // 16) * b
    kblocks = -(-cin // 32) if stride == 1 else 2 * (2 * cin // 32)
    per_cta_kb = tiles / min(tiles, sms) * kblocks
    us = e0.elapsed_time(e1) * 1e3
    clk = us * 1e-6 * 1.965e9 / per_cta_kb
    print(f"{h}x{w} {cin}->{cout} s{stride}: {us:7.1f} us, {tiles} tiles, {kblocks} k-blocks/tile, ~{clk:6.0f} clk per k-block (at 1.965 GHz)")
    print("   " + "  ".join(f"{n} {float(p[:, i].mean()) / per_cta_kb:6.0f}" for i, n in enumerate(NAMES)))


if __name__ == "__main__":
    args = [int(v) for v in sys.argv[1:]]
    cfgs = [tuple(args[i:i + 6]) for i in range(0, len(args), 6)] or [
        (8, 192, 640, 128, 128, 1), (8, 192, 640, 96, 64, 1), (8, 192, 640, 64, 32, 1), (8, 192, 640, 32, 16, 1), (8, 192, 640, 16, 5, 1),
        (8, 384, 1280, 16, 16, 2)]
    for c in cfgs:
        run(*c)

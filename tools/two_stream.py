#!/usr/bin/env python
"""Experiment (GPU box): the 8 sequences of a GPU as one batch of 8 on one stream vs G groups of 8/G on G streams.
The small-grid kernels of the deep pyramid levels (a few dozen CTAs on 148 SMs) of one group then overlap the large
kernels of another.  Prints frames/s of each arrangement (device-resident inputs, CUDA graphs)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import m4depth_b200 as m
from m4depth_b200.weights import init_random_weights
import bench

H, W, LEVELS = 384, 1280, 6
dev = torch.device("cuda", 0)
wts = init_random_weights(LEVELS, seed=7)


def run(groups, steps=30, warm=8):
    b = 8 // groups
    models, streams, cams, pools = [], [], [], []
    for gi in range(groups):
        mod = m.M4Depth(nbre_levels=LEVELS, use_cuda_graph=True)
        mod.load_weights(wts)
        models.append(mod)
        streams.append(torch.cuda.Stream(device=dev))
        cam = bench.kitti_camera(b)
        cams.append({k: v.to(dev) for k, v in cam.items()})
        pool = bench.synth_frames(4, b, seed=1234 + gi)
        pools.append([{k: v.to(dev) for k, v in fr.items()} for fr in pool])
    torch.cuda.synchronize()

    def step(t):
        for gi in range(groups):
            with torch.cuda.stream(streams[gi]):
                fr = pools[gi][t % 4]
                models[gi]([[{"RGB_im": fr["RGB_im"], "rot": fr["rot"], "trans": fr["trans"], "new_traj": [t == 0]}], cams[gi]])

    t = 0
    for _ in range(warm):
        step(t); t += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for _ in range(steps):
        step(t); t += 1
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{groups} group(s) of {b}: {ms:.3f} ms per step of 8 frames = {8 / ms * 1e3:.0f} frames/s")


for g in (1, 2, 4):
    run(g)

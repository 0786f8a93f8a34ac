#!/usr/bin/env python
"""Whole-model depth error against the golden vectors (recorded from the reference's Python on tools/tf_shim) for one
conv configuration: run it under M4D_CONV_ALGO=1 (FFMA2 everywhere), M4D_CONV_PREC=0 (tensor cores, 3xTF32) and with neither
(default: tensor cores, 3xFP16) to see what the convolution's summation error does to the depth maps after the fp16 stage
of the PSCV has amplified it (DESIGN.md section 3 quotes the numbers).  GPU box only."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
import m4depth_b200 as m

T = torch.from_numpy
for case in ("cfg1", "odd"):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"model_{case}.npz")))
    nl = int(g["nbre_levels"])
    model = m.M4Depth(nbre_levels=nl, use_cuda_graph=False)
    model.load_weights(oracle.init_weights(nl, seed=int(g["weights_seed"]), bias_std=0.05, dn_random=True))
    model.set_interp(m.INTERP_BP if bool(g["backproject"]) else m.INTERP_GATHER)
    cam = {"f": T(g["cam_f"]).cuda(), "c": T(g["cam_c"]).cuda()}
    t = 0
    while f"rgb_{t}" in g:
        b = g[f"rgb_{t}"].shape[0]
        s = {"RGB_im": T(g[f"rgb_{t}"]).cuda(), "rot": T(g[f"rot_{t}"]).cuda(), "trans": T(g[f"trans_{t}"]).cuda(), "new_traj": [t == 0] * b}
        got = model([[s], cam])["depth"].cpu().numpy()
        want = g[f"depth_{t}"]
        err = np.abs(got - want) / (np.abs(want) + 0.1)
        print(f"{case} frame {t}: median {np.median(err):.2e}  q90 {np.percentile(err, 90):.2e}  q99 {np.percentile(err, 99):.2e}  max {err.max():.2e}")
        t += 1

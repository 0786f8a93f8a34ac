#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python files on tools/tf_shim.

Run in the build container only (needs /root/reference; the GPU box never runs this):
    python tools/gen_golden.py

What executes: /root/reference/utils/depth_operations.py, utils/dense_image_warp.py and
m4depth_network.py, imported unmodified, with ``tensorflow`` resolved to tools/tf_shim (see its
docstring for the primitive semantics it restates).  For the BackProject branch of
dense_image_warp (:246-253) the op library cannot be loaded, so ``back_project`` is bound to
oracle.warp.back_project (the restated kernel body, backproject_op_gpu.cu.cc:19-79); the
python-gather branch needs no substitution.

Outputs (all float32 unless noted), seeds fixed:
  pscv_<case>.npz     inputs + cv/prev_disp from get_parallax_sweeping_cv (gather and backproject branch)
  sncv_<case>.npz     cost_volume
  geom_<case>.npz     prev_d2para / parallax2depth / depth2parallax / get_rot_mat / dense_image_warp
  dn.npz              DomainNormalization.call
  model_cfg1.npz      M4Depth (3 levels, 128x128, 2 frames, b=1 = BASELINE config 1): inputs, per-frame depth and
                      per-level state depth; weights = oracle.init_weights(nl, seed, 0.05, True) (seed stored)
  model_odd.npz       M4Depth (6 levels, 96x160 -> L6 2x3 with odd sizes on the way, b=2, 3 frames)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "tf_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tensorflow as tf                       # the shim  # noqa: E402
import utils.depth_operations as ref_ops      # reference  # noqa: E402
ref_diw = sys.modules["utils.dense_image_warp"]   # the module (utils/__init__.py rebinds the name to the function)
import m4depth_network as ref_net             # reference  # noqa: E402
from oracle.warp import back_project as oracle_back_project   # noqa: E402
from oracle.network import init_weights   # noqa: E402  (synthetic weights only; no oracle math is used here)

OUT = os.path.join(ROOT, "tests", "golden")
T = tf.Tensor


def npy(x):
    return x.t.numpy() if isinstance(x, tf.Tensor) else np.asarray(x)


def use_backproject(flag):
    """Flip dense_image_warp.py's module switch (:55-58); the op itself is the restated kernel."""
    ref_diw.use_cuda_backproject = flag
    ref_diw.back_project = lambda f_map, coords: T(oracle_back_project(f_map.t, coords.t))


def group_norm(x, cuts):
    b, h, w, c = x.shape
    g = x.reshape(b, h, w, cuts, c // cuts)
    return (g / torch.sqrt((g * g).sum(-1, keepdim=True))).reshape(b, h, w, c)


def camera_for(kind, b, h, w):
    if kind == "kitti":
        f, c = (0.580948 * w, 1.924101 * h), (0.490788 * w, 0.460944 * h)
    elif kind == "midair":
        f, c = (0.5 * w, 0.5 * h), (0.5 * w, 0.5 * h)
    else:
        f, c = (0.5 * w, 2.0 / 3.0 * h), (0.5 * w, 0.5 * h)
    return {"f": torch.tensor([f] * b, dtype=torch.float32), "c": torch.tensor([c] * b, dtype=torch.float32)}


def motion(g, b):
    rot = torch.cat([torch.ones(b, 1), 0.01 * torch.randn(b, 3, generator=g)], 1)
    rot = rot / rot.norm(dim=1, keepdim=True)
    trans = torch.tensor([0.0, 0.0, 1.0]) + torch.randn(b, 3, generator=g) * torch.tensor([0.05, 0.05, 0.3])
    return rot, trans


def gen_pscv(name, seed, b, h, w, c, cuts, kind):
    g = torch.Generator().manual_seed(seed)
    cam = camera_for(kind, b, h, w)
    rot, trans = motion(g, b)
    lk = lambda t: torch.where(t >= 0, t, 0.1 * t)
    c1 = group_norm(lk(torch.randn(b, h, w, c, generator=g)), cuts)
    c2 = group_norm(lk(torch.randn(b, h, w, c, generator=g)), cuts)
    para_l = torch.exp(torch.rand(b, h, w, 1, generator=g) * (np.log(16) - np.log(0.5)) + np.log(0.5))
    para_t = torch.exp(torch.rand(b, h, w, 1, generator=g) * (np.log(16) - np.log(0.05)) + np.log(0.05))
    tcam = {"f": T(cam["f"]), "c": T(cam["c"])}
    out = dict(c1=c1, c2=c2, para_prev_t=para_t, para_prev_l=para_l, rot=rot, trans=trans,
               cam_f=cam["f"], cam_c=cam["c"], cuts=np.int32(cuts), search_range=np.int32(4))
    for tag, flag in (("gather", False), ("bp", True)):
        use_backproject(flag)
        cv, pd = ref_ops.get_parallax_sweeping_cv(T(c1), T(c2), T(para_t), T(para_l), T(rot), T(trans), tcam, 4,
                                                  nbre_cuts=cuts)
        out["cv_" + tag], out["prev_disp_" + tag] = npy(cv), npy(pd)
    np.savez_compressed(os.path.join(OUT, f"pscv_{name}.npz"), **{k: npy(v) if not isinstance(v, torch.Tensor) else v.numpy() for k, v in out.items()})


def gen_sncv(name, seed, b, h, w, c, cuts):
    g = torch.Generator().manual_seed(seed)
    f = group_norm(torch.randn(b, h, w, c, generator=g), cuts)
    cv = ref_ops.cost_volume(T(f), T(f), 3, nbre_cuts=cuts)
    np.savez_compressed(os.path.join(OUT, f"sncv_{name}.npz"), f=f.numpy(), cuts=np.int32(cuts), out=npy(cv))


def gen_geom(name, seed, b, h, w, kind):
    g = torch.Generator().manual_seed(seed)
    cam = camera_for(kind, b, h, w)
    tcam = {"f": T(cam["f"]), "c": T(cam["c"])}
    rot, trans = motion(g, b)
    depth = torch.exp(torch.rand(b, h, w, 1, generator=g) * 3.0 + 1.0)
    para = torch.exp(torch.rand(b, h, w, 1, generator=g) * 3.0 - 1.0)
    img = torch.randn(b, h, w, 5, generator=g)
    flow = torch.randn(b, h, w, 2, generator=g) * 3.0
    out = dict(rot=rot, trans=trans, cam_f=cam["f"], cam_c=cam["c"], depth=depth, para=para, img=img, flow=flow,
               rot_mat=npy(ref_ops.get_rot_mat(T(rot))),
               prev_d2para=npy(ref_ops.prev_d2para(T(depth), T(rot), T(trans), tcam)),
               parallax2depth=npy(ref_ops.parallax2depth(T(para), T(rot), T(trans), tcam)),
               depth2parallax=npy(ref_ops.depth2parallax(T(depth), T(rot), T(trans), tcam)))
    use_backproject(False)
    out["warp_gather"] = npy(ref_diw.dense_image_warp(T(img), T(flow)))
    use_backproject(True)
    out["warp_bp"] = npy(ref_diw.dense_image_warp(T(img), T(flow)))
    use_backproject(False)
    np.savez_compressed(os.path.join(OUT, f"geom_{name}.npz"), **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})


def gen_dn(seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, 12, 20, 16, generator=g) * torch.rand(1, 1, 1, 16, generator=g) * 3 + torch.randn(1, 1, 1, 16, generator=g)
    dn = ref_net.DomainNormalization()
    _ = dn(T(x))
    scale = torch.rand(1, 1, 1, 16, generator=g) + 0.5
    bias = torch.randn(1, 1, 1, 16, generator=g) * 0.1
    dn.scale.assign(scale)
    dn.bias.assign(bias)
    np.savez_compressed(os.path.join(OUT, "dn.npz"), x=x.numpy(), scale=scale.numpy(), bias=bias.numpy(), out=npy(dn(T(x))))


def import_weights(model, w, nl):
    """Assign the synthetic weights (oracle.init_weights, checkpoint key layout) to the built reference model."""
    enc = model.encoder
    for i in range(nl):
        for tag, cv in (("s1", enc.conv_layers_s1[i]), ("s2", enc.conv_layers_s2[i])):
            cv.kernel.assign(w[f"encoder/conv_layers_{tag}/{i}/kernel"])
            cv.bias.assign(w[f"encoder/conv_layers_{tag}/{i}/bias"])
    enc.dn_layers[0].scale.assign(w["encoder/dn_layers/0/scale"])
    enc.dn_layers[0].bias.assign(w["encoder/dn_layers/0/bias"])
    for i, lvl in enumerate(model.d_estimator.levels):
        r = lvl.disp_refiner
        for tag, layers in (("prep_conv_layers", r.prep_conv_layers), ("est_d_conv_layers", r.est_d_conv_layers)):
            for j, cv in enumerate(layers):
                cv.kernel.assign(w[f"d_estimator/levels/{i}/disp_refiner/{tag}/{j}/kernel"])
                cv.bias.assign(w[f"d_estimator/levels/{i}/disp_refiner/{tag}/{j}/bias"])


def smooth(x):
    k = torch.ones(3, 1, 3, 3) / 9.0
    y = x.permute(0, 3, 1, 2)
    for _ in range(2):
        y = torch.nn.functional.conv2d(torch.nn.functional.pad(y, (1, 1, 1, 1), mode="replicate"), k, groups=3)
    return y.permute(0, 2, 3, 1).contiguous()


def gen_model(name, seed, nl, b, H, W, frames, kind, backproject):
    g = torch.Generator().manual_seed(seed)
    tf.set_initializer_seed(seed)
    use_backproject(backproject)
    model = ref_net.M4Depth(nbre_levels=nl, is_training=False)
    cam = camera_for(kind, b, H, W)
    out = {"cam_f": cam["f"].numpy(), "cam_c": cam["c"].numpy(), "nbre_levels": np.int32(nl),
           "backproject": np.int32(backproject)}
    base = smooth(torch.rand(b, H, W, 3, generator=g))
    first = True
    for t in range(frames):
        rot, trans = motion(g, b)
        rgb = torch.roll(base, shifts=(t, 2 * t), dims=(1, 2)) + 0.02 * torch.randn(b, H, W, 3, generator=g)
        sample = {"RGB_im": T(rgb), "rot": T(rot), "trans": T(trans), "new_traj": T(torch.tensor([t == 0] * b))}
        tcam = {"f": T(cam["f"].clone()), "c": T(cam["c"].clone())}
        if first:
            # build all layers (frame 0 never reaches the refiners), then load the synthetic weights
            # (non-zero biases and DN affine so that those are pinned too)
            _ = model([[sample], tcam])
            dummy_cam = {"f": T(cam["f"].clone()), "c": T(cam["c"].clone())}
            s2 = dict(sample)
            s2["new_traj"] = T(torch.tensor([False] * b))
            _ = model([[s2], dummy_cam])
            import_weights(model, init_weights(nl, seed=seed, bias_std=0.05, dn_random=True), nl)
            out["weights_seed"] = np.int32(seed)
            first = False
        res = model([[sample], tcam])
        out[f"rgb_{t}"], out[f"rot_{t}"], out[f"trans_{t}"] = rgb.numpy(), rot.numpy(), trans.numpy()
        out[f"depth_{t}"] = npy(res["depth"])
        for li, lvl in enumerate(model.d_estimator.levels):
            out[f"state_depth_{t}_l{li + 1}"] = npy(lvl.depth_prev_t)
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_pscv("l2_kitti", 11, 2, 12, 40, 32, 2, "kitti")
    gen_pscv("l1_midair", 12, 1, 24, 24, 16, 1, "midair")
    gen_pscv("l4_tartan", 13, 2, 6, 10, 96, 4, "tartan")
    gen_pscv("l6_kitti", 14, 3, 6, 20, 192, 8, "kitti")
    gen_sncv("l2", 21, 2, 12, 20, 32, 2)
    gen_sncv("l6", 22, 1, 6, 5, 192, 8)
    gen_geom("kitti", 31, 2, 12, 40, "kitti")
    gen_geom("tartan", 32, 1, 15, 20, "tartan")
    gen_dn(41)
    gen_model("cfg1", 51, 3, 1, 128, 128, 2, "midair", False)
    gen_model("cfg1_bp", 51, 3, 1, 128, 128, 2, "midair", True)
    gen_model("odd", 52, 6, 2, 96, 160, 3, "tartan", True)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

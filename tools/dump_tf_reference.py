#!/usr/bin/env python
"""Upgrade the parity pin from "restatement" to "TensorFlow": run the UNTOUCHED reference under real TensorFlow on the
seeded inputs of tests/golden/*.npz and write the same arrays next to them (``*_tf.npz``).

This repository's image has no TensorFlow, so this script is never run by its tests; anyone with the reference's
environment (tensorflow-gpu 2.7, README.md:79) can:

    python tools/dump_tf_reference.py --reference /path/to/M4Depth --golden tests/golden --out tests/golden_tf
    python tools/dump_tf_reference.py ... --compare          # also print max |tf - shim| per array

It imports the reference's own modules (utils/depth_operations.py, utils/dense_image_warp.py, m4depth_network.py) exactly
as tools/gen_golden.py does - the only difference is that ``import tensorflow`` resolves to the real package instead of
tools/tf_shim.  The fixtures carry their inputs, so no RNG has to match across frameworks.
"""
import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of michael-fonder/M4Depth")
    ap.add_argument("--golden", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", action="store_true")
    args = ap.parse_args()
    out_dir = args.out or args.golden + "_tf"
    os.makedirs(out_dir, exist_ok=True)
    try:
        import tensorflow as tf
    except ImportError:
        sys.exit("TensorFlow is not installed: this tool needs the reference's environment (tensorflow-gpu 2.7)")
    if not hasattr(tf, "sysconfig"):
        sys.exit("`tensorflow` resolved to tools/tf_shim, not the real package: remove it from sys.path")
    sys.path.insert(0, args.reference)
    from utils import depth_operations as dop          # the reference's own files
    from utils import dense_image_warp as diw

    def cam(g):
        return {"f": tf.constant(g["cam_f"]), "c": tf.constant(g["cam_c"])}

    for name in sorted(os.listdir(args.golden)):
        if not name.endswith(".npz"):
            continue
        g = dict(np.load(os.path.join(args.golden, name)))
        res = {}
        if name.startswith("pscv_"):
            cv, pd = dop.get_parallax_sweeping_cv(tf.constant(g["c1"]), tf.constant(g["c2"]), tf.constant(g["para_prev_t"]),
                                                  tf.constant(g["para_prev_l"]), tf.constant(g["rot"]), tf.constant(g["trans"]),
                                                  cam(g), 4, nbre_cuts=int(g["cuts"]))
            branch = "bp" if getattr(diw, "use_cuda_backproject", False) else "gather"
            res = {"cv_" + branch: cv.numpy(), "prev_disp_" + branch: pd.numpy()}
        elif name.startswith("sncv_"):
            f = tf.constant(g["f"])
            res = {"out": dop.cost_volume(f, f, 3, "cost_volume", nbre_cuts=int(g["cuts"])).numpy()}
        elif name.startswith("geom_"):
            rot, trans = tf.constant(g["rot"]), tf.constant(g["trans"])
            res = {"rot_mat": dop.get_rot_mat(rot).numpy(),
                   "prev_d2para": dop.prev_d2para(tf.constant(g["depth"]), rot, trans, cam(g)).numpy(),
                   "parallax2depth": dop.parallax2depth(tf.constant(g["para"]), rot, trans, cam(g)).numpy(),
                   "depth2parallax": dop.depth2parallax(tf.constant(g["depth"]), rot, trans, cam(g)).numpy()}
        else:
            continue                                   # model-level fixtures need the weights: see gen_golden.py --help
        np.savez_compressed(os.path.join(out_dir, name.replace(".npz", "_tf.npz")), **res)
        if args.compare:
            for k, v in res.items():
                if k in g:
                    d = np.abs(v.astype(np.float64) - g[k].astype(np.float64))
                    print(f"{name}:{k}: max |tf - fixture| = {d.max():.3e}  ({(d > 0).mean() * 100:.3f} % of entries differ)")
        print("wrote", name.replace(".npz", "_tf.npz"))


if __name__ == "__main__":
    main()

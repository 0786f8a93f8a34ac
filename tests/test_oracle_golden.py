"""Oracle vs the golden vectors recorded from the reference's own Python files (tools/gen_golden.py).

This is the parity pin of the oracle: the vectors come from /root/reference/utils/depth_operations.py,
utils/dense_image_warp.py and m4depth_network.py executed unmodified on tools/tf_shim.
"""
import os

import numpy as np
import pytest
import torch

import oracle

T = torch.from_numpy


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def cam_of(g):
    return {"f": T(g["cam_f"]), "c": T(g["cam_c"])}


@pytest.mark.parametrize("case", ["l2_kitti", "l1_midair", "l4_tartan", "l6_kitti"])
@pytest.mark.parametrize("branch", ["gather", "bp"])
def test_pscv_matches_reference(golden_dir, case, branch):
    g = load(golden_dir, f"pscv_{case}.npz")
    cv, pd = oracle.get_parallax_sweeping_cv(T(g["c1"]), T(g["c2"]), T(g["para_prev_t"]), T(g["para_prev_l"]),
                                             T(g["rot"]), T(g["trans"]), cam_of(g), 4, nbre_cuts=int(g["cuts"]),
                                             use_cuda_backproject=(branch == "bp"))
    # same primitive arithmetic in the same order -> bit-exact
    assert np.array_equal(cv.numpy(), g["cv_" + branch])
    assert np.array_equal(pd.numpy(), g["prev_disp_" + branch])


@pytest.mark.parametrize("case", ["l2_kitti", "l4_tartan"])
def test_pscv_branches_agree(golden_dir, case):
    """python-gather and BackProject branches give the same values."""
    g = load(golden_dir, f"pscv_{case}.npz")
    np.testing.assert_allclose(g["prev_disp_gather"], g["prev_disp_bp"], rtol=1e-5, atol=1e-6)
    # cv is fp16-quantised: allow 1 fp16 ulp of the largest magnitude
    assert np.max(np.abs(g["cv_gather"] - g["cv_bp"])) <= 2.0 ** -14


@pytest.mark.parametrize("case", ["l2", "l6"])
def test_sncv_matches_reference(golden_dir, case):
    g = load(golden_dir, f"sncv_{case}.npz")
    out = oracle.cost_volume(T(g["f"]), T(g["f"]), 3, nbre_cuts=int(g["cuts"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", ["kitti", "tartan"])
def test_geometry_matches_reference(golden_dir, case):
    g = load(golden_dir, f"geom_{case}.npz")
    cam = cam_of(g)
    rot, trans = T(g["rot"]), T(g["trans"])
    assert np.array_equal(oracle.get_rot_mat(rot).numpy(), g["rot_mat"])
    assert np.array_equal(oracle.prev_d2para(T(g["depth"]), rot, trans, cam).numpy(), g["prev_d2para"])
    assert np.array_equal(oracle.parallax2depth(T(g["para"]), rot, trans, cam).numpy(), g["parallax2depth"])
    assert np.array_equal(oracle.depth2parallax(T(g["depth"]), rot, trans, cam).numpy(), g["depth2parallax"])
    assert np.array_equal(oracle.dense_image_warp(T(g["img"]), T(g["flow"]), False).numpy(), g["warp_gather"])
    assert np.array_equal(oracle.dense_image_warp(T(g["img"]), T(g["flow"]), True).numpy(), g["warp_bp"])


def test_domain_normalization_matches_reference(golden_dir):
    g = load(golden_dir, "dn.npz")
    out = oracle.DomainNormalization(T(g["scale"]), T(g["bias"]))(T(g["x"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=1e-6, atol=1e-7)


def check_depth(got, want, frame):
    """Model-level tolerance (DESIGN.md "Tolerances").

    depth = (s/rho - tz)/alpha cancels near depth 0, so errors are measured against |depth| + 0.1.
    Frames 0/1: every pixel within 1e-4.  Later frames: an fp32 summation-order difference of 1e-7
    moves a query point enough to flip an fp16 rounding inside the PSCV (1 fp16 ulp = 5e-4 relative)
    and the flip propagates through the refiner, so the bound is statistical: median 1e-5,
    99th percentile 1e-3, max 1e-2.  Two CPU evaluations of the reference's own graph that differ
    only in summation order show exactly this: test_summation_order_alone_moves_the_depth_maps below.
    """
    err = np.abs(got - want) / (np.abs(want) + 0.1)
    if frame <= 1:
        assert err.max() <= 1e-4, err.max()
    else:
        assert np.median(err) <= 1e-5 and np.percentile(err, 99) <= 1e-3 and err.max() <= 1e-2, \
            (np.median(err), np.percentile(err, 99), err.max())


@pytest.mark.parametrize("case", ["cfg1", "cfg1_bp", "odd"])
def test_model_matches_reference(golden_dir, case):
    g = load(golden_dir, f"model_{case}.npz")
    nl = int(g["nbre_levels"])
    w = oracle.init_weights(nl, seed=int(g["weights_seed"]), bias_std=0.05, dn_random=True)
    model = oracle.M4Depth(w, nbre_levels=nl, pscv_kwargs={"use_cuda_backproject": bool(g["backproject"])})
    cam = cam_of(g)
    t = 0
    while f"rgb_{t}" in g:
        b = g[f"rgb_{t}"].shape[0]
        sample = {"RGB_im": T(g[f"rgb_{t}"]), "rot": T(g[f"rot_{t}"]), "trans": T(g[f"trans_{t}"]),
                  "new_traj": torch.tensor([t == 0] * b)}
        out = model([[sample], cam])
        check_depth(out["depth"].numpy(), g[f"depth_{t}"], t)
        for li, lvl in enumerate(model.d_estimator.levels):
            check_depth(lvl.depth_prev_t.numpy(), g[f"state_depth_{t}_l{li + 1}"], t)
        t += 1
    assert t >= 2


def test_oracle_sqrt_is_correctly_rounded():
    """oracle._ieee.sqrt (numpy / hardware sqrtps) equals the float64 route, which is correctly rounded for fp32."""
    from oracle._ieee import sqrt
    g = torch.Generator().manual_seed(0)
    x = torch.exp(torch.rand(1 << 18, generator=g) * 40 - 20)
    want = torch.from_numpy(np.sqrt(x.numpy().astype(np.float64)).astype(np.float32))
    assert torch.equal(sqrt(x), want)


def test_back_project_grad_restatement_matches_autograd_of_the_forward():
    """The oracle's BackProjectGrad (restated from backproject_op_gpu.cu.cc:108-196) equals the autograd gradient of the
    oracle's BackProject forward (fp64, away from integer coordinates where the bilinear kernel is not differentiable)."""
    g = torch.Generator().manual_seed(3)
    B, H, W, S, Fd, C = 2, 6, 7, 3, 2, 5
    inp = torch.randn(B, H, W, Fd, C, generator=g)
    frac = 0.1 + 0.8 * torch.rand(B, H, W, S, Fd, 2, generator=g)
    base = torch.stack((torch.randint(-1, W, (B, H, W, S, Fd), generator=g), torch.randint(-1, H, (B, H, W, S, Fd), generator=g)), -1)
    coords = (base + frac).to(torch.float32)               # some samples fall outside the image: zero gradient there
    grad = torch.randn(B, H, W, S, Fd, C, generator=g)
    ig, cg = oracle.back_project_grad(inp, coords, grad)

    def fwd64(i, c):
        x, y = c[..., 0], c[..., 1]
        inside = (x >= 0) & (y >= 0) & (x <= W - 1) & (y <= H - 1)
        x0, y0 = torch.floor(x).clamp(0, W - 1).long(), torch.floor(y).clamp(0, H - 1).long()
        x1, y1 = torch.ceil(x).clamp(0, W - 1).long(), torch.ceil(y).clamp(0, H - 1).long()
        dx, dy = (x - x0).unsqueeze(-1), (y - y0).unsqueeze(-1)
        bi = torch.arange(B).view(B, 1, 1, 1, 1).expand_as(x0)
        fi = torch.arange(Fd).view(1, 1, 1, 1, Fd).expand_as(x0)
        tap = lambda yy, xx: i[bi, yy, xx, fi]
        out = tap(y0, x0) * (1 - dy) * (1 - dx) + tap(y0, x1) * (1 - dy) * dx + tap(y1, x0) * dy * (1 - dx) + tap(y1, x1) * dy * dx
        return out * inside.unsqueeze(-1)

    i64 = inp.double().requires_grad_(True)
    c64 = coords.double().requires_grad_(True)
    (fwd64(i64, c64) * grad.double()).sum().backward()
    np.testing.assert_allclose(ig.numpy(), i64.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cg.numpy(), c64.grad.numpy(), rtol=1e-4, atol=1e-4)
    # and the forward used for the autograd check is the oracle's forward
    np.testing.assert_allclose(oracle.back_project(inp, coords).numpy(), fwd64(inp.double(), coords.double()).detach().numpy(), atol=1e-5)


def test_plain_c_restatement_of_the_backproject_kernels_agrees_with_the_torch_oracle():
    """oracle/m4d_oracle.c restates the reference's two BackProject kernel bodies (backproject_op_gpu.cu.cc:19-79, 108-196)
    as plain C; the torch oracle must agree with it: forward and tap grids bit for bit (same expression, separately rounded),
    the gradients to summation-order tolerance."""
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle")], check=True)
    lib = ctypes.CDLL(os.path.join(root, "oracle", "libm4d_oracle.so"))
    g = torch.Generator().manual_seed(17)
    B, H, W, S, Fd, C = 2, 7, 9, 3, 2, 5
    inp = torch.randn(B, H, W, Fd, C, generator=g)
    coords = torch.rand(B, H, W, S, Fd, 2, generator=g) * torch.tensor([W + 2.0, H + 2.0]) - 1.0
    coords[0, 0, 0, 0, 0, 0] = float("nan")
    coords[0, 1, 1, 0, 0] = torch.tensor([3.0, 2.0])
    coords[1, 2, 2, 1, 1] = torch.tensor([W - 1.0, H - 1.0])
    grad = torch.randn(B, H, W, S, Fd, C, generator=g)
    dim = (ctypes.c_int32 * 6)(B, H, W, S, Fd, C)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    out = torch.empty(B, H, W, S, Fd, C)
    idx = torch.empty(B, H, W, S, Fd, 4, dtype=torch.int32)
    lib.bp_forward(ptr(inp), ptr(coords), dim, ptr(out), ptr(idx))
    assert torch.equal(out, oracle.back_project(inp, coords))
    x0, x1, y0, y1, _ = oracle.back_project_index_grids(coords, H, W)
    assert torch.equal(idx, torch.stack((x0, x1, y0, y1), dim=-1))
    ig, cg = torch.empty_like(inp), torch.empty_like(coords)
    lib.bp_backward(ptr(grad), ptr(inp), ptr(coords), dim, ptr(ig), ptr(cg))
    want_i, want_c = oracle.back_project_grad(inp, coords, grad)
    np.testing.assert_allclose(ig.numpy(), want_i.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(cg.numpy(), want_c.numpy(), rtol=1e-5, atol=1e-5)


def depth_err_quantiles(x, y, qs=(0.5, 0.9, 0.99, 0.999)):
    """Relative depth error |x - y| / (|y| + 0.1): the quantiles in ``qs``, the maximum and the fraction within 1e-4."""
    err = ((x - y).abs() / (y.abs() + 0.1)).flatten().double()
    return [float(torch.quantile(err, q)) for q in qs], float(err.max()), float((err <= 1e-4).double().mean())


def run_oracle_sequence(frames, cam, nl, mode, weights_seed=1):
    """The oracle model over a sequence under one summation order (oracle.network.REDUCTION_ORDER); depth map per frame."""
    w = oracle.init_weights(nl, seed=weights_seed, bias_std=0.05, dn_random=True)
    outs = []
    with oracle.reduction_order(mode), torch.no_grad():
        model = oracle.M4Depth(w, nbre_levels=nl, pscv_kwargs={"use_cuda_backproject": False})
        for t, fr in enumerate(frames):
            s = dict(fr)
            s["new_traj"] = torch.tensor([t == 0] * fr["RGB_im"].shape[0])
            outs.append(model([[s], cam])["depth"].clone())
    return outs


def test_summation_order_alone_moves_the_depth_maps():
    """What the north-star tolerance (1e-4 relative) can and cannot mean for the recurrent model.  The SAME reference graph
    evaluated twice on the CPU in fp32 - convolutions and DomainNormalization means summed in two different orders, and once
    more with those reductions in fp64 - agrees to 1e-4 on the first estimated frame, and then diverges: an fp32 rounding
    difference flips an fp16 rounding in the PSCV (1 fp16 ulp = 5e-4), the refiner and exp() amplify it, and the recurrent
    state carries it on.  By the third estimated frame the 99th percentile of the error BETWEEN TWO CPU EVALUATIONS is beyond
    1e-3.  The GPU whole-model tests therefore bound the GPU-vs-oracle error by this oracle-vs-oracle error
    (tests/test_gpu_parity.py::test_model_vs_oracle_at_baseline_configs), not by a fixed number."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    from synth import synth_sequence
    frames, cam = synth_sequence(5, 1, 128, 192, "kitti", seed=77)
    ref = run_oracle_sequence(frames, cam, 6, "default")
    alt = run_oracle_sequence(frames, cam, 6, "reordered")
    f64 = run_oracle_sequence(frames, cam, 6, "fp64")
    assert torch.equal(ref[0], alt[0]) and float(ref[0].min()) == 1000.0          # frame 0: new-trajectory pass-through
    q1, mx1, in1 = depth_err_quantiles(alt[1], ref[1])
    assert mx1 <= 1e-4 and in1 == 1.0, (q1, mx1)                                  # first estimated frame: 1e-4 everywhere
    grew = []
    for t in range(2, 5):
        qa, mxa, ina = depth_err_quantiles(alt[t], ref[t])
        qb, mxb, inb = depth_err_quantiles(f64[t], ref[t])
        grew.append((qa[2], qb[2], ina))
        assert qa[0] <= 1e-4 and qb[0] <= 1e-4                                     # the bulk stays close ...
    assert grew[-1][0] > 1e-4 and grew[-1][1] > 1e-4 and grew[-1][2] < 1.0         # ... the tail does not, whatever the order
    assert grew[-1][0] > grew[0][0]                                                # and it grows with the recurrence

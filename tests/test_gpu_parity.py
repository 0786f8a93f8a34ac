"""GPU parity tests: libm4d (through the C ABI via the Python mirror) vs the CPU oracle and the golden vectors.

Tolerances (DESIGN.md "Numerics"):
  * integer tap grids of the BackProject convention: bit-exact;
  * PSCV cv (fp16-quantised) and prev_disp, geometry maps, warps: bit-exact against the oracle branch that uses the
    same bilinear convention (the kernels evaluate the reference's op sequence with one rounding per op);
  * fp32 reductions whose order is not defined by the reference (SNCV mean, DN statistics, convolutions):
    1e-5 relative to the tensor scale (north_star allows 1e-4);
  * whole model: the statistical bound of tests/test_oracle_golden.py::check_depth.
"""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

T = torch.from_numpy


def _m4d():
    import m4depth_b200
    return m4depth_b200


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def cu(x):
    return (T(x) if isinstance(x, np.ndarray) else x).cuda().contiguous()


def cam_of(g, dev=True):
    f, c = T(g["cam_f"]), T(g["cam_c"])
    return {"f": f.cuda(), "c": c.cuda()} if dev else {"f": f, "c": c}


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


def pscv_inputs(seed, b, h, w, c, cuts, kind):
    g = torch.Generator().manual_seed(seed)
    cam = camera_for(kind, b, h, w)
    rot, trans = motion(g, b)
    lk = lambda t: torch.where(t >= 0, t, 0.1 * t)
    c1 = oracle.group_l2_normalize(lk(torch.randn(b, h, w, c, generator=g)), cuts)
    c2 = oracle.group_l2_normalize(lk(torch.randn(b, h, w, c, generator=g)), cuts)
    para_l = torch.exp(torch.rand(b, h, w, 1, generator=g) * (np.log(16) - np.log(0.5)) + np.log(0.5))
    para_t = torch.exp(torch.rand(b, h, w, 1, generator=g) * (np.log(16) - np.log(0.05)) + np.log(0.05))
    return c1, c2, para_t, para_l, rot, trans, cam


def dev_cam(cam):
    return {"f": cam["f"].cuda(), "c": cam["c"].cuda()}


# ------------------------------------------------------------------------------------------ library
def test_library_loaded_and_counts_launches():
    m = _m4d()
    assert m._lib.ABI_VERSION == 1
    n0 = m.launch_count()
    x = torch.zeros(8, device="cuda")
    m._lib.check(m._lib.lib.m4d_fill(x.data_ptr(), 8, 3.0, m._lib.stream()))
    torch.cuda.synchronize()
    assert m.launch_count() == n0 + 1
    assert torch.all(x == 3.0)


def test_errors_are_returned_not_fatal():
    m = _m4d()
    rc = m._lib.lib.m4d_fill(None, 8, 1.0, None)
    assert rc == -1 and "null" in m._lib.last_error()
    with pytest.raises(m.M4DError):
        m.utils.cost_volume(torch.zeros(1, 4, 4, 6, device="cuda"), torch.zeros(1, 4, 4, 6, device="cuda"), 3, nbre_cuts=1)
    with pytest.raises(m.M4DError):
        m.utils.cost_volume(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8), 3)      # CPU tensors: no fallback
    with pytest.raises(ValueError):
        m.utils.get_rot_mat(torch.zeros(2, 5, device="cuda"))


# ------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("case", ["l2_kitti", "l1_midair", "l4_tartan", "l6_kitti"])
@pytest.mark.parametrize("branch", ["gather", "bp"])
def test_pscv_golden(golden_dir, case, branch):
    m = _m4d()
    g = load(golden_dir, f"pscv_{case}.npz")
    interp = m.INTERP_GATHER if branch == "gather" else m.INTERP_BP
    cv, pd = m.utils.get_parallax_sweeping_cv(cu(g["c1"]), cu(g["c2"]), cu(g["para_prev_t"]), cu(g["para_prev_l"]),
                                              cu(g["rot"]), cu(g["trans"]), cam_of(g), 4, nbre_cuts=int(g["cuts"]),
                                              interp=interp)
    assert np.array_equal(pd.cpu().numpy(), g["prev_disp_" + branch])
    assert np.array_equal(cv.cpu().numpy(), g["cv_" + branch])


@pytest.mark.parametrize("shape", [(2, 24, 80, 32, 2), (1, 15, 20, 128, 4), (2, 8, 10, 192, 8), (1, 30, 41, 96, 4),
                                   (1, 37, 53, 16, 1), (1, 9, 33, 64, 2), (1, 12, 16, 96, 3), (1, 7, 9, 64, 1)])
@pytest.mark.parametrize("interp", [0, 1, 2])
def test_pscv_specialised_kernel_equals_generic_kernel(shape, interp):
    """The K=9 kernels - warp-autonomous (the network's (c, cuts) pairs) and CTA-tile (2-D tiles, partial tiles at the
    borders, every pyramid channel count, odd group counts) - are bit-identical to the shape-generic kernel, including prev_disp, the tap grids and the fused
    log(centre) output the pipeline consumes."""
    m = _m4d()
    L = m._lib
    b, h, w, c, cuts = shape
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(100 + h + c, b, h, w, c, cuts, "kitti")
    pl[:, : h // 3] -= 6.0                     # negative parallax + k: clipped hypotheses; large ones leave the image
    pl[:, -2:] *= 40.0
    args = [cu(x) for x in (c1, c2, pt, pl, rot, trans)]
    dc = dev_cam(cam)
    res = []
    for flag in (0, L.INTERP_FLAG_TILE, L.INTERP_FLAG_GENERIC):
        cv, pd, idx = m.utils.get_parallax_sweeping_cv(*args, dc, 4, nbre_cuts=cuts, interp=interp | flag, return_index_grids=True)
        xs = cuts * 9 + 3
        wide = torch.full((b, h, w, xs), -7.0, device="cuda")
        L.check(L.lib.m4d_pscv_fused_fwd_ex(
            L.ptr(args[0]), L.ptr(args[1]), L.ptr(args[2]), L.ptr(args[3]), L.ptr(args[4]), 4, L.ptr(args[5]), L.ptr(dc["f"]),
            L.ptr(dc["c"]), b, h, w, c, cuts, 4, wide.data_ptr() + 4, xs, None, 0, wide.data_ptr() + 4 * (xs - 1), xs, 0.5,
            None, interp | flag, L.stream()))
        res.append((cv.cpu(), pd.cpu(), idx.cpu(), wide.cpu()))
    for other in res[1:]:
        for a_, b_ in zip(res[0], other):
            assert torch.equal(a_.view(torch.int32) if a_.dtype == torch.float32 else a_,
                               b_.view(torch.int32) if b_.dtype == torch.float32 else b_)
    cv, pd, _, wide = res[0]
    assert torch.equal(wide[..., 1:1 + cuts * 9], cv) and torch.all(wide[..., 0] == -7.0) and torch.all(wide[..., cuts * 9 + 1] == -7.0)
    want = torch.log(pd[..., 4] * 0.5)
    got = wide[..., -1]
    ok = torch.isfinite(want)
    assert torch.equal(torch.isfinite(got), ok) and torch.allclose(got[ok], want[ok], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("case", ["l2", "l6"])
def test_sncv_golden(golden_dir, case):
    m = _m4d()
    g = load(golden_dir, f"sncv_{case}.npz")
    f = cu(g["f"])
    out = m.utils.cost_volume(f, f, 3, nbre_cuts=int(g["cuts"]))
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", ["kitti", "tartan"])
def test_geometry_golden(golden_dir, case):
    m = _m4d()
    g = load(golden_dir, f"geom_{case}.npz")
    cam = cam_of(g)
    rot, trans = cu(g["rot"]), cu(g["trans"])
    assert np.array_equal(m.utils.get_rot_mat(rot).cpu().numpy(), g["rot_mat"])
    assert np.array_equal(m.utils.prev_d2para(cu(g["depth"]), rot, trans, cam).cpu().numpy(), g["prev_d2para"])
    assert np.array_equal(m.utils.parallax2depth(cu(g["para"]), rot, trans, cam).cpu().numpy(), g["parallax2depth"])
    assert np.array_equal(m.utils.depth2parallax(cu(g["depth"]), rot, trans, cam).cpu().numpy(), g["depth2parallax"])
    # dense_image_warp: the library implements the BackProject branch with the compiled kernel's FMA contraction;
    # against the separately-rounded restatement that is <= 2 ulp of the largest term
    w = m.utils.dense_image_warp(cu(g["img"]), cu(g["flow"])).cpu().numpy()
    np.testing.assert_allclose(w, g["warp_bp"], rtol=0, atol=4e-7 * np.abs(g["img"]).max())
    np.testing.assert_allclose(w, g["warp_gather"], rtol=0, atol=2e-6 * np.abs(g["img"]).max())


def test_domain_normalization_golden(golden_dir):
    m = _m4d()
    g = load(golden_dir, "dn.npz")
    dn = m.DomainNormalization()
    dn.scale, dn.bias = cu(g["scale"]), cu(g["bias"])
    out = dn(cu(g["x"]))
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-5, atol=1e-6)


def check_depth(got, want, frame):
    """Frame 0 (new-trajectory pass-through) must match to 1e-4 everywhere.  From frame 1 on the fp16 stage of the
    PSCV amplifies legitimate 1e-7 fp32 summation-order differences (convolutions, normalisations) into isolated
    fp16-ulp flips, so the bound is the statistical one of tests/test_oracle_golden.py (DESIGN.md "Numerics")."""
    err = np.abs(got - want) / (np.abs(want) + 0.1)
    if frame == 0:
        assert err.max() <= 1e-4, err.max()
    else:
        assert np.median(err) <= 1e-5 and np.percentile(err, 99) <= 1e-3 and err.max() <= 1e-2, \
            (np.median(err), np.percentile(err, 99), err.max())


@pytest.mark.parametrize("case", ["cfg1", "cfg1_bp", "odd"])
@pytest.mark.parametrize("graph", [False, True])
def test_model_golden(golden_dir, case, graph):
    """BASELINE config 1 (128x128, 3 levels, 2 frames) and an odd-sized 6-level case, frame by frame."""
    m = _m4d()
    g = load(golden_dir, f"model_{case}.npz")
    nl = int(g["nbre_levels"])
    model = m.M4Depth(nbre_levels=nl, use_cuda_graph=graph)
    model.load_weights(oracle.init_weights(nl, seed=int(g["weights_seed"]), bias_std=0.05, dn_random=True))
    model.set_interp(m.INTERP_BP if bool(g["backproject"]) else m.INTERP_GATHER)
    cam = cam_of(g)
    t = 0
    while f"rgb_{t}" in g:
        b = g[f"rgb_{t}"].shape[0]
        sample = {"RGB_im": cu(g[f"rgb_{t}"]), "rot": cu(g[f"rot_{t}"]), "trans": cu(g[f"trans_{t}"]),
                  "new_traj": [t == 0] * b}
        out = model([[sample], cam])
        check_depth(out["depth"].cpu().numpy(), g[f"depth_{t}"], t)
        for li, lvl in enumerate(model.d_estimator.levels):
            check_depth(lvl.depth_prev_t.cpu().numpy(), g[f"state_depth_{t}_l{li + 1}"], t)
        t += 1
    assert t >= 2


# ------------------------------------------------------------------- seeded inputs vs the oracle, all level shapes
LEVEL_SHAPES = [  # (name, b, h, w, c, cuts, camera kind): the level shapes of BASELINE configs 1-5
    ("cfg1_l1", 1, 64, 64, 16, 1, "midair"), ("cfg1_l3", 1, 16, 16, 64, 2, "midair"),
    ("cfg2_l2", 1, 96, 96, 32, 2, "midair"), ("cfg2_l4", 1, 24, 24, 96, 4, "midair"),
    ("cfg3_l3", 2, 48, 160, 64, 2, "kitti"), ("cfg3_l5", 2, 12, 40, 128, 4, "kitti"),
    ("cfg3_l6", 8, 6, 20, 192, 8, "kitti"), ("cfg5_l6", 2, 8, 10, 192, 8, "tartan"),
    ("cfg5_l5", 1, 15, 20, 128, 4, "tartan"), ("ragged", 3, 5, 7, 32, 2, "kitti"),
    ("cfg3_l1", 1, 192, 640, 16, 1, "kitti"), ("cfg3_l2", 1, 96, 320, 32, 2, "kitti"), ("cfg5_l1", 1, 240, 320, 16, 1, "tartan"),
    ("cfg5_l2", 1, 120, 160, 32, 2, "tartan"), ("cfg5_l3", 1, 60, 80, 64, 2, "tartan"), ("cfg5_l4", 1, 30, 40, 96, 4, "tartan"),
]


@pytest.mark.parametrize("shape", LEVEL_SHAPES, ids=[s[0] for s in LEVEL_SHAPES])
@pytest.mark.parametrize("interp", ["gather", "bp"])
def test_pscv_vs_oracle(shape, interp):
    m = _m4d()
    _, b, h, w, c, cuts, kind = shape
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(100 + h * w + c, b, h, w, c, cuts, kind)
    want_cv, want_pd = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts,
                                                       use_cuda_backproject=(interp == "bp"))
    cv, pd, idx = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4,
                                                   nbre_cuts=cuts, interp=m.INTERP_BP if interp == "bp" else m.INTERP_GATHER,
                                                   return_index_grids=True)
    # integer index grids of the backproject op: bit-exact
    qy, qx = oracle.pscv_query_points(pl, rot, trans, cam, 4)
    qx = torch.minimum(torch.maximum(qx, torch.zeros(())), torch.tensor(float(w - 1)))
    qy = torch.minimum(torch.maximum(qy, torch.zeros(())), torch.tensor(float(h - 1)))
    x0, x1, y0, y1, _ = oracle.back_project_index_grids(torch.stack((qx, qy), dim=-1), h, w)
    want_idx = torch.stack((x0, x1, y0, y1), dim=-1).permute(1, 2, 3, 0, 4)      # [K,b,h,w,4] -> [b,h,w,K,4]
    assert torch.equal(idx.cpu(), want_idx)
    assert torch.equal(pd.cpu(), want_pd)
    assert torch.equal(cv.cpu(), want_cv)


def test_pscv_bp_fma_close_to_bp():
    """BP_FMA (the reference GPU binary's contraction) vs separately rounded BP: fp32 values within 2 ulp of the
    largest term, fp16-quantised cv within one fp16 ulp and equal almost everywhere."""
    m = _m4d()
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(7, 2, 24, 80, 96, 4, "kitti")
    args = (cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4)
    cv_a, pd_a = m.utils.get_parallax_sweeping_cv(*args, nbre_cuts=4, interp=m.INTERP_BP)
    cv_b, pd_b = m.utils.get_parallax_sweeping_cv(*args, nbre_cuts=4, interp=m.INTERP_BP_FMA)
    assert torch.allclose(pd_a, pd_b, rtol=0, atol=4e-7 * float(pt.max()))
    d = (cv_a - cv_b).abs()
    assert float(d.max()) <= 2.0 ** -11 * float(cv_a.abs().max())
    assert float((d > 0).float().mean()) < 0.02


def test_pscv_unnormalised_features_within_one_fp16_ulp():
    """Features that are NOT group-normalised (ablation): partial sums are no longer exact in fp32, so the kernel's
    summation tree may differ from the oracle's channel-order sum by one rounding of the final fp16 value."""
    m = _m4d()
    g = torch.Generator().manual_seed(3)
    b, h, w, c, cuts = 1, 12, 40, 32, 2
    c1 = torch.randn(b, h, w, c, generator=g) * 3
    c2 = torch.randn(b, h, w, c, generator=g) * 3
    _, _, pt, pl, rot, trans, cam = pscv_inputs(4, b, h, w, c, cuts, "kitti")
    want_cv, _ = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts, use_cuda_backproject=False)
    cv, _ = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4, nbre_cuts=cuts)
    d = (cv.cpu() - want_cv).abs()
    assert float((d / want_cv.abs().clamp(min=2.0 ** -14)).max()) <= 2.0 ** -10
    assert float((d > 0).float().mean()) < 0.01


def test_pscv_strided_outputs_and_centre_log():
    """The fused-pipeline form: cv written into a wider refiner-input buffer, only log(centre prev_disp * scale) kept."""
    m = _m4d()
    L = m._lib
    b, h, w, c, cuts = 2, 12, 40, 32, 2
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(9, b, h, w, c, cuts, "kitti")
    dc = dev_cam(cam)
    cv, pd = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dc, 4, nbre_cuts=cuts)
    xs = 124
    x_in = torch.full((b, h, w, xs), -5.0, device="cuda")
    d = [cu(t) for t in (c1, c2, pt, pl, rot, trans)]
    L.check(L.lib.m4d_pscv_fused_fwd(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), 4,
                                     d[5].data_ptr(), dc["f"].data_ptr(), dc["c"].data_ptr(), b, h, w, c, cuts, 4,
                                     x_in.data_ptr() + 4 * 3, xs, None, 0, x_in.data_ptr() + 4 * 121, xs, 0.5, None, L.stream()))
    assert torch.equal(x_in[..., 3:21], cv)
    assert torch.allclose(x_in[..., 121], torch.log(pd[..., 4] * 0.5), rtol=1e-6, atol=1e-6)
    untouched = torch.ones(xs, dtype=torch.bool)
    untouched[3:21] = False
    untouched[121] = False
    assert torch.all(x_in[..., untouched.cuda()] == -5.0)


def test_pscv_edge_cases():
    """Parallax far outside the image (clamped taps), negative hypotheses (collapse onto p), tiny maps."""
    m = _m4d()
    b, h, w, c, cuts = 2, 2, 3, 32, 2
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(21, b, h, w, c, cuts, "tartan")
    pl = pl * 0 + torch.tensor([0.3, 500.0]).view(2, 1, 1, 1)
    for interp, flag in ((m.INTERP_GATHER, False), (m.INTERP_BP, True)):
        want_cv, want_pd = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts, use_cuda_backproject=flag)
        cv, pd = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4,
                                                  nbre_cuts=cuts, interp=interp)
        assert torch.equal(cv.cpu(), want_cv) and torch.equal(pd.cpu(), want_pd)


def test_pscv_and_sncv_kernels_agree_on_random_shapes():
    """Seeded fuzz: on 40 random image sizes (partial 8x4 warp tiles, partial 16x8 super tiles, images smaller than a tile)
    and the network's (c, cuts) pairs the specialised PSCV kernels equal the shape-generic one bit for bit (cv, prev_disp, tap
    grids), and the column-strip SNCV kernel equals the (pixel, dy) kernel."""
    m = _m4d()
    L = m._lib
    rng = np.random.default_rng(7)
    pairs = [(16, 1), (32, 2), (64, 2), (96, 4), (128, 4), (192, 8)]
    for case in range(40):
        c, cuts = pairs[int(rng.integers(0, len(pairs)))]
        b, h, w = int(rng.integers(1, 4)), int(rng.integers(2, 45)), int(rng.integers(2, 70))
        interp = int(rng.integers(0, 3))
        c1, c2, pt, pl, rot, trans, cam = pscv_inputs(1000 + case, b, h, w, c, cuts, ["kitti", "midair", "tartan"][case % 3])
        args = [cu(t) for t in (c1, c2, pt, pl, rot, trans)]
        dc = dev_cam(cam)
        ref = m.utils.get_parallax_sweeping_cv(*args, dc, 4, nbre_cuts=cuts, interp=interp | L.INTERP_FLAG_GENERIC, return_index_grids=True)
        got = m.utils.get_parallax_sweeping_cv(*args, dc, 4, nbre_cuts=cuts, interp=interp, return_index_grids=True)
        for a_, b_ in zip(got, ref):
            assert torch.equal(a_.view(torch.int32) if a_.dtype == torch.float32 else a_,
                               b_.view(torch.int32) if b_.dtype == torch.float32 else b_), (case, b, h, w, c, cuts, interp)
        if cuts <= 4:
            oc = 49 * cuts
            outs = []
            for variant in (2, 1):
                o = torch.empty(b, h, w, oc, device="cuda")
                L.check(L.lib.m4d_sncv_fwd_ex(L.ptr(args[0]), L.ptr(args[1]), b, h, w, c, cuts, 3, L.ptr(o), oc, variant, L.stream()))
                outs.append(o)
            assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32)), (case, b, h, w, c, cuts)


def test_fast_division_is_ieee():
    """Phase 0 of pscv9s_kernel divides with MUFU.RCP + five FFMA and no branch (csrc/pscv_smem.cu div_fast): wherever it does
    not flag its operands as unsafe the quotient must be the IEEE round-to-nearest one, bit for bit - random mantissas over
    the whole safe exponent range, operands straddling the range limits, exact and near-exact quotients, the values the
    geometry produces (s / 1e-6, d / 1e8, small integers)."""
    m = _m4d()
    L = m._lib
    g = torch.Generator().manual_seed(11)
    n = 1 << 22
    def rnd(lo, hi):
        mant = 1.0 + torch.rand(n, generator=g, dtype=torch.float64)
        e = torch.randint(lo, hi + 1, (n,), generator=g).to(torch.float64)
        sgn = torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0).to(torch.float64)
        return (sgn * mant * torch.pow(torch.tensor(2.0, dtype=torch.float64), e)).to(torch.float32)
    cases = [(rnd(-60, 59), rnd(-60, 59)), (rnd(-70, 70), rnd(-70, 70)), (rnd(-20, 20), rnd(-2, 2)),
             (rnd(-126, 127), rnd(-126, 127))]
    k = torch.arange(1, n + 1, dtype=torch.float32)
    cases.append((k, torch.full((n,), 3.0)))
    cases.append((torch.rand(n, generator=g) * 300.0 + 1e-3, torch.full((n,), 1e-6)))
    cases.append((torch.randn(n, generator=g) * 200.0, torch.rand(n, generator=g) * 3e8 + 1e6))
    special = torch.tensor([0.0, -0.0, float("inf"), float("nan"), 1e-45, 1.17549435e-38, 3.4e38, 1.0, 2.0 ** -60, 2.0 ** 60,
                            2.0 ** -61, 2.0 ** 61])
    cases.append((special.repeat_interleave(len(special)), special.repeat(len(special))))
    safe_total = 0
    for a, b in cases:
        a, b = a.cuda().contiguous(), b.cuda().contiguous()
        nn = a.numel()
        qf, qi = torch.empty(nn, device="cuda"), torch.empty(nn, device="cuda")
        fl = torch.empty(nn, dtype=torch.int32, device="cuda")
        L.check(L.lib.m4d_debug_div_check(L.ptr(a), L.ptr(b), nn, L.ptr(qf), L.ptr(qi), L.ptr(fl), L.stream()))
        safe = fl == 0
        assert torch.equal(qf[safe].view(torch.int32), qi[safe].view(torch.int32))
        # the IEEE reference itself: fp64 quotient rounded once to fp32 (exact for fp32 operands up to double rounding, which
        # cannot occur for a 24-bit by 24-bit quotient in 53 bits)
        want = (a.double() / b.double()).float()
        ok = safe & torch.isfinite(want) & (want.abs() > 1e-37)
        assert torch.equal(qi[ok].view(torch.int32), want[ok].view(torch.int32))
        ea, eb = (a.view(torch.int32) >> 23) & 0xFF, (b.view(torch.int32) >> 23) & 0xFF
        inrange = (ea >= 67) & (ea <= 187) & (eb >= 67) & (eb <= 187)
        assert torch.equal(safe, inrange)
        safe_total += int(safe.sum())
    assert safe_total > 4 * n


@pytest.mark.parametrize("shape", [(8, 96, 320), (2, 96, 96), (3, 37, 53), (1, 120, 160), (2, 9, 17), (1, 2, 2)])
@pytest.mark.parametrize("data", ["micro", "insitu", "far", "nan"])
def test_pscv_smem_staged_kernel_equals_ldg_kernels(shape, data):
    """pscv9s_kernel (c2 window staged in shared memory by bulk copies, csrc/pscv_smem.cu) against the warp-autonomous LDG
    kernel and, at small sizes, the shape-generic one: cv, prev_disp, the integer tap grids and the fused log(centre) output
    bit for bit.  Data: the microbench distribution (boxes of 300-800 pixels), in-situ-like small parallax (3-4 of the 9
    hypotheses collapse onto one point: the duplicate-record skip), parallax far beyond the window (tiles that overflow
    the window buffer take the LDG taps inside the same kernel; taps clamped at the image border), NaN / Inf parallax
    (records without taps, tiles without any tap)."""
    m = _m4d()
    L = m._lib
    b, h, w = shape
    c, cuts = 32, 2
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(500 + h + w, b, h, w, c, cuts, "kitti")
    g = torch.Generator().manual_seed(h * w)
    if data == "insitu":
        pl = 0.6 + 1.7 * torch.rand(b, h, w, 1, generator=g)
    elif data == "far":
        pl = pl * torch.where(torch.rand(b, h, w, 1, generator=g) < 0.02, 30.0, 1.0)
        pl[:, : max(1, h // 4)] *= -1.0
    elif data == "nan":
        pl[0, : min(h, 17)] = float("nan")
        pl[-1, -1, -1] = float("inf")
        pl[torch.rand(b, h, w, 1, generator=g) < 0.05] = float("nan")
    args = [cu(x) for x in (c1, c2, pt, pl, rot, trans)]
    dc = dev_cam(cam)
    flags = (0, L.INTERP_VARIANT_STAGED_8, L.INTERP_VARIANT_STAGED_4, L.INTERP_FLAG_WARP) + (
        (L.INTERP_FLAG_GENERIC,) if b * h * w <= 20000 else ())
    res = []
    for flag in flags:
        cv, pd, idx = m.utils.get_parallax_sweeping_cv(*args, dc, 4, nbre_cuts=cuts, interp=flag, return_index_grids=True)
        xs = 124
        wide = torch.full((b, h, w, xs), -7.0, device="cuda")
        L.check(L.lib.m4d_pscv_fused_fwd_ex(
            L.ptr(args[0]), L.ptr(args[1]), L.ptr(args[2]), L.ptr(args[3]), L.ptr(args[4]), 4, L.ptr(args[5]), L.ptr(dc["f"]),
            L.ptr(dc["c"]), b, h, w, c, cuts, 4, wide.data_ptr(), xs, None, 0, wide.data_ptr() + 4 * 121, xs, 0.5,
            None, flag, L.stream()))
        res.append((cv.cpu(), pd.cpu(), idx.cpu(), wide.cpu()))
    bits = lambda t: t.view(torch.int32) if t.dtype == torch.float32 else t
    for other in res[1:]:
        for a_, b_ in zip(res[0], other):
            assert torch.equal(bits(a_), bits(b_))
    cv, _, _, wide = res[0]
    assert torch.equal(bits(wide[..., :18]), bits(cv)) and torch.all(wide[..., 18:121] == -7.0) and torch.all(wide[..., 122:] == -7.0)


@pytest.mark.parametrize("shape", [(8, 192, 640), (2, 96, 96), (3, 37, 53), (1, 120, 160), (2, 9, 17), (1, 2, 2)])
@pytest.mark.parametrize("data", ["micro", "insitu", "far", "nan"])
def test_pscv_smem_staged_kernel_level1_shape_equals_ldg_kernels(shape, data):
    """The staged kernel instantiated for level 1 (c = 16, one group: 32x8-pixel tiles, 64-byte pixel rows, quad rotation by pixel
    pairs) against the warp-autonomous LDG kernel that level runs by default and, at small sizes, the shape-generic one: cv,
    prev_disp, integer tap grids and the fused log(centre) output bit for bit, on the same four parallax distributions."""
    m = _m4d()
    L = m._lib
    b, h, w = shape
    c, cuts = 16, 1
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(700 + h + w, b, h, w, c, cuts, "kitti")
    g = torch.Generator().manual_seed(h * w + 1)
    if data == "insitu":
        pl = 0.6 + 1.7 * torch.rand(b, h, w, 1, generator=g)
    elif data == "far":
        pl = pl * torch.where(torch.rand(b, h, w, 1, generator=g) < 0.02, 30.0, 1.0)
        pl[:, : max(1, h // 4)] *= -1.0
    elif data == "nan":
        pl[0, : min(h, 17)] = float("nan")
        pl[-1, -1, -1] = float("inf")
        pl[torch.rand(b, h, w, 1, generator=g) < 0.05] = float("nan")
    args = [cu(x) for x in (c1, c2, pt, pl, rot, trans)]
    dc = dev_cam(cam)
    flags = (L.INTERP_VARIANT_STAGED_C16, 0) + ((L.INTERP_FLAG_GENERIC,) if b * h * w <= 20000 else ())
    res = []
    for flag in flags:
        cv, pd, idx = m.utils.get_parallax_sweeping_cv(*args, dc, 4, nbre_cuts=cuts, interp=flag, return_index_grids=True)
        xs = 64
        wide = torch.full((b, h, w, xs), -7.0, device="cuda")
        L.check(L.lib.m4d_pscv_fused_fwd_ex(
            L.ptr(args[0]), L.ptr(args[1]), L.ptr(args[2]), L.ptr(args[3]), L.ptr(args[4]), 4, L.ptr(args[5]), L.ptr(dc["f"]),
            L.ptr(dc["c"]), b, h, w, c, cuts, 4, wide.data_ptr(), xs, None, 0, wide.data_ptr() + 4 * 63, xs, 2.0,
            None, flag, L.stream()))
        res.append((cv.cpu(), pd.cpu(), idx.cpu(), wide.cpu()))
    bits = lambda t: t.view(torch.int32) if t.dtype == torch.float32 else t
    for other in res[1:]:
        for a_, b_ in zip(res[0], other):
            assert torch.equal(bits(a_), bits(b_))
    cv, _, _, wide = res[0]
    assert torch.equal(bits(wide[..., :9]), bits(cv)) and torch.all(wide[..., 9:63] == -7.0)


@pytest.mark.parametrize("interp", ["gather", "bp"])
def test_pscv_degenerate_inputs(interp):
    """What the reference's arithmetic does with degenerate inputs must come out the same: zero translation (s = 0: 0/0 in the
    epipolar direction, App. B), NaN / +-Inf / zero parallax at some pixels, a 90-degree rotation, zero features (0/0 in nothing:
    the PSCV takes normalised features as they are).  NaN patterns and every finite value are compared bit for bit."""
    m = _m4d()
    b, h, w, c, cuts = 3, 12, 16, 32, 2
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(77, b, h, w, c, cuts, "kitti")
    flag = interp == "bp"
    rot[1] = torch.tensor([0.70710678, 0.0, 0.70710678, 0.0])  # 90 degrees about y: alpha = r_z changes sign across the image
    pl[2, 0, 1] = float("inf"); pl[2, 0, 2] = float("-inf"); pl[2, 0, 3] = 0.0
    pt[2, 1, :4] = torch.tensor([float("nan"), float("inf"), 0.0, -1.0]).view(4, 1)
    c2[2, 5:7] = 0.0
    if flag:
        # NaN query points exist only for the BackProject branch (its guard writes zeros).  On the gather branch the reference
        # itself fails there: floor(NaN) becomes an out-of-range tf.gather index (an error on CPU), so those inputs are not part
        # of its contract (libm4d returns zeros for them).
        trans[0] = 0.0                                         # pure rotation: d = 0, s = 0 -> 0/0
        pl[2, 0, 0] = float("nan")
    want_cv, want_pd = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts, use_cuda_backproject=flag)
    for kernel_flag in (0, m._lib.INTERP_FLAG_TILE, m._lib.INTERP_FLAG_GENERIC):
        cv, pd = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4, nbre_cuts=cuts,
                                                  interp=(m.INTERP_BP if flag else m.INTERP_GATHER) | kernel_flag)
        for got, want, name in ((cv.cpu(), want_cv, "cv"), (pd.cpu(), want_pd, "prev_disp")):
            assert torch.equal(torch.isnan(got), torch.isnan(want)), f"{name}: NaN pattern differs (kernel flag {kernel_flag:#x})"
            ok = ~torch.isnan(want)
            assert torch.equal(got[ok], want[ok]), f"{name}: finite values differ (kernel flag {kernel_flag:#x})"


@pytest.mark.parametrize("shape", [(2, 10, 14, 32, 2), (1, 9, 11, 16, 1), (1, 6, 8, 96, 4)])
def test_pscv_backward_vs_autograd_of_the_oracle(shape):
    """Backward of the fused PSCV against torch autograd through the oracle's literal forward (the reference's backward IS
    autodiff of that graph): gradients wrt both feature maps, the previous-frame parallax and this level's parallax.  The
    fp16 stage makes the gradients fp16-quantised; summation orders differ: 2e-3 of each gradient's scale, and the bulk
    (99 % of the entries) within 1e-4."""
    m = _m4d()
    b, h, w, c, cuts = shape
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(300 + c, b, h, w, c, cuts, "kitti")
    pl[:, : h // 3] -= 3.0                                                          # some clipped hypotheses (no gradient through the clip)
    g = torch.Generator().manual_seed(c)
    d_cv = torch.randn(b, h, w, cuts * 9, generator=g)
    d_pd = torch.randn(b, h, w, 9, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (c1, c2, pt, pl)]
    cv, pd = oracle.get_parallax_sweeping_cv(leaves[0], leaves[1], leaves[2], leaves[3], rot, trans, cam, 4, nbre_cuts=cuts,
                                             use_cuda_backproject=False)
    want = torch.autograd.grad([cv, pd], leaves, [d_cv, d_pd])
    got = m.utils.get_parallax_sweeping_cv_grad(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4, cu(d_cv), cu(d_pd),
                                                nbre_cuts=cuts)
    for name, gg, ww in zip(("d_c1", "d_c2", "d_disp_prev_t", "d_disp"), got, want):
        gg = gg.cpu()
        scale = float(ww.abs().max()) + 1e-12
        err = (gg - ww).abs() / scale
        assert float(err.max()) <= 2e-3, (name, float(err.max()))
        assert float(err.flatten().quantile(0.99)) <= 1e-4, (name, float(err.flatten().quantile(0.99)))


@pytest.mark.parametrize("shape", LEVEL_SHAPES, ids=[s[0] for s in LEVEL_SHAPES])
def test_sncv_vs_oracle(shape):
    m = _m4d()
    _, b, h, w, c, cuts, _ = shape
    g = torch.Generator().manual_seed(h * w + c)
    f = oracle.group_l2_normalize(torch.randn(b, h, w, c, generator=g), cuts)
    f2 = oracle.group_l2_normalize(torch.randn(b, h, w, c, generator=g), cuts)
    want = oracle.cost_volume(f, f2, 3, nbre_cuts=cuts)
    out = m.utils.cost_volume(cu(f), cu(f2), 3, nbre_cuts=cuts)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("shape", [(2, 24, 80, 32, 2), (1, 37, 53, 16, 1), (1, 9, 33, 64, 2), (1, 30, 41, 96, 4),
                                   (1, 15, 20, 128, 4), (1, 3, 5, 16, 1)])
def test_sncv_column_strip_kernel_equals_pixel_dy_kernel(shape):
    """The column-strip kernel (group widths 16 / 24 / 32, cuts <= 4; partial tiles, images smaller than a tile) is
    bit-identical to the (pixel, dy) kernel, also when writing into a wider pixel stride."""
    m = _m4d()
    L = m._lib
    b, h, w, c, cuts = shape
    g = torch.Generator().manual_seed(h * w + c)
    f = cu(torch.randn(b, h, w, c, generator=g))
    f2 = cu(torch.randn(b, h, w, c, generator=g))
    oc = 49 * cuts
    outs = []
    for variant in (2, 1, 0):
        wide = torch.full((b, h, w, oc + 5), -7.0, device="cuda")
        L.check(L.lib.m4d_sncv_fwd_ex(L.ptr(f), L.ptr(f2), b, h, w, c, cuts, 3, wide.data_ptr() + 8, oc + 5, variant, L.stream()))
        outs.append(wide.cpu())
    assert torch.equal(outs[0].view(torch.int32), outs[1].view(torch.int32))
    assert torch.equal(outs[0].view(torch.int32), outs[2].view(torch.int32))
    assert torch.all(outs[0][..., :2] == -7.0) and torch.all(outs[0][..., oc + 2:] == -7.0)
    want = oracle.cost_volume(f.cpu(), f2.cpu(), 3, nbre_cuts=cuts)
    np.testing.assert_allclose(outs[0][..., 2:oc + 2].numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_backproject_op_vs_oracle():
    """General BackProject signature (S, F > 1), out-of-range and NaN coordinates, index grids bit-exact."""
    m = _m4d()
    g = torch.Generator().manual_seed(5)
    B, H, W, S, Fd, C = 2, 7, 9, 3, 2, 5
    inp = torch.randn(B, H, W, Fd, C, generator=g)
    coords = torch.rand(B, H, W, S, Fd, 2, generator=g) * torch.tensor([W + 2.0, H + 2.0]) - 1.0
    coords[0, 0, 0, 0, 0, 0] = float("nan")
    coords[0, 1, 1, 0, 0] = torch.tensor([3.0, 2.0])            # integral coordinate: ceil == floor
    coords[1, 2, 2, 1, 1] = torch.tensor([W - 1.0, H - 1.0])    # last pixel
    want = oracle.back_project(inp, coords)
    x0, x1, y0, y1, _ = oracle.back_project_index_grids(coords, H, W)
    out, idx = m.utils.back_project(cu(inp), cu(coords), return_index_grids=True)
    assert torch.equal(idx.cpu(), torch.stack((x0, x1, y0, y1), dim=-1))
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=0, atol=1e-6)
    # vectorised path (C % 4 == 0)
    inp8 = torch.randn(B, H, W, Fd, 8, generator=g)
    np.testing.assert_allclose(m.utils.back_project(cu(inp8), cu(coords)).cpu().numpy(),
                               oracle.back_project(inp8, coords).numpy(), rtol=0, atol=1e-6)


# ------------------------------------------------------- the reference's own compiled BackProject kernels as the oracle
def _ref_binary():
    from oracle import ref_binary
    if not ref_binary.available():
        pytest.fail("oracle/_ref/libbackproject_ref.so is missing: __graft_entry__.build() compiles it from /root/reference "
                    "(cuda_backproject/backproject_op_gpu.cu.cc, unmodified) and it travels with the snapshot")
    return ref_binary


def _bp_case(shape, seed):
    B, H, W, S, Fd, C = shape
    g = torch.Generator().manual_seed(seed)
    inp = torch.randn(B, H, W, Fd, C, generator=g)
    coords = torch.rand(B, H, W, S, Fd, 2, generator=g) * torch.tensor([W + 2.0, H + 2.0]) - 1.0
    coords[0, 0, 0, 0, 0, 0] = float("nan")
    coords[0, 1, 1, 0, 0] = torch.tensor([3.0, 2.0])            # integral coordinate: ceil == floor
    coords[-1, 2, 2, -1, -1] = torch.tensor([W - 1.0, H - 1.0])   # last pixel
    coords[0, 2, 1, 0, 0] = torch.tensor([0.0, 0.0])
    return inp, coords, g


@pytest.mark.parametrize("shape", [(2, 7, 9, 3, 2, 5), (1, 12, 16, 9, 1, 32), (2, 6, 5, 1, 1, 33), (1, 4, 4, 2, 3, 8),
                                   (9, 24, 80, 1, 1, 33), (2, 48, 64, 1, 1, 16)])
def test_backproject_fwd_equals_reference_binary(shape):
    """m4d_backproject_fwd against BackProjectForwardLauncher of the reference itself (backproject_op_gpu.cu.cc:19-103 compiled
    unmodified for sm_100a): every output bit, including the zeros the reference gets from its memset where the coordinate is
    outside the image or NaN.  (9, 24, 80, 1, 1, 33) is the shape class the reference runs it on: 9b tiled copies, c+1 channels.)"""
    m = _m4d()
    ref = _ref_binary()
    inp, coords, _ = _bp_case(shape, 17 + shape[5])
    want = ref.back_project(inp, coords)
    got = m.utils.back_project(cu(inp), cu(coords)).cpu()
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    # and the restated oracle the CPU tests use is within one rounding of the binary (it rounds the products separately)
    np.testing.assert_allclose(oracle.back_project(inp, coords).numpy(), want.numpy(), rtol=0, atol=4e-7 * float(inp.abs().max()))


@pytest.mark.parametrize("shape", [(2, 7, 9, 3, 2, 5), (1, 12, 16, 9, 1, 32), (2, 6, 5, 1, 1, 33), (1, 4, 4, 2, 3, 8)])
def test_backproject_bwd_vs_reference_binary(shape):
    """m4d_backproject_bwd against BackProjectBackwardLauncher of the reference (:108-223).  Both scatter-add inputs_grad with
    floating-point atomics (order not fixed: 1e-5 of the scale); the reference sums coords_grad serially over the channels,
    libm4d in a fixed shuffle order (1e-5 of the scale)."""
    m = _m4d()
    ref = _ref_binary()
    inp, coords, g = _bp_case(shape, 23 + shape[5])
    grad = torch.randn(*coords.shape[:-1], shape[5], generator=g)
    want_i, want_c = ref.back_project_grad(inp, coords, grad)
    got_i, got_c = m.utils.back_project_grad(cu(inp), cu(coords), cu(grad))
    np.testing.assert_allclose(got_i.cpu().numpy(), want_i.numpy(), rtol=1e-5, atol=1e-5 * float(want_i.abs().max()))
    np.testing.assert_allclose(got_c.cpu().numpy(), want_c.numpy(), rtol=1e-4, atol=1e-5 * float(want_c.abs().max()))
    # the restated oracle gradient agrees with the binary as well
    oi, oc = oracle.back_project_grad(inp, coords, grad)
    np.testing.assert_allclose(oi.numpy(), want_i.numpy(), rtol=1e-5, atol=1e-5 * float(want_i.abs().max()))
    np.testing.assert_allclose(oc.numpy(), want_c.numpy(), rtol=1e-4, atol=1e-5 * float(want_c.abs().max()))


@pytest.mark.parametrize("shape", [(2, 24, 80, 32, 2), (1, 37, 53, 16, 1), (1, 12, 40, 128, 4), (2, 6, 20, 192, 8)])
def test_pscv_bp_fma_equals_pipeline_through_reference_binary(shape):
    """The fused PSCV in BP_FMA mode against the reference pipeline with the reference's OWN compiled BackProject kernel doing
    the warp: utils/depth_operations.py:223-281 restated literally (9x tile_in_batch copies, clip, reverse, BackProject with
    S = F = 1 on the [9b, h, w, c+1] tensor, fp16 correlate) where the BackProject call is the binary.  cv and prev_disp bit for
    bit: this is the arithmetic of the reference's GPU path."""
    m = _m4d()
    ref = _ref_binary()
    b, h, w, c, cuts = shape
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(800 + h + c, b, h, w, c, cuts, "kitti")
    pl[:, : h // 3] -= 5.0
    want_cv, want_pd = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts, use_cuda_backproject=True,
                                                       back_project_fn=ref.back_project)
    cv, pd = m.utils.get_parallax_sweeping_cv(cu(c1), cu(c2), cu(pt), cu(pl), cu(rot), cu(trans), dev_cam(cam), 4, nbre_cuts=cuts,
                                              interp=m.INTERP_BP_FMA)
    assert torch.equal(pd.cpu().view(torch.int32), want_pd.view(torch.int32))
    assert torch.equal(cv.cpu().view(torch.int32), want_cv.view(torch.int32))


@pytest.mark.parametrize("shape", [(2, 7, 9, 3, 2, 5), (1, 12, 16, 9, 1, 32), (2, 6, 5, 1, 1, 33), (1, 4, 4, 2, 3, 8)])
def test_backproject_grad_vs_oracle(shape):
    """BackProjectGrad (the gradient TF registers for BackProject): scatter of grad*w onto the taps and the coordinate
    gradient, general (S, F), channel counts with and without the float4 path, coordinates outside the image, NaN and
    integral coordinates.  inputs_grad is a floating-point scatter-add (order not fixed, as in the reference): 1e-5."""
    m = _m4d()
    B, H, W, S, Fd, C = shape
    g = torch.Generator().manual_seed(B * H + C)
    inp = torch.randn(B, H, W, Fd, C, generator=g)
    coords = torch.rand(B, H, W, S, Fd, 2, generator=g) * torch.tensor([W + 2.0, H + 2.0]) - 1.0
    coords[0, 0, 0, 0, 0, 0] = float("nan")
    coords[0, 1, 1, 0, 0] = torch.tensor([3.0, 2.0])            # integral coordinate: ceil == floor, both taps coincide
    coords[-1, 2, 2, -1, -1] = torch.tensor([W - 1.0, H - 1.0])
    grad = torch.randn(B, H, W, S, Fd, C, generator=g)
    want_i, want_c = oracle.back_project_grad(inp, coords, grad)
    d_inp, d_coords, d_grad = cu(inp), cu(coords), cu(grad)
    got_i, got_c = m.utils.back_project_grad(d_inp, d_coords, d_grad)
    scale = float(want_i.abs().max())
    np.testing.assert_allclose(got_i.cpu().numpy(), want_i.numpy(), rtol=1e-5, atol=1e-5 * scale)
    np.testing.assert_allclose(got_c.cpu().numpy(), want_c.numpy(), rtol=1e-4, atol=1e-5 * float(want_c.abs().max()))
    # coords_grad is deterministic and fully written (zeros outside), inputs_grad starts from zero on every call
    poison_i, poison_c = torch.full_like(got_i, 9.0), torch.full_like(got_c, 9.0)
    L = m._lib
    import ctypes
    dim = (ctypes.c_int32 * 6)(B, H, W, S, Fd, C)
    L.check(L.lib.m4d_backproject_bwd(L.ptr(d_grad), L.ptr(d_inp), L.ptr(d_coords), dim, L.ptr(poison_i), L.ptr(poison_c), L.stream()))
    assert torch.equal(poison_c, got_c)
    np.testing.assert_allclose(poison_i.cpu().numpy(), got_i.cpu().numpy(), rtol=1e-5, atol=1e-5 * scale)
    outside = ~((coords[..., 0] >= 0) & (coords[..., 1] >= 0) & (coords[..., 0] <= W - 1) & (coords[..., 1] <= H - 1))
    assert torch.all(got_c.cpu()[outside] == 0)


@pytest.mark.parametrize("cfg", [(2, 9, 13, 3, 16, 1), (1, 16, 16, 16, 16, 2), (2, 15, 20, 24, 40, 2), (1, 8, 12, 122, 128, 1),
                                 (1, 6, 20, 470, 128, 1), (2, 10, 7, 16, 5, 1), (1, 33, 47, 64, 96, 2), (1, 12, 40, 238, 128, 1)])
def test_conv3x3_vs_oracle(cfg):
    m = _m4d()
    b, h, w, cin, cout, stride = cfg
    g = torch.Generator().manual_seed(cin * cout + h)
    x = torch.randn(b, h, w, cin, generator=g)
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, stride))
    from m4depth_b200.m4depth_network import _Conv2D
    conv = _Conv2D(cout, stride)
    conv.assign(k, bias, "cuda")
    out = conv(cu(x), alpha=0.1, algo=1).clone()
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    # input living inside a wider (strided) buffer, as the refiner input does
    xs = (cin + 3) // 4 * 4 + 4
    wide = torch.full((b, h, w, xs), 7.0)
    wide[..., :cin] = x
    out2 = conv(cu(wide)[..., :cin], alpha=0.1, algo=1)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("cfg", [(1, 16, 8, 32, 16), (2, 16, 16, 64, 128), (1, 24, 80, 122, 128), (2, 13, 21, 128, 96),
                                 (1, 6, 20, 470, 128), (1, 33, 47, 96, 64), (1, 12, 40, 238, 32), (1, 20, 9, 16, 16),
                                 (1, 48, 64, 128, 128), (2, 19, 23, 16, 5), (1, 16, 8, 40, 20)])
def test_conv3x3_tcgen05_3xtf32_vs_oracle(cfg):
    """The tensor-core path (tcgen05, 3xTF32: hi*hi + hi*lo + lo*hi with fp32 accumulation) against the fp32 oracle conv
    and against the FFMA2 kernel; partial tiles, channel counts that are not multiples of 32 (TMA zero fill), inputs
    living inside a wider pixel stride.  Tolerance: 1e-5 of the tensor scale (north_star allows 1e-4)."""
    m = _m4d()
    b, h, w, cin, cout = cfg
    g = torch.Generator().manual_seed(cin * cout + h)
    x = torch.randn(b, h, w, cin, generator=g)
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, 1))
    from m4depth_b200.m4depth_network import _Conv2D
    conv = _Conv2D(cout, 1)
    conv.assign(k, bias, "cuda")
    assert conv.packed is not None
    xs = (cin + 3) // 4 * 4
    wide = torch.full((b, h, w, xs + 4), 7.0)
    wide[..., :cin] = x
    xin = cu(wide)[..., :cin]
    ffma = conv(xin, alpha=0.1, algo=1).clone()
    tc = conv(xin, alpha=0.1, algo=2).clone()
    scale = float(want.abs().max())
    np.testing.assert_allclose(tc.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * scale)
    np.testing.assert_allclose(tc.cpu().numpy(), ffma.cpu().numpy(), rtol=1e-5, atol=1e-5 * scale)
    # against an fp64 evaluation: the 3xTF32 error must stay in the fp32 class (single-pass TF32 would be ~5e-4 of the
    # scale).  The tensor core adds each K=8 partial product to the fp32 accumulator with truncation, so its error
    # grows with the number of accumulations (9 taps x cin/8) and is a few times the FFMA kernel's.
    ref64 = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), k.double().permute(3, 2, 0, 1), bias.double(), padding=1)
    ref64 = torch.nn.functional.leaky_relu(ref64, 0.1).permute(0, 2, 3, 1)
    e_tc = float((tc.cpu().double() - ref64).abs().max()) / scale
    e_ffma = float((ffma.cpu().double() - ref64).abs().max()) / scale
    print(f"cin={cin} cout={cout}: max error / scale  tcgen05 3xTF32 {e_tc:.2e}   FFMA2 {e_ffma:.2e}")
    assert e_tc < 4e-6


@pytest.mark.parametrize("cfg", [(2, 32, 48, 16, 16), (1, 24, 80, 32, 32), (1, 20, 36, 64, 64), (2, 12, 40, 96, 96),
                                 (1, 12, 40, 128, 128), (1, 12, 40, 192, 192), (1, 2, 2, 16, 16), (1, 34, 18, 48, 80)])
def test_conv3x3_stride2_tcgen05_vs_oracle(cfg):
    """FeaturePyramid's stride-2 convs on the tensor cores (2x2-cell formulation over a 5-D tensor map, k-steps without
    weights skipped) against the fp32 oracle conv (TF SAME: even sizes pad bottom / right only), the FFMA2 kernel and an
    fp64 evaluation; includes the 192-channel level (two launches over output-channel halves) and partial tiles."""
    m = _m4d()
    b, h, w, cin, cout = cfg
    g = torch.Generator().manual_seed(cin * cout + h)
    x = torch.randn(b, h, w, cin, generator=g)
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, 2))
    from m4depth_b200.m4depth_network import _Conv2D
    conv = _Conv2D(cout, 2)
    conv.assign(k, bias, "cuda")
    assert conv.packed is not None
    ffma = conv(cu(x), alpha=0.1, algo=1).clone()
    tc = conv(cu(x), alpha=0.1, algo=2).clone()
    assert tuple(tc.shape) == tuple(want.shape)
    scale = float(want.abs().max())
    np.testing.assert_allclose(tc.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * scale)
    np.testing.assert_allclose(tc.cpu().numpy(), ffma.cpu().numpy(), rtol=1e-5, atol=1e-5 * scale)
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref64 = torch.nn.functional.conv2d(xp, k.double().permute(3, 2, 0, 1), bias.double(), stride=2)
    ref64 = torch.nn.functional.leaky_relu(ref64, 0.1).permute(0, 2, 3, 1)
    assert float((tc.cpu().double() - ref64).abs().max()) / scale < 4e-6
    # odd sizes (TF pads both sides): shifted into a zeroed even-sized buffer, then the same tensor-core kernel (auto and forced)
    xo = cu(torch.randn(1, 15, 20, cin, generator=g))
    want_o = oracle.leaky_relu(oracle.conv2d_same(xo.cpu(), k, bias, 2))
    np.testing.assert_allclose(conv(xo, alpha=0.1).cpu().numpy(), want_o.numpy(), rtol=1e-5, atol=1e-5 * float(want_o.abs().max()))
    np.testing.assert_allclose(conv(xo, alpha=0.1, algo=2).cpu().numpy(), want_o.numpy(), rtol=1e-5, atol=1e-5 * float(want_o.abs().max()))
    np.testing.assert_allclose(conv(xo, alpha=0.1, algo=1).cpu().numpy(), want_o.numpy(), rtol=1e-5, atol=1e-5 * float(want_o.abs().max()))


@pytest.mark.parametrize("cfg", [(1, 6, 20, 470, 128, 1), (2, 12, 40, 238, 128, 1), (1, 12, 40, 128, 96, 1), (1, 9, 17, 96, 64, 1),
                                 (1, 12, 40, 128, 128, 2), (1, 16, 8, 64, 192, 1)])
def test_conv3x3_tcgen05_output_channel_slices(cfg):
    """A pixel tile's output channels split over 1 / 2 / 4 CTAs (what the library does for layers with fewer tiles than
    SMs): every legal split matches the oracle and the fp64 evaluation to the same tolerance as the unsplit kernel."""
    m = _m4d()
    b, h, w, cin, cout, stride = cfg
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(b, h, w, cin, generator=g)
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, stride))
    scale = float(want.abs().max())
    from m4depth_b200.m4depth_network import _Conv2D
    conv = _Conv2D(cout, stride)
    conv.assign(k, bias, "cuda")
    xs = (cin + 3) // 4 * 4                       # 470 / 238 channels live in a 472 / 240-float pixel stride (refiner input)
    wide = torch.zeros(b, h, w, xs)
    wide[..., :cin] = x
    xin = cu(wide)[..., :cin]
    ran = 0
    for slices in (0, 1, 2, 4):
        cp = (cout + 15) // 16 * 16
        legal = slices == 0 or (cp % slices == 0 and cp // slices <= 128 and (slices == 1 or (cp // slices) % 32 == 0))
        if not legal:
            with pytest.raises(m.M4DError):
                conv(xin, alpha=0.1, algo=2, slices=slices)
            continue
        out = conv(xin, alpha=0.1, algo=2, slices=slices).clone()
        np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * scale)
        ran += 1
    assert ran >= 2


@pytest.mark.parametrize("cfg", [(1, 16, 8, 32, 16, 1), (2, 16, 16, 64, 128, 1), (1, 24, 80, 122, 128, 1), (2, 13, 21, 128, 96, 1),
                                 (1, 6, 20, 470, 128, 1), (1, 33, 47, 96, 64, 1), (1, 12, 40, 238, 32, 1), (1, 20, 9, 16, 16, 1),
                                 (2, 19, 23, 16, 5, 1), (1, 16, 8, 40, 20, 1), (2, 32, 48, 16, 16, 2), (1, 20, 36, 64, 64, 2),
                                 (1, 12, 40, 192, 192, 2), (1, 34, 18, 48, 80, 2), (1, 12, 40, 128, 192, 1),
                                 (1, 1, 1, 32, 16, 1), (1, 2, 3, 16, 128, 1), (2, 2, 2, 16, 16, 2), (1, 1, 40, 64, 32, 1)])
@pytest.mark.parametrize("dyn", ["unit", "wide"])
def test_conv3x3_tcgen05_3xfp16_vs_oracle(cfg, dyn):
    """The 3xFP16 mode of the tensor-core conv (operands scaled per layer / per tile-k-block and split into fp16 h1 + 2^-11 h2;
    hi*hi + hi*lo + lo*hi in fp32) is in the same error class as 3xTF32 and the FFMA kernel: 1e-5 of the tensor scale against
    the oracle, 4e-6 against fp64 - also when the input's magnitude varies over 12 orders between image regions and
    channels ("wide"), which fp16's exponent range alone could not represent."""
    m = _m4d()
    b, h, w, cin, cout, stride = cfg
    g = torch.Generator().manual_seed(cin * cout + h)
    x = torch.randn(b, h, w, cin, generator=g)
    if dyn == "wide":
        x = x * torch.exp(torch.randn(b, h, 1, 1, generator=g) * 6.0) * torch.exp(torch.randn(1, 1, 1, cin, generator=g) * 3.0)
        x[:, : h // 4] = 0.0                                                     # an all-zero region: scale falls back to 1
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    k[..., : cout // 3] *= 1e-3                                                  # output channels with tiny weights
    bias = torch.randn(cout, generator=g) * 0.1
    from m4depth_b200.m4depth_network import _Conv2D
    conv = _Conv2D(cout, stride, prec=1)
    conv.assign(k, bias, "cuda")
    xs = (cin + 3) // 4 * 4
    if stride == 1:
        wide = torch.zeros(b, h, w, xs + 4)
        wide[..., :cin] = x
        xin = cu(wide)[..., :cin]
    else:
        xin = cu(x)
    tc = conv(xin, alpha=0.1, algo=2).clone()
    pad = (0, 1, 0, 1) if stride == 2 else (1, 1, 1, 1)
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), pad)
    ref64 = torch.nn.functional.conv2d(xp, k.double().permute(3, 2, 0, 1), bias.double(), stride=stride)
    ref64 = torch.nn.functional.leaky_relu(ref64, 0.1).permute(0, 2, 3, 1)
    # error measured against the local magnitude of the sum (|x| * |w| convolved), so that the small-magnitude regions of the
    # "wide" input are held to the same relative standard as the large ones
    mag = torch.nn.functional.conv2d(xp.abs(), k.double().abs().permute(3, 2, 0, 1), bias.double().abs(), stride=stride).permute(0, 2, 3, 1)
    err = ((tc.cpu().double() - ref64).abs() / (mag + 1e-30)).max()
    print(f"cin={cin} cout={cout} s{stride} {dyn}: max error / local magnitude {float(err):.2e}")
    assert float(err) < (2e-6 if dyn == "unit" else 1e-5)
    if dyn == "unit":
        want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, stride))
        np.testing.assert_allclose(tc.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * float(want.abs().max()))


@pytest.mark.parametrize("cfg", [(4, 96, 160, 16, 16, 1), (4, 192, 320, 16, 16, 2), (2, 128, 160, 64, 32, 1), (2, 96, 160, 96, 64, 1),
                                 (2, 96, 160, 64, 128, 1), (3, 96, 128, 32, 5, 1)])
def test_conv3x3_tcgen05_many_tiles_per_cta(cfg):
    """Several hundred pixel tiles, i.e. every persistent CTA walks many tiles: the mbarrier rings (halo, weights - streamed or
    resident -, the two accumulator sets, the alternating epilogue groups of thin layers) wrap around many times."""
    m = _m4d()
    b, h, w, cin, cout, stride = cfg
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(b, h, w, cin, generator=g)
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    pad = (0, 1, 0, 1) if stride == 2 else (1, 1, 1, 1)
    ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x.permute(0, 3, 1, 2), pad), k.permute(3, 2, 0, 1), bias, stride=stride)
    ref = torch.nn.functional.leaky_relu(ref, 0.1).permute(0, 2, 3, 1)
    from m4depth_b200.m4depth_network import _Conv2D
    for prec in (1, 0):
        conv = _Conv2D(cout, stride, prec=prec)
        conv.assign(k, bias, "cuda")
        for _ in range(2):                                   # twice: a second launch right behind the first
            out = conv(cu(x), alpha=0.1, algo=2)
        np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-5, atol=2e-5 * float(ref.abs().max()))


def test_conv3x3_tcgen05_random_shapes():
    """Seeded fuzz over the tensor-core conv's shape space: any cin (TMA zero fill of partial k-blocks), cout 1..192 (padding,
    slices), stride 2 where the cell formulation applies, 1..3 images of 1..70 pixels a side (partial and empty tile rows),
    both precision modes, inputs inside a wider pixel stride."""
    m = _m4d()
    from m4depth_b200.m4depth_network import _Conv2D
    rng = np.random.default_rng(2024)
    g = torch.Generator().manual_seed(2024)
    ran = 0
    for case in range(48):
        stride = int(rng.integers(1, 3))
        cin = int(rng.choice([16, 32, 48, 64, 96, 128, 192])) if stride == 2 else int(rng.integers(16, 260))
        cout = int(rng.integers(1, 193))
        if cout > 128 and (-(-cout // 16) * 16) % 64 != 0:
            cout = 192                                          # beyond 128 channels the slices must be multiples of 32
        b = int(rng.integers(1, 4))
        h, w = int(rng.integers(1, 71)), int(rng.integers(1, 71))
        if stride == 2:
            h, w = 2 * ((h + 1) // 2), 2 * ((w + 1) // 2)
        prec = int(rng.integers(0, 2))
        x = torch.randn(b, h, w, cin, generator=g)
        k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
        bias = torch.randn(cout, generator=g) * 0.1
        pad = (0, 1, 0, 1) if stride == 2 else (1, 1, 1, 1)
        ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), pad), k.double().permute(3, 2, 0, 1),
                                         bias.double(), stride=stride)
        ref = torch.nn.functional.leaky_relu(ref, 0.1).permute(0, 2, 3, 1)
        conv = _Conv2D(cout, stride, prec=prec)
        conv.assign(k, bias, "cuda")
        if stride == 1:
            xs = (cin + 3) // 4 * 4 + 4 * int(rng.integers(0, 2))
            wide = torch.full((b, h, w, xs), 3.0)
            wide[..., :cin] = x
            xin = cu(wide)[..., :cin]
        else:
            xin = cu(x)
        out = conv(xin, alpha=0.1, algo=2)
        err = float((out.cpu().double() - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 5e-6, (case, b, h, w, cin, cout, stride, prec, err)
        ran += 1
    assert ran == 48


def test_conv3x3_tcgen05_wide_output_split():
    """cout = 192 (> 128 TMEM-friendly columns): output channels sliced over two CTAs per tile, stride 1 (128->192)."""
    m = _m4d()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 12, 40, 128, generator=g)
    k = torch.randn(3, 3, 128, 192, generator=g) * (2.0 / (9 * 128)) ** 0.5
    bias = torch.randn(192, generator=g) * 0.1
    want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, 1))
    from m4depth_b200.m4depth_network import _Conv2D
    conv = _Conv2D(192, 1)
    conv.assign(k, bias, "cuda")
    tc = conv(cu(x), alpha=0.1, algo=2)
    np.testing.assert_allclose(tc.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * float(want.abs().max()))


def test_model_loads_reference_format_checkpoint(tmp_path):
    """Weights written in the reference's TF tensor-bundle format (object-graph key layout) load through
    M4Depth.load_checkpoint and give the same depth, bit for bit, as the same weights passed as a dict."""
    m = _m4d()
    from m4depth_b200.checkpoint import save_reference_weights
    nl, H, W = 3, 64, 96
    wts = oracle.init_weights(nl, seed=4, bias_std=0.05, dn_random=True)
    prefix = str(tmp_path / "cp-0001.ckpt")
    save_reference_weights(prefix, wts)
    g = torch.Generator().manual_seed(2)
    cam = {"f": torch.tensor([[0.5 * W, 0.5 * H]]).cuda(), "c": torch.tensor([[0.5 * W, 0.5 * H]]).cuda()}
    rot, trans = motion(g, 1)
    frames = [torch.rand(1, H, W, 3, generator=g).cuda() for _ in range(2)]
    outs = []
    for loader in ("dict", "bundle"):
        mod = m.M4Depth(nbre_levels=nl, use_cuda_graph=False)
        if loader == "dict":
            mod.load_weights(wts)
        else:
            mod.load_checkpoint(prefix)
        for t, rgb in enumerate(frames):
            out = mod([[{"RGB_im": rgb, "rot": rot.cuda(), "trans": trans.cuda(), "new_traj": [t == 0]}], cam])["depth"]
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0]).all()


@pytest.mark.parametrize("shape", [(2, 32, 48), (1, 37, 53), (3, 8, 8)])
def test_first_encoder_layer_fused_conv_dn(shape):
    """conv3x3(RGB -> 16) + DomainNormalization + leaky_relu in one call that never stores the conv output (both DN passes
    recompute it) against the oracle and against the two separate ops."""
    m = _m4d()
    b, h, w = shape
    g = torch.Generator().manual_seed(b * h + w)
    wts = oracle.init_weights(1, seed=9, bias_std=0.05, dn_random=True)
    rgb = torch.rand(b, h, w, 3, generator=g)
    conv = oracle.conv2d_same(rgb, wts["encoder/conv_layers_s1/0/kernel"], wts["encoder/conv_layers_s1/0/bias"], 1)
    dn = oracle.DomainNormalization(wts["encoder/dn_layers/0/scale"], wts["encoder/dn_layers/0/bias"])
    want = oracle.leaky_relu(dn(conv))
    settings = {"nbre_lvls": 1, "is_training": False, "ablation": m.M4depthAblationParameters()}
    outs = []
    for mode in (2, 0, 1, 3):           # weights as kernel parameters | two ops | fused, weights in shared memory | conv once + stats
        unfused = mode == 0
        enc = m.FeaturePyramid(settings)
        enc.unfused_first_layer, enc.first_layer_mode = unfused, mode
        enc.conv_layers_s1[0].assign(wts["encoder/conv_layers_s1/0/kernel"], wts["encoder/conv_layers_s1/0/bias"], "cuda")
        enc.conv_layers_s2[0].assign(wts["encoder/conv_layers_s2/0/kernel"], wts["encoder/conv_layers_s2/0/bias"], "cuda")
        enc.dn_layers[0].scale = wts["encoder/dn_layers/0/scale"].cuda().reshape(1, 1, 1, -1).contiguous()
        enc.dn_layers[0].bias = wts["encoder/dn_layers/0/bias"].cuda().reshape(1, 1, 1, -1).contiguous()
        x = cu(rgb)
        tmp = enc._first_layer_fused(enc.conv_layers_s1[0], x) if not unfused else enc.dn_layers[0].call(
            enc.conv_layers_s1[0](x, alpha=1.0), leaky_alpha=0.1)
        outs.append(tmp.clone())
    np.testing.assert_allclose(outs[0].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[1].cpu().numpy(), rtol=1e-5, atol=1e-6)
    # same FMA chain; the double-precision statistics are summed in another order (a last-bit difference of a mean at most)
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[2].cpu().numpy(), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[3].cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_test_step_protocol_single_frames_and_kitti_sequence_mode():
    """M4Depth.test_step (m4depth_network.py:433-474): per-frame mode skips the frame that starts a trajectory and clips
    gt / estimate before scoring; sequence mode (5-D inputs) scores only the last frame; both agree with the oracle's
    metrics on the same depth maps."""
    m = _m4d()
    nl, H, W, T = 3, 64, 96, 3
    wts = oracle.init_weights(nl, seed=6, bias_std=0.05, dn_random=True)
    g = torch.Generator().manual_seed(12)
    cam = {"f": torch.tensor([[0.5 * W, 0.5 * H]]).cuda(), "c": torch.tensor([[0.5 * W, 0.5 * H]]).cuda()}
    rot, trans = motion(g, 1)
    rgbs = [torch.rand(1, H, W, 3, generator=g) for _ in range(T)]
    gts = [torch.rand(1, H, W, 1, generator=g) * 100.0 for _ in range(T)]          # some values above the 80 m clip
    mod = m.M4Depth(nbre_levels=nl, use_cuda_graph=False)
    mod.load_weights(wts)
    want = []
    for t in range(T):
        data = {"RGB_im": rgbs[t].cuda(), "rot": rot.cuda(), "trans": trans.cuda(), "new_traj": [t == 0], "depth": gts[t].cuda(),
                "camera": cam}
        res = mod.test_step(data)
        if t > 0:
            want.append(oracle.depth_metrics(gts[t], mod._out.cpu()))
    assert mod.compiled_metrics.count == T - 1                                     # frame 0 (new_traj) is not scored
    pred = mod.predict_step({"RGB_im": rgbs[0].cuda(), "rot": rot.cuda(), "trans": trans.cuda(), "new_traj": [True], "camera": cam})
    assert sorted(pred) == ["depth", "image", "new_traj"] and tuple(pred["depth"].shape) == (1, H, W, 1)      # :481-488
    mod.predict_step({"RGB_im": rgbs[1].cuda(), "rot": rot.cuda(), "trans": trans.cuda(), "new_traj": [False], "camera": cam})
    mod.predict_step({"RGB_im": rgbs[2].cuda(), "rot": rot.cuda(), "trans": trans.cuda(), "new_traj": [False], "camera": cam})
    mean = torch.stack(want).mean(dim=0)                                          # keras Mean of the per-batch values
    for k, v in zip(oracle.METRIC_NAMES, mean.tolist()):
        assert abs(res[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, res[k], v)
    last_single = mod._out.clone()
    # sequence mode on a fresh model: same frames stacked on axis 1, only the last frame scored
    mod2 = m.M4Depth(nbre_levels=nl, use_cuda_graph=False)
    mod2.load_weights(wts)
    seq = {"RGB_im": torch.stack(rgbs, 1).cuda(), "rot": torch.stack([rot] * T, 1).cuda(), "trans": torch.stack([trans] * T, 1).cuda(),
           "new_traj": torch.tensor([[True] + [False] * (T - 1)]), "depth": torch.stack(gts, 1).cuda(), "camera": cam}
    res2 = mod2.test_step(seq)
    assert mod2.compiled_metrics.count == 1
    assert torch.equal(mod2._out, last_single)
    one = oracle.depth_metrics(gts[-1], last_single.cpu())
    for k, v in zip(oracle.METRIC_NAMES, one.tolist()):
        assert abs(res2[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, res2[k], v)


def test_resize_and_prologue_epilogue_vs_oracle():
    m = _m4d()
    L = m._lib
    g = torch.Generator().manual_seed(11)
    b, ih, iw, h, w = 2, 8, 10, 15, 20
    cam = camera_for("tartan", b, h, w)
    rot, trans = motion(g, b)
    other = torch.randn(b, ih, iw, 4, generator=g)
    para = torch.rand(b, ih, iw, 1, generator=g) * 4 + 0.2
    depth = torch.rand(b, ih, iw, 1, generator=g) * 40 + 1
    state = torch.rand(b, h, w, 1, generator=g) * 40 + 3
    want_other = oracle.resize_bilinear_legacy(other, h, w)
    want_para = oracle.resize_bilinear_legacy(para, h, w) * 2.0
    want_depth = oracle.resize_bilinear_legacy(depth, h, w)
    want_pt = oracle.prev_d2para(state, rot, trans, cam)
    e = lambda *s: torch.empty(*s, device="cuda")
    o_pl, o_dl, o_ot, o_pt, x_in = e(b, h, w, 1), e(b, h, w, 1), e(b, h, w, 4), e(b, h, w, 1), torch.zeros(b, h, w, 12, device="cuda")
    d = [cu(t) for t in (other, para, depth, state, rot, trans, cam["f"], cam["c"])]
    L.check(L.lib.m4d_level_prologue(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), ih, iw, d[3].data_ptr(), d[4].data_ptr(), 4,
                                     d[5].data_ptr(), d[6].data_ptr(), d[7].data_ptr(), b, h, w, o_pl.data_ptr(), o_dl.data_ptr(),
                                     o_ot.data_ptr(), o_pt.data_ptr(), x_in.data_ptr(), 12, 2, 5, 0.25, L.stream()))
    assert torch.equal(o_ot.cpu(), want_other) and torch.equal(o_pl.cpu(), want_para) and torch.equal(o_dl.cpu(), want_depth)
    assert torch.equal(o_pt.cpu(), want_pt)
    assert torch.equal(x_in[..., 5:9].cpu(), want_other)
    assert torch.allclose(x_in[..., 2].cpu(), torch.log(want_para[..., 0] * 0.25), rtol=1e-6, atol=1e-6)
    # legacy resize standalone + nearest
    r = e(b, h, w, 4)
    L.check(L.lib.m4d_resize_bilinear_legacy(d[0].data_ptr(), b, ih, iw, 4, h, w, 1.0, r.data_ptr(), 4, L.stream()))
    assert torch.equal(r.cpu(), want_other)
    n = e(b, 2 * h, 2 * w, 1)
    L.check(L.lib.m4d_resize_nearest(d[3].data_ptr(), b, h, w, 1, 2 * h, 2 * w, n.data_ptr(), L.stream()))
    assert torch.equal(n.cpu(), oracle.resize_nearest(state, 2 * h, 2 * w))
    for oh2, ow2 in ((3 * h + 1, 4 * ((5 * w) // 8)), (h + 3, 2 * w + 1)):      # non-integer factors: 16-byte-store kernel, generic kernel
        n2 = e(b, oh2, ow2, 1)
        L.check(L.lib.m4d_resize_nearest(d[3].data_ptr(), b, h, w, 1, oh2, ow2, n2.data_ptr(), L.stream()))
        assert torch.equal(n2.cpu(), oracle.resize_nearest(state, oh2, ow2))
    # epilogue
    rr = torch.randn(b, h, w, 5, generator=g) * 3
    rr[0, 0, 0, 0] = 9.0       # clipped at 7
    o_p, o_d, o_o, o_s = e(b, h, w, 1), e(b, h, w, 1), e(b, h, w, 4), e(b, h, w, 1)
    L.check(L.lib.m4d_level_epilogue(cu(rr).data_ptr(), 5, d[4].data_ptr(), 4, d[5].data_ptr(), d[6].data_ptr(), d[7].data_ptr(),
                                     b, h, w, 2.0, o_p.data_ptr(), o_d.data_ptr(), o_o.data_ptr(), o_s.data_ptr(), L.stream()))
    want_p = torch.exp(torch.clamp(rr[..., :1], -7., 7.)) / 0.5
    assert torch.allclose(o_p.cpu(), want_p, rtol=2e-6, atol=0)
    want_d = oracle.parallax2depth(o_p.cpu(), rot, trans, cam)
    assert torch.equal(o_d.cpu(), want_d) and torch.equal(o_s, o_d) and torch.equal(o_o.cpu(), rr[..., 1:])


def test_group_l2norm_and_dn_vs_oracle():
    m = _m4d()
    L = m._lib
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 9, 11, 96, generator=g)
    out = torch.empty_like(x, device="cuda")
    L.check(L.lib.m4d_group_l2norm(cu(x).data_ptr(), 2 * 9 * 11, 96, 4, out.data_ptr(), L.stream()))
    np.testing.assert_allclose(out.cpu().numpy(), oracle.group_l2_normalize(x, 4).numpy(), rtol=5e-7, atol=0)
    # every (channels, groups) pair of the network (m4depth_network.py:172-189): 16- and 32-channel groups run the
    # warp-cooperative kernel, 24-channel groups the one-thread-per-group kernel; in place as well; a zero group gives NaN (no epsilon)
    for c, cuts in ((16, 1), (32, 2), (64, 2), (96, 4), (128, 4), (192, 8)):
        x = torch.randn(3, 7, 13, c, generator=g) * 3.0
        x[1, 2, 3, : c // cuts] = 0.0
        want = oracle.group_l2_normalize(x, cuts)
        buf = cu(x).clone()
        L.check(L.lib.m4d_group_l2norm(buf.data_ptr(), 3 * 7 * 13, c, cuts, buf.data_ptr(), L.stream()))
        got = buf.cpu()
        assert torch.equal(torch.isnan(got), torch.isnan(want))
        ok = ~torch.isnan(want)
        np.testing.assert_allclose(got[ok].numpy(), want[ok].numpy(), rtol=5e-7, atol=0)
    # DN at an encoder-like shape, random affine, fused leaky
    xx = torch.randn(2, 48, 64, 16, generator=g) * (torch.rand(1, 1, 1, 16, generator=g) * 2 + 0.1) + torch.randn(1, 1, 1, 16, generator=g)
    sc, bi = torch.rand(1, 1, 1, 16, generator=g) + 0.5, torch.randn(1, 1, 1, 16, generator=g) * 0.1
    dn = m.DomainNormalization()
    dn.scale, dn.bias = cu(sc), cu(bi)
    want = oracle.leaky_relu(oracle.DomainNormalization(sc, bi)(xx))
    np.testing.assert_allclose(dn.call(cu(xx), leaky_alpha=0.1).cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6)


def test_metrics_vs_oracle():
    m = _m4d()
    g = torch.Generator().manual_seed(17)
    gt = torch.rand(2, 32, 48, 1, generator=g) * 100 - 5
    est = gt * torch.exp(torch.randn(2, 32, 48, 1, generator=g) * 0.3)
    want = oracle.depth_metrics(gt, est)
    got = m.metrics.depth_metrics(cu(gt), cu(est)).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-5, atol=1e-7)


# ------------------------------------------------------------------------- headline size (BASELINE config 3, L2)
def test_pscv_headline_size_vs_oracle_and_batch_independence():
    """96x320x32, cuts 2, r=4, b=8 (the roofline configuration): bit-exact vs the oracle on the full tensor, and the
    batched launch equals eight single-image launches (sequences never interact)."""
    m = _m4d()
    b, h, w, c, cuts = 8, 96, 320, 32, 2
    c1, c2, pt, pl, rot, trans, cam = pscv_inputs(1234, b, h, w, c, cuts, "kitti")
    dc = dev_cam(cam)
    d = [cu(t) for t in (c1, c2, pt, pl, rot, trans)]
    cv, pd = m.utils.get_parallax_sweeping_cv(d[0], d[1], d[2], d[3], d[4], d[5], dc, 4, nbre_cuts=cuts)
    want_cv, want_pd = oracle.get_parallax_sweeping_cv(c1, c2, pt, pl, rot, trans, cam, 4, nbre_cuts=cuts, use_cuda_backproject=False)
    assert torch.equal(cv.cpu(), want_cv) and torch.equal(pd.cpu(), want_pd)
    for i in (0, 5):
        cvi, pdi = m.utils.get_parallax_sweeping_cv(d[0][i:i + 1], d[1][i:i + 1], d[2][i:i + 1], d[3][i:i + 1], d[4][i:i + 1],
                                                    d[5][i:i + 1], {"f": dc["f"][i:i + 1], "c": dc["c"][i:i + 1]}, 4, nbre_cuts=cuts)
        assert torch.equal(cvi[0], cv[i]) and torch.equal(pdi[0], pd[i])


@pytest.mark.parametrize("off", ["DINL", "SNCV", "time_recurr", "normalize_features", "subdivide_features", "level_memory"])
def test_model_ablations_vs_oracle(off):
    """M4depthAblationParameters (m4depth_network.py:21-22): each switch turned off in turn changes the refiner input layout
    / the feature preparation / the encoder; the whole model must still follow the oracle model built with the same settings."""
    m = _m4d()
    nl, b, H, W = 3, 1, 64, 96
    flags = dict(DINL=True, SNCV=True, time_recurr=True, normalize_features=True, subdivide_features=True, level_memory=True)
    flags[off] = False
    ab_ref = oracle.M4depthAblationParameters(**flags)
    ab_gpu = m.M4depthAblationParameters(**flags)
    # the refiner input width depends on the switches (m4depth_network.py:223-242): first refiner kernel of every level to match
    wts = oracle.init_weights(nl, seed=11, bias_std=0.05, dn_random=True)
    cuts = lambda lvl: (2 ** (lvl // 2)) if flags["subdivide_features"] else 1
    for lvl in range(1, nl + 1):
        cin = 9 * cuts(lvl) + 1 + (4 if flags["level_memory"] else 0) + (49 * cuts(lvl) if flags["SNCV"] else 0) + (1 if flags["time_recurr"] else 0)
        gk = torch.Generator().manual_seed(100 + lvl)
        wts[f"d_estimator/levels/{lvl - 1}/disp_refiner/prep_conv_layers/0/kernel"] = torch.randn(3, 3, cin, 128, generator=gk) * (2.0 / (9 * cin)) ** 0.5
    ref = oracle.M4Depth(wts, nbre_levels=nl, ablation_settings=ab_ref, pscv_kwargs={"use_cuda_backproject": False})
    mod = m.M4Depth(nbre_levels=nl, ablation_settings=ab_gpu, use_cuda_graph=False)
    mod.load_weights(wts)
    g = torch.Generator().manual_seed(31)
    cam = camera_for("kitti", b, H, W)
    for t in range(3):
        rot, trans = motion(g, b)
        rgb = torch.rand(b, H, W, 3, generator=g)
        s = {"RGB_im": rgb, "rot": rot, "trans": trans, "new_traj": [t == 0] * b}
        want = ref([[s], cam])["depth"].numpy()
        got = mod([[{k: (cu(v) if hasattr(v, "shape") else v) for k, v in s.items()}], dev_cam(cam)])["depth"].cpu().numpy()
        # A mis-wired layout (wrong channel offsets, a missing input) shows as O(0.1-1) errors everywhere.  The bound is looser
        # than check_depth's: without DN / feature normalisation the fp16 stage of the PSCV sees unnormalised magnitudes and
        # amplifies the legitimate 1e-7 summation-order differences more (measured: median <= 9e-5, 99th percentile <= 5e-3).
        err = np.abs(got - want) / (np.abs(want) + 0.1)
        if t == 0:
            assert err.max() <= 1e-4
        else:
            assert np.median(err) <= 5e-4 and np.percentile(err, 99) <= 2e-2 and err.max() <= 0.2, \
                (off, t, np.median(err), np.percentile(err, 99), err.max())


def test_model_shapes_are_static_like_the_reference_state_variables():
    """The levels' recurrent state is allocated for one (batch, height, width), as the reference's tf.Variables are
    (m4depth_network.py:160-163): a later call with another shape is refused with an error, not silently re-built; a second
    model object handles the other shape."""
    m = _m4d()
    nl = 4
    wts = oracle.init_weights(nl, seed=8, bias_std=0.05, dn_random=True)

    def frame(b, H, W, t, g):
        rot, trans = motion(g, b)
        return {"RGB_im": cu(torch.rand(b, H, W, 3, generator=g)), "rot": cu(rot), "trans": cu(trans), "new_traj": [t == 0] * b}

    g = torch.Generator().manual_seed(1)
    model = m.M4Depth(nbre_levels=nl, use_cuda_graph=True)
    model.load_weights(wts)
    for t in range(3):
        model([[frame(2, 64, 96, t, g)], dev_cam(camera_for("kitti", 2, 64, 96))])
    with pytest.raises(m.M4DError, match="static shapes"):
        model([[frame(1, 96, 64, 0, g)], dev_cam(camera_for("kitti", 1, 96, 64))])
    other = m.M4Depth(nbre_levels=nl, use_cuda_graph=True)
    other.load_weights(wts)
    out = other([[frame(1, 96, 64, 0, g)], dev_cam(camera_for("kitti", 1, 96, 64))])["depth"]
    assert tuple(out.shape) == (1, 96, 64, 1) and float(out.min()) == 1000.0


def test_model_streaming_graph_equals_eager():
    """6 levels, 5 frames with a trajectory reset in the middle: CUDA-graph replay is bit-identical to eager execution."""
    m = _m4d()
    g = torch.Generator().manual_seed(23)
    nl, b, H, W = 6, 2, 128, 192
    w = oracle.init_weights(nl, seed=3, bias_std=0.05, dn_random=True)
    cam = dev_cam(camera_for("kitti", b, H, W))
    outs = {}
    for graph in (False, True):
        model = m.M4Depth(nbre_levels=nl, use_cuda_graph=graph)
        model.load_weights(w)
        gg = torch.Generator().manual_seed(29)
        res = []
        for t in range(9):
            rot, trans = motion(gg, b)
            rgb = torch.rand(b, H, W, 3, generator=gg)
            s = {"RGB_im": cu(rgb), "rot": cu(rot), "trans": cu(trans), "new_traj": [t in (0, 5)] * b}
            res.append(model([[s], cam])["depth"].clone())
        outs[graph] = res
    for a, bb in zip(outs[False], outs[True]):
        assert torch.equal(a, bb)
    assert float(outs[True][0].min()) == 1000.0        # frame 0 is the new-trajectory pass-through


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_model_side_stream_preparation_is_the_same_computation(graph):
    """DepthEstimatorPyramid forks the feature preparation and SNCV of every level onto a side stream (they depend on the encoder
    output only): depth maps and every level's state are bit-identical to the serial order, eagerly and under CUDA graphs,
    across a trajectory reset, and with ablations that remove the SNCV / the normalisation."""
    m = _m4d()
    nl, b, H, W = 6, 2, 128, 192
    w = oracle.init_weights(nl, seed=3, bias_std=0.05, dn_random=True)
    cam = dev_cam(camera_for("kitti", b, H, W))
    for abl in (None, m.M4depthAblationParameters(SNCV=False), m.M4depthAblationParameters(normalize_features=False)):
        outs = {}
        for side in (0, 1, 2):           # serial | one fork after the encoder | per-level forks from inside the encoder
            model = m.M4Depth(nbre_levels=nl, use_cuda_graph=graph, ablation_settings=abl)
            model.d_estimator.side_stream_prep = side
            wts = dict(w)
            if abl is not None:       # the first refiner kernel's input width follows the switches (m4depth_network.py:223-242)
                for lvl in range(1, nl + 1):
                    cuts = 2 ** (lvl // 2)
                    cin = 9 * cuts + 1 + 4 + (49 * cuts if abl.SNCV else 0) + 1
                    gk = torch.Generator().manual_seed(100 + lvl)
                    wts[f"d_estimator/levels/{lvl - 1}/disp_refiner/prep_conv_layers/0/kernel"] = \
                        torch.randn(3, 3, cin, 128, generator=gk) * (2.0 / (9 * cin)) ** 0.5
            model.load_weights(wts)
            gg = torch.Generator().manual_seed(31)
            res = []
            for t in range(8):
                rot, trans = motion(gg, b)
                rgb = torch.rand(b, H, W, 3, generator=gg)
                s = {"RGB_im": cu(rgb), "rot": cu(rot), "trans": cu(trans), "new_traj": [t in (0, 4)] * b}
                res.append(model([[s], cam])["depth"].clone())
            torch.cuda.synchronize()
            res += [lvl.depth_prev_t.clone() for lvl in model.d_estimator.levels]
            res += [lvl.prev_f_maps.clone() for lvl in model.d_estimator.levels]
            outs[side] = res
        for side in (1, 2):
            for a, bb in zip(outs[0], outs[side]):
                assert torch.equal(a, bb)


# ----------------------------------------------------- whole model at the BASELINE.json sizes, bounded by oracle-vs-oracle
def _synth():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.join(root, "tools") not in sys.path:
        sys.path.insert(0, os.path.join(root, "tools"))
    import synth
    return synth


def _report(name, rows):
    """Per-frame error table of a whole-model comparison -> gpurun_out/ (numbers quoted in DESIGN.md)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"model_parity_{name}.json"), "w") as f:
        json.dump(rows, f, indent=1)


def _gpu_sequence(m, frames, cam, nl, new_traj_at=(0,), graph=True, weights_seed=1, weights=None):
    model = m.M4Depth(nbre_levels=nl, use_cuda_graph=graph)
    model.load_weights(weights if weights is not None else oracle.init_weights(nl, seed=weights_seed, bias_std=0.05, dn_random=True))
    dcam = dev_cam(cam)
    outs = []
    for t, fr in enumerate(frames):
        b = fr["RGB_im"].shape[0]
        s = {"RGB_im": cu(fr["RGB_im"]), "rot": cu(fr["rot"]), "trans": cu(fr["trans"]), "new_traj": [t in new_traj_at] * b}
        outs.append(model([[s], dcam])["depth"].cpu().clone())
    return outs


def _oracle_sequence(frames, cam, nl, mode, new_traj_at=(0,), weights_seed=1, weights=None):
    w = weights if weights is not None else oracle.init_weights(nl, seed=weights_seed, bias_std=0.05, dn_random=True)
    outs = []
    with oracle.reduction_order(mode), torch.no_grad():
        model = oracle.M4Depth(w, nbre_levels=nl, pscv_kwargs={"use_cuda_backproject": False})
        for t, fr in enumerate(frames):
            s = dict(fr)
            s["new_traj"] = torch.tensor([t in new_traj_at] * fr["RGB_im"].shape[0])
            outs.append(model([[s], cam])["depth"].clone())
    return outs


QS = (0.5, 0.9, 0.99, 0.999)


def _err(x, y):
    e = ((x - y).abs() / (y.abs() + 0.1)).flatten().double()
    return [float(torch.quantile(e, q)) for q in QS], float(e.max()), float((e <= 1e-4).double().mean())


def _check_against_oracle_pair(name, gpu, ref, alt, f64, new_traj_at=(0,)):
    """GPU-vs-oracle error against the error between CPU evaluations of the same graph that differ only in how the fp32
    reductions are summed (tests/test_oracle_golden.py::test_summation_order_alone_moves_the_depth_maps).
      * pass-through frame of a trajectory start: exact;
      * first estimated frame: 1e-4 (north_star) at the 99.9th percentile and for at least 99.9 % of the pixels;
      * every estimated frame: every quantile (50 / 90 / 99 / 99.9 %) of the GPU error within 4x the larger of the two
        oracle-vs-oracle errors at that quantile, the maximum (one pixel: a noisy statistic) within 10x, the fraction of pixels
        inside 1e-4 within 3 points of the oracles'.
    Why 4x and not 1x: besides summation order the GPU path differs from the torch oracle in expf / logf / rsqrt (<= 2 ulp
    each) and in the 3xFP16 split products of the tensor-core convolutions (2^-22 relative per product); measured on B200 the
    ratio is 2-3x at every quantile and does not grow with the frame index (gpurun_out/model_parity_*.json)."""
    rows = []
    since = 0
    for t in range(len(gpu)):
        since = 0 if t in new_traj_at else since + 1
        qg, mg, ing = _err(gpu[t], ref[t])
        qa, ma, ina = _err(alt[t], ref[t])
        qb, mb, inb = _err(f64[t], ref[t])
        rows.append({"frame": t, "frames_since_reset": since, "gpu_vs_oracle": {"q": qg, "max": mg, "within_1e-4": ing},
                     "reordered_vs_oracle": {"q": qa, "max": ma, "within_1e-4": ina}, "fp64_vs_oracle": {"q": qb, "max": mb, "within_1e-4": inb}})
    _report(name, rows)
    K_ = 4.0
    for r in rows:
        t, g_ = r["frame"], r["gpu_vs_oracle"]
        omax = max(r["reordered_vs_oracle"]["max"], r["fp64_vs_oracle"]["max"])
        if r["frames_since_reset"] == 0:
            assert g_["max"] == 0.0, (name, t, g_)
        else:
            if r["frames_since_reset"] == 1:
                assert g_["within_1e-4"] >= 0.999 and g_["q"][3] <= 1e-4, (name, t, g_)
            for i, q in enumerate(QS):
                bound = K_ * max(r["reordered_vs_oracle"]["q"][i], r["fp64_vs_oracle"]["q"][i]) + 2e-6
                assert g_["q"][i] <= bound, (name, t, q, g_["q"][i], bound)
            assert g_["max"] <= 10.0 * omax + 1e-4, (name, t, g_["max"], omax)
            assert g_["within_1e-4"] >= min(r["reordered_vs_oracle"]["within_1e-4"], r["fp64_vs_oracle"]["within_1e-4"]) - 0.03, (name, t)


BASELINE_MODEL_CASES = [  # (name, camera, b, H, W, frames): BASELINE.json configs[1], [2] (two of the eight sequences), [4]
    ("cfg1_384x384", "midair", 1, 384, 384, 5),
    ("cfg2_384x1280", "kitti", 2, 384, 1280, 5),
    ("cfg4_480x640", "tartan", 1, 480, 640, 5),
]


@pytest.mark.parametrize("case", BASELINE_MODEL_CASES, ids=[c[0] for c in BASELINE_MODEL_CASES])
def test_model_vs_oracle_at_baseline_configs(case):
    """Whole model (6 levels, CUDA graphs) against the oracle model at the BASELINE.json image sizes, four estimated frames."""
    m = _m4d()
    name, kind, b, H, W, n = case
    frames, cam = _synth().synth_sequence(n, b, H, W, kind, seed=1234 + H + W)
    gpu = _gpu_sequence(m, frames, cam, 6)
    ref = _oracle_sequence(frames, cam, 6, "default")
    alt = _oracle_sequence(frames, cam, 6, "reordered")
    f64 = _oracle_sequence(frames, cam, 6, "fp64")
    _check_against_oracle_pair(name, gpu, ref, alt, f64)


def test_model_batch_independence_at_headline_config():
    """BASELINE configs[2] as benched (384x1280, b = 8): each of the eight sequences gets bit for bit the depth maps it gets in
    a batch of two (the pair the oracle comparison above runs), frame by frame - no kernel mixes batch elements."""
    m = _m4d()
    frames, cam = _synth().synth_sequence(4, 2, 384, 1280, "kitti", seed=1234 + 384 + 1280)
    two = _gpu_sequence(m, frames, cam, 6)
    rep = lambda t: t.repeat(4, *([1] * (t.dim() - 1)))
    frames8 = [{k: rep(v) for k, v in fr.items()} for fr in frames]
    cam8 = {k: rep(v) for k, v in cam.items()}
    eight = _gpu_sequence(m, frames8, cam8, 6)
    for a_, b_ in zip(two, eight):
        assert torch.equal(rep(a_).view(torch.int32), b_.view(torch.int32))


def test_model_16_frame_stream_drift():
    """BASELINE configs[4]-shaped stream (TartanAir camera, 16 frames, a trajectory restart at frame 9; 240x320 so that the three
    CPU evaluations stay within a minute): the recurrent state (m4depth_network.py:160-163) carried over many frames stays
    within the oracle-vs-oracle envelope at every frame, and the restart brings the error back to the first-frame bound."""
    m = _m4d()
    frames, cam = _synth().synth_sequence(16, 1, 240, 320, "tartan", seed=4242)
    at = (0, 9)
    gpu = _gpu_sequence(m, frames, cam, 6, new_traj_at=at)
    ref = _oracle_sequence(frames, cam, 6, "default", new_traj_at=at)
    alt = _oracle_sequence(frames, cam, 6, "reordered", new_traj_at=at)
    f64 = _oracle_sequence(frames, cam, 6, "fp64", new_traj_at=at)
    _check_against_oracle_pair("stream16_240x320", gpu, ref, alt, f64, new_traj_at=at)


@pytest.mark.parametrize("which", ["midair", "kitti"])
def test_model_with_the_reference_checkpoints_vs_oracle(which):
    """The reference's shipped weights (pretrained_weights.zip -> tests/golden/_real/weights_*.npz, extracted by
    __graft_entry__.build() with the TensorFlow-free bundle reader) through the CUDA path against the oracle with the same
    weights: Mid-Air weights on a Mid-Air-shaped 384x384 stream, KITTI weights on a KITTI-shaped 256x768 stream (the network
    input sizes of dataloaders/midair.py:13 and kitti.py:14), five plane-warped synthetic frames, same bound as the random-weight
    cases.  Trained weights make the recurrence contractive: the report shows how much closer the evaluations stay."""
    m = _m4d()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "tests", "golden", "_real", f"weights_{which}.npz")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: __graft_entry__.build() extracts it from /root/reference/pretrained_weights.zip")
    w = {k: torch.from_numpy(v) for k, v in np.load(path).items()}
    kind, H, W = ("midair", 384, 384) if which == "midair" else ("kitti", 256, 768)
    frames, cam = _synth().synth_sequence(5, 1, H, W, kind, seed=99)
    gpu = _gpu_sequence(m, frames, cam, 6, weights=w)
    ref = _oracle_sequence(frames, cam, 6, "default", weights=w)
    alt = _oracle_sequence(frames, cam, 6, "reordered", weights=w)
    f64 = _oracle_sequence(frames, cam, 6, "fp64", weights=w)
    assert all(torch.isfinite(x).all() for x in gpu)
    _check_against_oracle_pair(f"real_{which}", gpu, ref, alt, f64)


def test_load_weights_without_dn_variables_and_with_wrong_layout():
    """A checkpoint trained with DINL off has no dn variables (Keras builds the layer on first call): it loads into a model
    with DINL off, is refused with a clear error by a model that needs them, and a kernel whose input channels do not fit
    the ablation layout is refused at load time, not at the first forward call."""
    m = _m4d()
    w = oracle.init_weights(3, seed=2)
    no_dn = {k: v for k, v in w.items() if "dn_layers" not in k}
    ab = m.M4depthAblationParameters(DINL=False)
    model = m.M4Depth(nbre_levels=3, ablation_settings=ab, use_cuda_graph=False)
    model.load_weights(no_dn)
    with pytest.raises(m.M4DError, match="missing"):
        m.M4Depth(nbre_levels=3, use_cuda_graph=False).load_weights(no_dn)
    with pytest.raises(m.M4DError, match="input channels"):
        m.M4Depth(nbre_levels=3, ablation_settings=m.M4depthAblationParameters(SNCV=False), use_cuda_graph=False).load_weights(w)


def test_inputs_consumed_event_orders_host_buffer_reuse():
    """Pinned host inputs: call() returns before the uploads ran; inputs_consumed() is the event after which the caller may
    refill the same buffers.  Overwriting them only after that event must not change the result."""
    m = _m4d()
    frames, cam = _synth().synth_sequence(4, 2, 128, 192, "kitti", seed=5)
    w = oracle.init_weights(6, seed=1, bias_std=0.05, dn_random=True)
    want = _gpu_sequence(m, frames, cam, 6, weights=w)
    model = m.M4Depth(nbre_levels=6, use_cuda_graph=True)
    model.load_weights(w)
    buf = {k: torch.empty_like(v).pin_memory() for k, v in frames[0].items()}
    hcam = {k: v.clone().pin_memory() for k, v in cam.items()}
    for t, fr in enumerate(frames):
        for k in buf:
            buf[k].copy_(fr[k])
        out = model([[{"RGB_im": buf["RGB_im"], "rot": buf["rot"], "trans": buf["trans"], "new_traj": [t == 0] * 2}], hcam])
        model.inputs_consumed().synchronize()
        for k in buf:
            buf[k].fill_(float("nan"))                     # the loader reuses the buffers immediately
        assert torch.equal(out["depth"].cpu(), want[t])


@pytest.mark.parametrize("cfg", [(1, 15, 20, 128, 192), (2, 33, 47, 64, 96), (1, 30, 41, 96, 128), (2, 7, 9, 16, 32)])
def test_stride2_conv_on_odd_sizes_runs_on_the_tensor_cores(cfg):
    """Keras Conv2D(strides=2, 'same') on an odd dimension pads one pixel on both sides (BASELINE configs[4]: 15 -> 8
    at level 6).  The layer shifts the input into a zeroed even-sized buffer (m4d_pad_shift) and runs the tcgen05 stride-2
    kernel: same results as the oracle convolution, and no FFMA2 fallback is recorded."""
    m = _m4d()
    from m4depth_b200 import m4depth_network as net
    b, h, w, cin, cout = cfg
    g = torch.Generator().manual_seed(h * w + cin)
    x = torch.randn(b, h, w, cin, generator=g)
    k = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    want = oracle.leaky_relu(oracle.conv2d_same(x, k, bias, 2))
    layer = net._Conv2D(cout, 2)
    layer.assign(k, bias, "cuda")
    assert layer.packed is not None
    before = dict(net.conv_fallbacks)
    got = layer(cu(x), alpha=0.1)
    assert dict(net.conv_fallbacks) == before
    assert tuple(got.shape) == tuple(want.shape)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    # twice through the same layer: the zero border of the padded buffer is never written
    got2 = layer(cu(x * 2.0), alpha=0.1)
    np.testing.assert_allclose(got2.cpu().numpy(), oracle.leaky_relu(oracle.conv2d_same(x * 2.0, k, bias, 2)).numpy(), rtol=1e-5,
                               atol=1e-5 * float(want.abs().max()) * 2)


def test_fetch_depth_overlaps_the_next_frame_without_tearing():
    """M4Depth.fetch_depth: the depth map of frame t read into pinned host memory while frame t+1 is already running must be
    frame t's map, bit for bit (it leaves through a device staging buffer, not the live output buffer)."""
    m = _m4d()
    frames, cam = _synth().synth_sequence(6, 2, 128, 192, "kitti", seed=9)
    w = oracle.init_weights(6, seed=1, bias_std=0.05, dn_random=True)
    want = _gpu_sequence(m, frames, cam, 6, weights=w)
    model = m.M4Depth(nbre_levels=6, use_cuda_graph=True)
    model.load_weights(w)
    dcam = dev_cam(cam)
    hosts = [torch.empty(2, 128, 192, 1).pin_memory() for _ in frames]
    events = []
    for t, fr in enumerate(frames):
        model([[{"RGB_im": cu(fr["RGB_im"]), "rot": cu(fr["rot"]), "trans": cu(fr["trans"]), "new_traj": [t == 0] * 2}], dcam])
        events.append(model.fetch_depth(hosts[t]))           # no synchronisation: the next frame is enqueued right behind
    for t, ev in enumerate(events):
        ev.synchronize()
        assert torch.equal(hosts[t], want[t])

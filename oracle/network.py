"""Oracle: network layers and the per-level state machine (test infrastructure, see oracle/__init__.py).

Restates ``/root/reference/m4depth_network.py``:
  DomainNormalization 24-48 · FeaturePyramid 51-90 · DispRefiner 93-135 · DepthEstimatorLevel 138-262 ·
  DepthEstimatorPyramid 265-323 · M4Depth.call 351-369,
with the TF 2.7 primitive semantics they rely on (Keras Conv2D ``padding='same'`` incl. the
asymmetric stride-2 padding, ``tf.compat.v1.image.resize_bilinear`` legacy sampling,
``tf.image.resize(NEAREST)``, ``tf.linalg.normalize`` without epsilon, ``tf.math.l2_normalize``).

Weights are a flat dict keyed like the reference checkpoints' object graph
(``encoder/conv_layers_s1/0/kernel`` ..., kernels HWIO ``[3,3,Cin,Cout]``).
"""
from collections import namedtuple
import math

import torch

from ._ieee import sqrt as ieee_sqrt
import torch.nn.functional as TF

from .geometry import parallax2depth, prev_d2para
from .cost_volumes import get_parallax_sweeping_cv, cost_volume

F32 = torch.float32

# m4depth_network.py:21-22
M4depthAblationParameters = namedtuple(
    'M4depthAblationParameters',
    ('DINL', 'SNCV', 'time_recurr', 'normalize_features', 'subdivide_features', 'level_memory'),
    defaults=(True, True, True, True, True, True))

ENC_CHANNELS = [16, 32, 64, 96, 128, 192]          # m4depth_network.py:59
PREP_CHANNELS = [128, 128, 96]                     # :102
EST_CHANNELS = [64, 32, 16, 5]                     # :109


def level_channels(lvl_depth):
    """(c, cuts, refiner_cin) for level ``lvl_depth`` in 1..6 (m4depth_network.py:59,174,223-242)."""
    c = ENC_CHANNELS[lvl_depth - 1]
    cuts = 2 ** (lvl_depth // 2)
    return c, cuts, 9 * cuts + 1 + 4 + 49 * cuts + 1


def leaky_relu(x, alpha=0.1):
    return torch.where(x >= 0, x, x * alpha)


def same_padding(size, stride, k=3):
    """TF 'SAME': out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total // 2."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2, total - total // 2


# Summation order of the fp32 reductions whose order the reference does not define (Conv2D, the DomainNormalization means).
# "default": torch's conv2d / mean.  "reordered": the same sums in another order (tap-major matmuls, row-then-column means) -
# a second legitimate fp32 evaluation of the same graph; tests/test_oracle_golden.py and the GPU whole-model tests measure the
# GPU-vs-oracle error against the oracle-vs-reordered-oracle error (what summation order alone does to the depth maps).
# "fp64": every such reduction in double precision, rounded once (the least-error evaluation).
REDUCTION_ORDER = "default"


class reduction_order:
    """Context manager: ``with oracle.network.reduction_order("reordered"): ...``"""

    def __init__(self, mode):
        assert mode in ("default", "reordered", "fp64"), mode
        self.mode = mode

    def __enter__(self):
        global REDUCTION_ORDER
        self.prev, REDUCTION_ORDER = REDUCTION_ORDER, self.mode

    def __exit__(self, *a):
        global REDUCTION_ORDER
        REDUCTION_ORDER = self.prev


def conv2d_same(x, kernel, bias, stride=1):
    """Keras Conv2D(k=3, padding='same') on NHWC input, HWIO kernel, bias add (no activation)."""
    b, h, w, cin = x.shape
    oh, pt, pb = same_padding(h, stride)
    ow, pl, pr = same_padding(w, stride)
    if REDUCTION_ORDER == "reordered":
        # tap-major: nine [pixels, cin] x [cin, cout] products accumulated in fp32 in raster order of the taps
        xp = TF.pad(x, (0, 0, pl, pr, pt, pb))
        y = torch.zeros(b, oh, ow, kernel.shape[3], dtype=F32)
        for ky in range(3):
            for kx in range(3):
                sl = xp[:, ky:ky + (oh - 1) * stride + 1:stride, kx:kx + (ow - 1) * stride + 1:stride, :]
                y = y + (sl.reshape(-1, cin) @ kernel[ky, kx]).reshape(b, oh, ow, -1)
        return y + bias.view(1, 1, 1, -1)
    xn = TF.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    wt = kernel.permute(3, 2, 0, 1).contiguous()
    if REDUCTION_ORDER == "fp64":
        y = TF.conv2d(xn.double(), wt.double(), bias.double(), stride=stride).to(F32)
    else:
        y = TF.conv2d(xn, wt, bias, stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def _mean_hw(t):
    """mean over H and W (keepdim) in the summation order REDUCTION_ORDER selects."""
    if REDUCTION_ORDER == "reordered":
        return t.mean(dim=2, keepdim=True).mean(dim=1, keepdim=True)
    if REDUCTION_ORDER == "fp64":
        return t.double().mean(dim=(1, 2), keepdim=True).to(F32)
    return t.mean(dim=(1, 2), keepdim=True)


def resize_bilinear_legacy(x, out_h, out_w):
    """tf.compat.v1.image.resize_bilinear(align_corners=False, half_pixel_centers=False).

    src = dst * (in/out); lo = floor(src); hi = min(ceil(src), in-1); l = src - lo;
    top = TL + (TR-TL)*lx; bot = BL + (BR-BL)*lx; out = top + (bot-top)*ly   (SURVEY A.12).
    """
    b, h, w, c = x.shape

    def axis(in_size, out_size):
        scale = torch.tensor(in_size / out_size, dtype=F32) if out_size > 0 else torch.tensor(0., dtype=F32)
        src = torch.arange(out_size, dtype=F32) * scale
        lo_f = torch.floor(src)
        lo = lo_f.to(torch.int64).clamp(min=0)
        hi = torch.ceil(src).to(torch.int64).clamp(max=in_size - 1)
        return lo, hi, src - lo_f

    ylo, yhi, ly = axis(h, out_h)
    xlo, xhi, lx = axis(w, out_w)
    lx = lx.view(1, 1, out_w, 1)
    ly = ly.view(1, out_h, 1, 1)
    tl = x[:, ylo][:, :, xlo]
    tr = x[:, ylo][:, :, xhi]
    bl = x[:, yhi][:, :, xlo]
    br = x[:, yhi][:, :, xhi]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return top + (bot - top) * ly


def resize_nearest(x, out_h, out_w):
    """tf.image.resize(method=NEAREST) (TF2: half-pixel centres): src = min(floor((dst+.5)*in/out), in-1)."""
    b, h, w, c = x.shape

    def axis(in_size, out_size):
        scale = torch.tensor(in_size / out_size, dtype=F32)
        src = torch.floor((torch.arange(out_size, dtype=F32) + 0.5) * scale).to(torch.int64)
        return src.clamp(max=in_size - 1)

    return x[:, axis(h, out_h)][:, :, axis(w, out_w)]


def group_l2_normalize(f, nbre_cuts):
    """reshape [b,h,w,cuts,-1]; x / sqrt(sum(x^2)) per group, no epsilon (m4depth_network.py:180-186)."""
    b, h, w, c = f.shape
    g = f.reshape(b, h, w, nbre_cuts, c // nbre_cuts)
    n = ieee_sqrt((g * g).sum(dim=-1, keepdim=True))
    return (g / n).reshape(b, h, w, c)


class DomainNormalization:
    """m4depth_network.py:24-48.  (f-mean)/(var+1e-12) [variance, not sigma], channel l2_normalize, affine."""

    def __init__(self, scale, bias):
        self.scale = scale.reshape(1, 1, 1, -1).to(F32)
        self.bias = bias.reshape(1, 1, 1, -1).to(F32)

    def __call__(self, f_map):
        mean = _mean_hw(f_map)
        dev = f_map - mean
        var = _mean_hw(dev * dev)                              # tf.math.reduce_variance
        g = dev / (var + 1e-12)
        sq = (g * g).sum(dim=-1, keepdim=True)
        normed = g * torch.rsqrt(torch.clamp(sq, min=1e-12))   # tf.math.l2_normalize
        return self.scale * normed + self.bias

    call = __call__


class FeaturePyramid:
    """m4depth_network.py:51-90."""

    def __init__(self, settings, weights):
        self.use_dinl = settings["ablation"].DINL
        self.n = settings["nbre_lvls"]
        self.w = weights
        self.dn = DomainNormalization(weights["encoder/dn_layers/0/scale"], weights["encoder/dn_layers/0/bias"])

    def __call__(self, images):
        f = images
        outs = []
        for i in range(self.n):
            t = conv2d_same(f, self.w[f"encoder/conv_layers_s1/{i}/kernel"], self.w[f"encoder/conv_layers_s1/{i}/bias"], 1)
            if self.use_dinl and i == 0:
                t = self.dn(t)
            t = leaky_relu(t)
            t = conv2d_same(t, self.w[f"encoder/conv_layers_s2/{i}/kernel"], self.w[f"encoder/conv_layers_s2/{i}/bias"], 2)
            f = leaky_relu(t)
            outs.append(f)
        return outs

    call = __call__


class DispRefiner:
    """m4depth_network.py:93-135.  Returns [t5, t96]: the 2nd element is the untouched prep output."""

    def __init__(self, weights, prefix):
        self.w = weights
        self.p = prefix

    def __call__(self, feature_map):
        x = feature_map
        for j in range(len(PREP_CHANNELS)):
            x = leaky_relu(conv2d_same(x, self.w[f"{self.p}/prep_conv_layers/{j}/kernel"],
                                       self.w[f"{self.p}/prep_conv_layers/{j}/bias"], 1))
        prep = x
        for j in range(len(EST_CHANNELS)):
            x = conv2d_same(x, self.w[f"{self.p}/est_d_conv_layers/{j}/kernel"],
                            self.w[f"{self.p}/est_d_conv_layers/{j}/bias"], 1)
            if j < len(EST_CHANNELS) - 1:
                x = leaky_relu(x)
        return [x, prep]

    call = __call__


class DepthEstimatorLevel:
    """m4depth_network.py:138-262 (inference mode: state held in prev_f_maps / depth_prev_t)."""

    def __init__(self, settings, depth, weights, pscv_kwargs=None):
        self.ablation = settings["ablation"]
        self.is_training = settings["is_training"]
        self.lvl_depth = depth
        self.lvl_mul = depth - 3
        self.disp_refiner = DispRefiner(weights, f"d_estimator/levels/{depth - 1}/disp_refiner")
        self.prev_f_maps = None
        self.depth_prev_t = None
        self.pscv_kwargs = pscv_kwargs or {}
        self.trace = None          # optional dict filled with intermediates (tests)

    def __call__(self, curr_f_maps, prev_l_est, rot, trans, camera, new_traj, prev_f_maps=None, prev_t_depth=None):
        b, h, w, c = curr_f_maps.shape
        nbre_cuts = 2 ** (self.lvl_depth // 2) if self.ablation.subdivide_features else 1
        if self.ablation.normalize_features:
            curr_f_maps = group_l2_normalize(curr_f_maps, nbre_cuts)
            if prev_f_maps is not None:
                prev_f_maps = group_l2_normalize(prev_f_maps, nbre_cuts)

        if (not self.is_training) and prev_f_maps is None and prev_t_depth is None:
            prev_t_depth = self.depth_prev_t
            prev_f_maps = self.prev_f_maps

        if prev_l_est is None:
            para_prev_l = torch.ones(b, h, w, 1, dtype=F32)
            depth_prev_l = 1000. * torch.ones(b, h, w, 1, dtype=F32)
            other_prev_l = torch.zeros(b, h, w, 4, dtype=F32)
        else:
            other_prev_l = resize_bilinear_legacy(prev_l_est["other"], h, w)
            para_prev_l = resize_bilinear_legacy(prev_l_est["parallax"], h, w) * 2.
            depth_prev_l = resize_bilinear_legacy(prev_l_est["depth"], h, w)

        if prev_t_depth is None or bool(new_traj[0]):
            prev_t_depth = torch.ones(b, h, w, 1, dtype=F32) * 1000.
            if not self.is_training:
                self.prev_f_maps = curr_f_maps
                self.depth_prev_t = prev_t_depth
            return {"depth": depth_prev_l, "parallax": para_prev_l, "other": other_prev_l}

        scale = 2.0 ** self.lvl_mul
        para_prev_t = prev_d2para(prev_t_depth, rot, trans, camera)
        cv, para_prev_t_reproj = get_parallax_sweeping_cv(
            curr_f_maps, prev_f_maps, para_prev_t, para_prev_l, rot, trans, camera, 4,
            nbre_cuts=nbre_cuts, **self.pscv_kwargs)
        feats = [cv, torch.log(para_prev_l * scale)]
        if self.ablation.level_memory:
            feats.append(other_prev_l)
        if self.ablation.SNCV:
            feats.append(cost_volume(curr_f_maps, curr_f_maps, 3, nbre_cuts=nbre_cuts))
        if self.ablation.time_recurr:
            feats.append(torch.log(para_prev_t_reproj[..., 4:5] * scale))
        f_input = torch.cat(feats, dim=3)

        out = self.disp_refiner(f_input)[0]
        para = out[..., :1]
        other = out[..., 1:]
        para_curr_l = torch.exp(torch.clamp(para, -7., 7.)) / scale
        depth = parallax2depth(para_curr_l, rot, trans, camera)
        if self.trace is not None:
            self.trace.update(curr_f_maps=curr_f_maps, para_prev_t=para_prev_t, cv=cv,
                              prev_disp=para_prev_t_reproj, f_input=f_input, refiner_out=out,
                              para_prev_l=para_prev_l)
        if not self.is_training:
            self.prev_f_maps = curr_f_maps
            self.depth_prev_t = depth
        return {"other": other, "depth": depth, "parallax": para_curr_l}

    call = __call__


class DepthEstimatorPyramid:
    """m4depth_network.py:265-323 (inference: temporal state lives in the levels)."""

    def __init__(self, settings, weights, pscv_kwargs=None):
        self.levels = [DepthEstimatorLevel(settings, i + 1, weights, pscv_kwargs) for i in range(settings["nbre_lvls"])]
        self.is_training = settings["is_training"]

    def __call__(self, f_maps_pyrs, traj_samples, camera, training=False):
        d_est_seq = []
        for f_pyr_curr, sample in zip(f_maps_pyrs, traj_samples):
            rot, trans = sample['rot'], sample['trans']
            cnter = float(len(self.levels))
            d_est_curr = None
            for l, (f_maps_curr, level) in enumerate(zip(f_pyr_curr[::-1], self.levels[::-1])):
                local_camera = {"f": camera["f"] / 2. ** cnter, "c": camera["c"] / 2. ** cnter}
                d_est = d_est_curr[-1].copy() if l != 0 else None
                est = level(f_maps_curr, d_est, rot, trans, local_camera, sample["new_traj"])
                d_est_curr = [est] if d_est_curr is None else d_est_curr + [est]
                cnter -= 1.
            d_est_seq.append(d_est_curr[::-1])
        return d_est_seq

    call = __call__


class M4Depth:
    """m4depth_network.py:325-369 (inference call only)."""

    def __init__(self, weights, depth_type="map", nbre_levels=6, is_training=False, ablation_settings=None,
                 pscv_kwargs=None):
        self.ablation_settings = ablation_settings or M4depthAblationParameters()
        self.model_settings = {"nbre_lvls": nbre_levels, "is_training": is_training, "ablation": self.ablation_settings}
        self.encoder = FeaturePyramid(self.model_settings, weights)
        self.d_estimator = DepthEstimatorPyramid(self.model_settings, weights, pscv_kwargs)

    def __call__(self, data, training=False):
        traj_samples, camera = data
        f_maps_pyrs = [self.encoder(s['RGB_im']) for s in traj_samples]
        d_maps_pyrs = self.d_estimator(f_maps_pyrs, traj_samples, camera, training)
        if training:
            return d_maps_pyrs
        h, w = traj_samples[-1]['RGB_im'].shape[1:3]
        return {"depth": resize_nearest(d_maps_pyrs[-1][0]["depth"], h, w)}

    call = __call__


def init_weights(nbre_levels=6, seed=7, bias_std=0.0, dn_random=False):
    """Synthetic stand-in for the checkpoints; key layout matches SURVEY.md section 5.

    He-normal kernels (Keras HeNormal is a truncated normal with stddev sqrt(2/fan_in); a plain normal
    with that stddev is used here), biases zero like Keras' default (m4depth_network.py:35-38,61) or
    N(0, bias_std) so that bias handling is exercised, DN scale=1/bias=0 or randomised."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(name, cin, cout):
        std = math.sqrt(2.0 / (9 * cin))
        w[name + "/kernel"] = torch.randn(3, 3, cin, cout, generator=g, dtype=F32) * std
        w[name + "/bias"] = torch.randn(cout, generator=g, dtype=F32) * bias_std

    cin = 3
    for i in range(nbre_levels):
        conv(f"encoder/conv_layers_s1/{i}", cin, ENC_CHANNELS[i])
        conv(f"encoder/conv_layers_s2/{i}", ENC_CHANNELS[i], ENC_CHANNELS[i])
        cin = ENC_CHANNELS[i]
    if dn_random:
        w["encoder/dn_layers/0/scale"] = torch.rand(1, 1, 1, ENC_CHANNELS[0], generator=g, dtype=F32) + 0.5
        w["encoder/dn_layers/0/bias"] = torch.randn(1, 1, 1, ENC_CHANNELS[0], generator=g, dtype=F32) * 0.1
    else:
        w["encoder/dn_layers/0/scale"] = torch.ones(1, 1, 1, ENC_CHANNELS[0], dtype=F32)
        w["encoder/dn_layers/0/bias"] = torch.zeros(1, 1, 1, ENC_CHANNELS[0], dtype=F32)
    for lvl in range(1, nbre_levels + 1):
        _, _, cin = level_channels(lvl)
        p = f"d_estimator/levels/{lvl - 1}/disp_refiner"
        for j, co in enumerate(PREP_CHANNELS):
            conv(f"{p}/prep_conv_layers/{j}", cin, co)
            cin = co
        for j, co in enumerate(EST_CHANNELS):
            conv(f"{p}/est_d_conv_layers/{j}", cin, co)
            cin = co
    return w

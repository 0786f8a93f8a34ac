"""Host-side mirror of the reference's ``m4depth_network.py`` layer API for INFERENCE, backed by libm4d (sm_100a).

Same class names, constructor arguments, ``call`` signatures and return structures as the reference
(m4depth_network.py:24-369), on torch CUDA tensors (NHWC fp32) instead of TF tensors:

    DomainNormalization(regularizer_weight).call(f_map)
    FeaturePyramid(settings, regularizer_weight, trainable).call(images) -> list[nbre_lvls]
    DispRefiner(regularizer_weight).call(feature_map) -> [t5, t96]
    DepthEstimatorLevel(settings, depth, regularizer_weight).call(curr_f_maps, prev_l_est, rot, trans, camera,
                        new_traj, prev_f_maps=None, prev_t_depth=None) -> {"depth","parallax","other"}
    DepthEstimatorPyramid(settings, ...).call(f_maps_pyrs, traj_samples, camera, training=False)
    M4Depth(depth_type, nbre_levels, is_training, ablation_settings).call([traj_samples, camera], training=False)

Differences a maintainer must know:
  * inference only (``is_training=True`` / ``training=True`` raise): the north-star path has no backward pass;
  * weights are assigned with ``load_weights(dict)`` keyed like the reference checkpoints' object graph;
  * returned tensors are views into per-layer workspaces that the next call overwrites (clone to keep) - this is
    what lets a whole frame be one CUDA graph (``M4Depth(use_cuda_graph=True)``, the default);
  * ``new_traj`` is read on the host (bool / list / CPU tensor); a CUDA tensor costs a device sync.
"""
from collections import namedtuple

import os

import torch

from . import _lib as L

# conv kernel choice for layers called with algo=0: 0 = auto (tcgen05 where packed), 1 = FFMA2 everywhere (tools/model_error.py)
DEFAULT_CONV_ALGO = int(os.environ.get("M4D_CONV_ALGO", "0"))

# m4depth_network.py:21-22
M4depthAblationParameters = namedtuple(
    'M4depthAblationParameters',
    ('DINL', 'SNCV', 'time_recurr', 'normalize_features', 'subdivide_features', 'level_memory'),
    defaults=(True, True, True, True, True, True))

LEAKY = 0.1

# How often a convolution that could have used the tensor cores did not (odd-sized stride-2 input, unaligned or strided view):
# {reason: count}.  The FFMA2 kernel gives the same results at a fraction of the rate; m4depth_b200.m4depth_network.conv_fallbacks
# lets a caller see it (the first occurrence of each reason is also logged once).
conv_fallbacks = {}


def _note_fallback(reason, shape):
    n = conv_fallbacks.get(reason, 0)
    conv_fallbacks[reason] = n + 1
    if n == 0:
        import warnings
        warnings.warn(f"m4depth_b200: convolution off the tcgen05 path ({reason}; input {tuple(shape)}): FFMA2 kernel used", stacklevel=3)


def _nvtx(name):
    """NVTX range around a host-side section (eager passes and CUDA-graph capture; graph replays launch nothing from Python)."""
    return torch.cuda.nvtx.range(name)


def _pix_stride(t):
    """Pixel stride (floats) of a [b,h,w,c] tensor laid out as rows of pixels, possibly inside a wider buffer."""
    b, h, w, c = t.shape
    sb, sh, sw, sc = t.stride()
    if t.dtype != torch.float32 or not t.is_cuda:
        raise L.M4DError("expected a float32 CUDA tensor")
    if c > 1 and sc != 1:
        raise L.M4DError("channel stride must be 1 (NHWC)")
    ps = sw
    if ps < c or (h > 1 and sh != w * ps) or (b > 1 and sb != h * w * ps):
        raise L.M4DError(f"tensor is not a dense NHWC pixel grid (shape {tuple(t.shape)}, strides {t.stride()})")
    return ps


def _new_traj_flag(new_traj):
    """m4depth_network.py:206-208: batch element 0 decides for the whole batch."""
    if isinstance(new_traj, (bool, int)):
        return bool(new_traj)
    if isinstance(new_traj, torch.Tensor):
        return bool(new_traj.reshape(-1)[0].item())
    return bool(new_traj[0])


# precision mode of the tensor-core convs (include/m4d.h M4D_CONV_PREC_*): 0 = 3xTF32, 1 = 3xFP16 (scaled fp16 hi/lo planes:
# same error class, twice the tensor-core rate).  M4D_CONV_PREC in the environment overrides the default.
DEFAULT_CONV_PREC = int(os.environ.get("M4D_CONV_PREC", "1"))
CONV_PDL, CONV_PDL_WEIGHTS_STABLE = 1 << 8, 1 << 9
USE_PDL = os.environ.get("M4D_CONV_PDL", "1") != "0"


class _Conv2D:
    """ks.layers.Conv2D(filters, 3, strides, padding='same') with fused bias + optional leaky_relu (libm4d)."""

    def __init__(self, filters, strides=1, prec=None):
        self.filters, self.strides = filters, strides
        self.prec = DEFAULT_CONV_PREC if prec is None else prec
        self.kernel = None          # [3,3,cin,cout] HWIO
        self.packed = None          # TF32 hi/lo planes for the tcgen05 path (m4d_conv3x3_tc_pack)
        self.tc_min_cin = 16        # thinner inputs (the RGB conv) stay on the FFMA2 kernel
        self.events = None          # optional list: receives (start, end) CUDA events around the launch (bench.py roofline)
        self.bias = None            # [cout]
        self.host = None            # (kernel, bias) on the host, made on demand (first encoder layer, weights as kernel parameters)
        self._out = {}
        self._calls = 0             # tensor-core launches since the weights were (re)packed

    def assign(self, kernel, bias, device):
        k = kernel.to(device=device, dtype=torch.float32).contiguous()
        if k.dim() != 4 or k.shape[0] != 3 or k.shape[1] != 3 or k.shape[3] != self.filters:
            raise L.M4DError(f"conv kernel must be [3,3,cin,{self.filters}], got {tuple(k.shape)}")
        self.kernel = k
        self.bias = bias.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        self._calls = 0
        # the first encoder layer's fused kernels take their 3x3x3x16 weights as kernel parameters: host copies made here, not
        # lazily inside a call (a device-to-host copy is not allowed while a CUDA graph is being captured)
        self.host = (k.detach().cpu().contiguous(), self.bias.detach().cpu().contiguous()) if k.shape[2] == 3 and self.filters == 16 else None
        # tensor-core path (stride 1, or stride 2 with cin % 16 == 0; cout <= 256): TF32 hi/lo planes packed once per layer
        self.packed = None
        n = L.lib.m4d_conv3x3_tc_packed_floats_p(k.shape[2], self.filters, self.strides, self.prec)
        if n > 0:
            self.packed = torch.empty(n, dtype=torch.float32, device=k.device)
            L.check(L.lib.m4d_conv3x3_tc_pack_p(L.ptr(k), k.shape[2], self.filters, self.strides, self.prec, L.ptr(self.packed), L.stream()))

    def out_shape(self, x):
        b, h, w, _ = x.shape
        s = self.strides
        return (b, -(-h // s), -(-w // s), self.filters)

    def __call__(self, x, alpha=1.0, out=None, algo=0, slices=0):
        if self.kernel is None:
            raise L.M4DError("conv layer has no weights: call load_weights() first")
        b, h, w, cin = x.shape
        if cin != self.kernel.shape[2]:
            raise L.M4DError(f"conv input has {cin} channels, kernel expects {self.kernel.shape[2]}")
        if out is None:
            key = (b, h, w)
            out = self._out.get(key)
            if out is None:
                out = self._out[key] = torch.empty(self.out_shape(x), dtype=torch.float32, device=x.device)
        xs, ys = _pix_stride(x), _pix_stride(out)
        algo = algo or DEFAULT_CONV_ALGO
        # algo: 0 = auto (tcgen05 3xTF32 where the layer was packed and the strides allow, else FFMA2), 1 = FFMA2, 2 = tcgen05
        tc_ok = self.packed is not None and cin >= self.tc_min_cin
        if tc_ok and not (xs % 4 == 0 and x.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0):
            tc_ok = False           # TMA needs 16-byte aligned bases and pixel strides
            _note_fallback("unaligned view", x.shape)
        if tc_ok and self.strides == 2 and not (h % 2 == 0 and w % 2 == 0 and xs == cin):
            # The 2x2-cell formulation needs even sizes (TF SAME then pads bottom / right only) and dense pixels.  An odd
            # dimension is padded one pixel on both sides: shifted by one pixel into a zeroed even-sized buffer it is the even case.
            if cin % 4 == 0 and algo != 1:
                sy, sx = h % 2, w % 2
                key = ("pad", b, h, w)
                pad = self._out.get(key)
                if pad is None:
                    pad = self._out[key] = torch.zeros(b, h + sy, w + sx, cin, dtype=torch.float32, device=x.device)
                L.check(L.lib.m4d_pad_shift(L.ptr(x), xs, b, h, w, cin, sy, sx, L.ptr(pad), L.stream()))
                x, h, w, xs = pad, h + sy, w + sx, cin
            else:
                tc_ok = False
                _note_fallback("stride 2 on a strided input with cin % 4 != 0", x.shape)
        if algo != 1 and tc_ok:
            if self.events is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            # programmatic dependent launch (include/m4d.h M4D_CONV_PDL): from a layer's second call on its packed weights are
            # older than whatever precedes the launch in the stream, so they may be fetched before the grid dependency resolves
            flags = (CONV_PDL | (CONV_PDL_WEIGHTS_STABLE if self._calls > 0 else 0)) if USE_PDL else 0
            self._calls += 1
            L.check(L.lib.m4d_conv3x3_tc_fwd_p(L.ptr(x), xs, L.ptr(self.packed), L.ptr(self.bias), b, h, w, cin, self.filters,
                                               self.strides, self.prec, float(alpha), L.ptr(out), ys, int(slices) | flags, L.stream()))
            if self.events is not None:
                ev1.record()
                self.events.append((ev0, ev1))
            return out
        if algo == 2:
            raise L.M4DError("conv layer is outside the tcgen05 path (cout <= 256, cin >= 16, input pixel stride % 4 == 0; "
                             "stride 2: even sizes, cin % 16 == 0, dense pixels)")
        L.check(L.lib.m4d_conv3x3_nhwc(L.ptr(x), xs, L.ptr(self.kernel), L.ptr(self.bias), b, h, w, cin,
                                       self.filters, self.strides, float(alpha), L.ptr(out), ys, 1, L.stream()))
        return out


class DomainNormalization:
    """m4depth_network.py:24-48 (Zhang et al., domain-invariant normalisation)."""

    def __init__(self, regularizer_weight=0.0004):
        self.regularizer_weight = regularizer_weight
        self.scale = None
        self.bias = None
        self._ws = {}

    def build(self, input_shape, device):
        c = input_shape[-1]
        self.scale = torch.ones(1, 1, 1, c, dtype=torch.float32, device=device)
        self.bias = torch.zeros(1, 1, 1, c, dtype=torch.float32, device=device)

    def call(self, f_map, leaky_alpha=1.0, out=None):
        L.f32c(f_map, "f_map")
        b, h, w, c = f_map.shape
        if self.scale is None:
            self.build(f_map.shape, f_map.device)
        key = (b, h, w, c)
        ws = self._ws.get(key)
        if ws is None:
            ws = self._ws[key] = (torch.empty(2 * b * c, dtype=torch.float64, device=f_map.device),
                                  torch.empty_like(f_map))
        stats, buf = ws
        out = buf if out is None else out
        L.check(L.lib.m4d_domain_norm(L.ptr(f_map), b, h, w, c, L.ptr(self.scale), L.ptr(self.bias), float(leaky_alpha),
                                      L.ptr(stats), L.ptr(out), L.stream()))
        return out

    __call__ = call


class FeaturePyramid:
    """Encoder (m4depth_network.py:51-90)."""

    def __init__(self, settings, regularizer_weight=0.0004, trainable=True):
        self.use_dinl = settings["ablation"].DINL
        self.out_sizes = [16, 32, 64, 96, 128, 192][:settings["nbre_lvls"]]
        self.conv_layers_s1 = [_Conv2D(n, 1) for n in self.out_sizes]
        self.conv_layers_s2 = [_Conv2D(n, 2) for n in self.out_sizes]
        self.dn_layers = [DomainNormalization(regularizer_weight) for _ in self.out_sizes]   # only [0] is used (:82-83)
        # conv -> DN as two ops (default) or m4d_rgb_conv_dn, which never stores the conv output but evaluates it twice.
        # Measured on B200 (config 3): the fused call is 0.18 ms per step SLOWER - the 3->16 conv is instruction-bound
        # (~400 instructions per pixel), not bound by the 252 MB it writes, so recomputing it costs more than the traffic saved.
        # M4D_FUSED_FIRST_LAYER: 3 (default) = the conv evaluated once with its weights in the kernel parameter block, output stored
        # and DN statistics accumulated on the way, then DN's apply pass (m4d_rgb_conv_stats_hostw + m4d_domain_norm_apply);
        # 2 = conv recomputed in both DN passes, weights as parameters (m4d_rgb_conv_dn_hostw); 1 = the same with the weights in
        # shared memory; 0 = conv (shared-memory weights), DN statistics, DN apply as three kernels.
        # Measured on B200 (config 3, frames/s, same box): 3: 1467-1471, 2: 1435-1452, 0: 1439-1451.
        self.first_layer_mode = int(os.environ.get("M4D_FUSED_FIRST_LAYER", "3"))
        self.unfused_first_layer = self.first_layer_mode == 0

    def _first_layer_fused(self, conv1, images):
        dn = self.dn_layers[0]
        if conv1.kernel is None:
            raise L.M4DError("conv layer has no weights: call load_weights() first")
        b, h, w, _ = images.shape
        if dn.scale is None:
            dn.build((b, h, w, 16), images.device)
        key = (b, h, w, 16)
        ws = dn._ws.get(key)
        if ws is None:
            ws = dn._ws[key] = (torch.empty(2 * b * 16, dtype=torch.float64, device=images.device),
                                torch.empty((b, h, w, 16), dtype=torch.float32, device=images.device))
        stats, out = ws
        if self.first_layer_mode == 3:
            # conv evaluated once (weights as kernel parameters), stored, statistics accumulated on the way; then DN's apply pass
            if conv1.host is None:
                conv1.host = (conv1.kernel.detach().cpu().contiguous(), conv1.bias.detach().cpu().contiguous())
            hk, hb = conv1.host
            y = dn._ws.get(("y",) + key)
            if y is None:
                y = dn._ws[("y",) + key] = torch.empty((b, h, w, 16), dtype=torch.float32, device=images.device)
            L.check(L.lib.m4d_rgb_conv_stats_hostw(L.ptr(images), _pix_stride(images), hk.data_ptr(), hb.data_ptr(), b, h, w,
                                                   L.ptr(y), L.ptr(stats), L.stream()))
            L.check(L.lib.m4d_domain_norm_apply(L.ptr(y), b, h, w, 16, L.ptr(dn.scale), L.ptr(dn.bias), float(LEAKY), L.ptr(stats),
                                                L.ptr(out), L.stream()))
            return out
        if self.first_layer_mode == 2:
            # weights in the kernel parameter block (constant operands of the FMAs): host copies, made once per assignment
            if conv1.host is None:
                conv1.host = (conv1.kernel.detach().cpu().contiguous(), conv1.bias.detach().cpu().contiguous())
            hk, hb = conv1.host
            L.check(L.lib.m4d_rgb_conv_dn_hostw(L.ptr(images), _pix_stride(images), hk.data_ptr(), hb.data_ptr(), b, h, w,
                                                L.ptr(dn.scale), L.ptr(dn.bias), float(LEAKY), L.ptr(stats), L.ptr(out), L.stream()))
            return out
        L.check(L.lib.m4d_rgb_conv_dn(L.ptr(images), _pix_stride(images), L.ptr(conv1.kernel), L.ptr(conv1.bias), b, h, w,
                                      L.ptr(dn.scale), L.ptr(dn.bias), float(LEAKY), L.ptr(stats), L.ptr(out), L.stream()))
        return out

    def call(self, images, on_level=None):
        """``on_level(i, feature_map)`` (optional) is called as soon as level i's feature map has been enqueued: M4Depth uses it
        to start the decoder's encoder-only work of that level on a side stream while the deeper encoder levels run."""
        L.f32c(images, "images")
        prev_out = images
        out_features = []
        for i, (conv1, conv2) in enumerate(zip(self.conv_layers_s1, self.conv_layers_s2)):
            if self.use_dinl and i == 0 and prev_out.shape[-1] == 3 and conv1.filters == 16 and not self.unfused_first_layer:
                tmp = self._first_layer_fused(conv1, prev_out)              # conv + DN + leaky_relu, conv output never stored
            elif self.use_dinl and i == 0:
                tmp = conv1(prev_out, alpha=1.0)
                tmp = self.dn_layers[0].call(tmp, leaky_alpha=LEAKY)        # DN then leaky_relu (:83-84)
            else:
                tmp = conv1(prev_out, alpha=LEAKY)
            prev_out = conv2(tmp, alpha=LEAKY)
            out_features.append(prev_out)
            if on_level is not None:
                on_level(i, prev_out)
        return out_features

    __call__ = call


class DispRefiner:
    """Parallax refiner (m4depth_network.py:93-135): 7 convs in -> 128,128,96 | 64,32,16,5."""

    def __init__(self, regularizer_weight=0.0004):
        self.prep_conv_layers = [_Conv2D(n, 1) for n in (128, 128, 96)]
        self.est_d_conv_layers = [_Conv2D(n, 1) for n in (64, 32, 16, 5)]

    def call(self, feature_map):
        prev_out = feature_map
        for conv in self.prep_conv_layers:
            prev_out = conv(prev_out, alpha=LEAKY)
        prep = prev_out
        n = len(self.est_d_conv_layers)
        for i, conv in enumerate(self.est_d_conv_layers):
            prev_out = conv(prev_out, alpha=LEAKY if i < n - 1 else 1.0)
        # the reference returns [estimate, untouched 96-channel tensor] (zip quirk, :125-135)
        return [prev_out, prep]

    __call__ = call


class DepthEstimatorLevel:
    """One decoder level (m4depth_network.py:138-262) with its recurrent state (prev_f_maps, depth_prev_t)."""

    def __init__(self, settings, depth, regularizer_weight=0.0004):
        self.is_training = settings["is_training"]
        if self.is_training:
            raise NotImplementedError("m4depth_b200 implements the inference path only (is_training=False)")
        self.ablation = settings["ablation"]
        self.disp_refiner = DispRefiner(regularizer_weight=regularizer_weight)
        self.lvl_depth = depth
        self.lvl_mul = depth - 3
        self.interp = L.INTERP_GATHER
        self.shape = None
        self.trace = None           # optional dict: tests set it to {} to receive clones of intermediates
        self._prepared = None       # set by prepare(): {"done": event, "sncv": bool}, consumed by the next call
        self._prep_event = None
        self.pscv_events = None     # optional list: receives (start, end) CUDA events around the PSCV launch

    # ---- state variables, as in the reference (:160-163)
    @property
    def prev_f_maps(self):
        return self._f[1 - self._parity] if self.shape else None

    @property
    def depth_prev_t(self):
        return self._state_depth if self.shape else None

    def build(self, input_shape, device):
        b, h, w, c = input_shape
        self.shape = tuple(input_shape)
        ab = self.ablation
        self.nbre_cuts = 2 ** (self.lvl_depth // 2) if ab.subdivide_features else 1
        cuts = self.nbre_cuts
        # refiner-input channel layout (m4depth_network.py:223-242)
        ch = 0
        self.ch_cv = ch; ch += 9 * cuts
        self.ch_logpara = ch; ch += 1
        self.ch_other = ch if ab.level_memory else -1
        ch += 4 if ab.level_memory else 0
        self.ch_sncv = ch if ab.SNCV else -1
        ch += 49 * cuts if ab.SNCV else 0
        self.ch_logprev = ch if ab.time_recurr else -1
        ch += 1 if ab.time_recurr else 0
        self.cin = ch
        self.xs = (ch + 3) // 4 * 4
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=device)
        self._f = [torch.zeros(b, h, w, c, dtype=torch.float32, device=device) for _ in range(2)]
        self._parity = 0            # _f[_parity] receives the current frame, _f[1-_parity] holds the state
        self._state_depth = torch.ones(b, h, w, 1, dtype=torch.float32, device=device)
        self._para_prev_l, self._depth_prev_l, self._para_prev_t = e(b, h, w, 1), e(b, h, w, 1), e(b, h, w, 1)
        self._other_prev_l = e(b, h, w, 4)
        self._x_in = torch.zeros(b, h, w, self.xs, dtype=torch.float32, device=device)
        self._para, self._depth, self._other = e(b, h, w, 1), e(b, h, w, 1), e(b, h, w, 4)
        self._prev_norm = None

    def prepare(self, curr_f_maps, with_sncv):
        """The part of ``call`` that depends on the encoder output only (:172-189 feature preparation and, when
        ``with_sncv``, the :232 SNCV into the refiner-input buffer), enqueued on the CURRENT stream; the next ``call`` of
        this level waits for it.  DepthEstimatorPyramid runs it for every level on a side stream, so that this
        bandwidth-bound work overlaps the coarse levels, whose small launches leave most of the SMs idle."""
        L.f32c(curr_f_maps, "curr_f_maps")
        if self.shape is None:
            self.build(curr_f_maps.shape, curr_f_maps.device)
        if tuple(curr_f_maps.shape) != self.shape:
            raise L.M4DError(f"level {self.lvl_depth} was built for {self.shape}, got {tuple(curr_f_maps.shape)} "
                             "(static shapes, like the reference's state variables)")
        b, h, w, c = self.shape
        cuts, st = self.nbre_cuts, L.stream()
        cur = self._f[self._parity]
        if self.ablation.normalize_features:
            L.check(L.lib.m4d_group_l2norm(L.ptr(curr_f_maps), b * h * w, c, cuts, L.ptr(cur), st))
        else:
            cur.copy_(curr_f_maps)
        sncv = bool(with_sncv and self.ch_sncv >= 0)
        if sncv:
            with _nvtx("cost_volume"):
                L.check(L.lib.m4d_sncv_fwd(L.ptr(cur), L.ptr(cur), b, h, w, c, cuts, 3,
                                           self._x_in.data_ptr() + 4 * self.ch_sncv, self.xs, st))
        if self._prep_event is None:
            self._prep_event = torch.cuda.Event()
        self._prep_event.record()
        self._prepared = {"done": self._prep_event, "sncv": sncv}

    def call(self, curr_f_maps, prev_l_est, rot, trans, camera, new_traj, prev_f_maps=None, prev_t_depth=None):
        L.f32c(curr_f_maps, "curr_f_maps")
        if self.shape is None:
            self.build(curr_f_maps.shape, curr_f_maps.device)
        if tuple(curr_f_maps.shape) != self.shape:
            raise L.M4DError(f"level {self.lvl_depth} was built for {self.shape}, got {tuple(curr_f_maps.shape)} "
                             "(static shapes, like the reference's state variables)")
        b, h, w, c = self.shape
        cuts, st = self.nbre_cuts, L.stream()
        rot, trans = L.f32c(rot, "rot"), L.f32c(trans, "trans")
        cam_f, cam_c = L.f32c(camera["f"], "camera['f']"), L.f32c(camera["c"], "camera['c']")
        rd = rot.shape[1]
        cur = self._f[self._parity]
        prepared, self._prepared = self._prepared, None
        if prepared is not None:
            # prepare() ran on the pyramid's side stream: everything below that reads `cur` or the SNCV channels waits for it
            torch.cuda.current_stream().wait_event(prepared["done"])

        # :172-189 feature preparation
        if prepared is not None:
            pass
        elif self.ablation.normalize_features:
            L.check(L.lib.m4d_group_l2norm(L.ptr(curr_f_maps), b * h * w, c, cuts, L.ptr(cur), st))
        else:
            cur.copy_(curr_f_maps)
        if self.ablation.normalize_features:
            if prev_f_maps is not None:
                if self._prev_norm is None:
                    self._prev_norm = torch.empty_like(cur)
                L.check(L.lib.m4d_group_l2norm(L.ptr(L.f32c(prev_f_maps, "prev_f_maps")), b * h * w, c, cuts,
                                               L.ptr(self._prev_norm), st))
                prev_f_maps = self._prev_norm

        # :191-194 recurrent state
        if prev_f_maps is None and prev_t_depth is None:
            prev_t_depth = self._state_depth
            prev_f_maps = self._f[1 - self._parity]
        is_new = prev_t_depth is None or _new_traj_flag(new_traj)

        # :196-204 prologue (+ :218 prev_d2para, :224/:227 refiner-input channels)
        if prev_l_est is None:
            po = pp = pdp = None
            ih = iw = 0
        else:
            po, pp, pdp = (L.f32c(prev_l_est[k], k) for k in ("other", "parallax", "depth"))
            ih, iw = pp.shape[1:3]
        scale = 2.0 ** self.lvl_mul
        L.check(L.lib.m4d_level_prologue(
            L.ptr(po), L.ptr(pp), L.ptr(pdp), ih, iw,
            None if is_new else L.ptr(L.f32c(prev_t_depth, "prev_t_depth")),
            L.ptr(rot), rd, L.ptr(trans), L.ptr(cam_f), L.ptr(cam_c), b, h, w,
            L.ptr(self._para_prev_l), L.ptr(self._depth_prev_l), L.ptr(self._other_prev_l),
            None if is_new else L.ptr(self._para_prev_t),
            None if is_new else L.ptr(self._x_in), self.xs, self.ch_logpara, self.ch_other, scale, st))

        if is_new:                                                     # :208-214
            L.check(L.lib.m4d_fill(L.ptr(self._state_depth), b * h * w, 1000.0, st))
            self._parity ^= 1
            return {"depth": self._depth_prev_l, "parallax": self._para_prev_l, "other": self._other_prev_l}

        # :220-221 fused backproject + PSCV -> cv channels and log(prev_disp centre) (:238)
        want_prev = self.ch_logprev >= 0
        if self.pscv_events is not None:               # bench.py: CUDA events around this launch (in-situ roofline)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        torch.cuda.nvtx.range_push("get_parallax_sweeping_cv")          # reference scope: m4depth_network.py:216-221
        L.check(L.lib.m4d_pscv_fused_fwd_ex(
            L.ptr(cur), L.ptr(prev_f_maps), L.ptr(self._para_prev_t) if want_prev else None, L.ptr(self._para_prev_l),
            L.ptr(rot), rd, L.ptr(trans), L.ptr(cam_f), L.ptr(cam_c), b, h, w, c, cuts, 4,
            self._x_in.data_ptr() + 4 * self.ch_cv, self.xs, None, 0,
            (self._x_in.data_ptr() + 4 * self.ch_logprev) if want_prev else None, self.xs, scale, None,
            self.interp, st))
        torch.cuda.nvtx.range_pop()
        if self.pscv_events is not None:
            ev1.record()
            self.pscv_events.append((ev0, ev1))
        # :232 SNCV
        if self.ch_sncv >= 0 and not (prepared is not None and prepared["sncv"]):
            with _nvtx("cost_volume"):                              # m4depth_network.py:232
                L.check(L.lib.m4d_sncv_fwd(L.ptr(cur), L.ptr(cur), b, h, w, c, cuts, 3,
                                           self._x_in.data_ptr() + 4 * self.ch_sncv, self.xs, st))
        f_input = self._x_in[..., :self.cin]
        # :245 refiner
        with _nvtx("DispRefiner"):                                  # m4depth_network.py:244-245
            prev_out = self.disp_refiner(f_input)
        r = prev_out[0]
        # :247-260 epilogue + state update
        L.check(L.lib.m4d_level_epilogue(L.ptr(r), _pix_stride(r), L.ptr(rot), rd, L.ptr(trans), L.ptr(cam_f), L.ptr(cam_c),
                                         b, h, w, 1.0 / scale, L.ptr(self._para), L.ptr(self._depth), L.ptr(self._other),
                                         L.ptr(self._state_depth), st))
        if self.trace is not None:
            self.trace.update(curr_f_maps=cur.clone(), para_prev_t=self._para_prev_t.clone(),
                              para_prev_l=self._para_prev_l.clone(), f_input=f_input.clone(), refiner_out=r.clone())
        self._parity ^= 1
        return {"other": self._other, "depth": self._depth, "parallax": self._para}

    __call__ = call


class DepthEstimatorPyramid:
    """Decoder (m4depth_network.py:265-323), inference mode: temporal state lives in the levels."""

    def __init__(self, settings, regularizer_weight=0.0004, trainable=True):
        self.levels = [DepthEstimatorLevel(settings, i + 1, regularizer_weight=regularizer_weight)
                       for i in range(settings["nbre_lvls"])]
        self.is_training = settings["is_training"]
        self._cam = None
        # M4D_SIDE_STREAM: 0 = everything on the calling stream; 1 = fork after the encoder (one side stream); 2 (default) = the
        # preparation of level l forks as soon as the ENCODER has produced level l (M4Depth._forward), one side stream per level
        self.side_stream_prep = int(os.environ.get("M4D_SIDE_STREAM", "2"))
        self._side = None
        self._fork = None
        self._lvl_side = None

    def fork_prepare(self, l, f_map, new_traj):
        """Enqueue level l's encoder-only work (DepthEstimatorLevel.prepare) on that level's side stream, behind everything
        enqueued on the current stream so far.  The level's next ``call`` waits for it."""
        dev = f_map.device
        if self._lvl_side is None or self._lvl_side[0][0].device != dev:
            self._lvl_side = [(torch.cuda.Stream(device=dev), torch.cuda.Event()) for _ in self.levels]
        side, fork = self._lvl_side[l]
        fork.record(torch.cuda.current_stream())
        side.wait_event(fork)
        with torch.cuda.stream(side):
            self.levels[l].prepare(f_map, with_sncv=not _new_traj_flag(new_traj))

    def call(self, f_maps_pyrs, traj_samples, camera, training=False):
        if training:
            raise NotImplementedError("m4depth_b200 implements the inference path only")
        nl = len(self.levels)
        cam_f, cam_c = L.f32c(camera["f"], "camera['f']"), L.f32c(camera["c"], "camera['c']")
        b = cam_f.shape[0]
        if self._cam is None or self._cam[0].shape[1] != b:
            self._cam = (torch.empty(nl, b, 2, dtype=torch.float32, device=cam_f.device),
                         torch.empty(nl, b, 2, dtype=torch.float32, device=cam_f.device))
        # local_camera["f"|"c"] /= 2**cnter for every level at once (:300-302)
        L.check(L.lib.m4d_camera_pyramid(L.ptr(cam_f), L.ptr(cam_c), b, nl, L.ptr(self._cam[0]), L.ptr(self._cam[1]), L.stream()))
        d_est_seq = []
        for f_pyr_curr, sample in zip(f_maps_pyrs, traj_samples):
            rot, trans, new_traj = sample['rot'], sample['trans'], sample["new_traj"]
            d_est_curr = None
            if self.side_stream_prep and all(lvl._prepared is None for lvl in self.levels):
                # Feature preparation and SNCV of every level depend on the encoder output only: fork them onto a side stream,
                # coarse level first (the order the decoder needs them in).  Levels 6-3 are chains of small launches that leave
                # most SMs idle; the bandwidth-bound preparation of levels 1-2 runs underneath them.  Each level's call waits
                # for its own event, so the result is the same computation (same kernels, same buffers).  Works eagerly and
                # under CUDA-graph capture (the fork / join events become graph dependencies).
                is_new = _new_traj_flag(new_traj)
                main = torch.cuda.current_stream()
                if self._side is None or self._side.device != f_pyr_curr[0].device:
                    self._side = torch.cuda.Stream(device=f_pyr_curr[0].device)
                    self._fork = torch.cuda.Event()
                self._fork.record(main)
                self._side.wait_event(self._fork)
                with torch.cuda.stream(self._side):
                    for l in range(nl - 1, -1, -1):
                        self.levels[l].prepare(f_pyr_curr[l], with_sncv=not is_new)
            for l in range(nl - 1, -1, -1):                                 # coarse -> fine (:293)
                level = self.levels[l]
                local_camera = {"f": self._cam[0][l], "c": self._cam[1][l]}
                d_est = dict(d_est_curr[-1]) if d_est_curr else None
                # a level also waits for the side-stream work of the next finer level (the largest of all: level 1's SNCV), so
                # that the bandwidth-bound kernels of levels 1-2 on this stream do not share the machine with it
                if l >= 1 and self.levels[l - 1]._prepared is not None:
                    torch.cuda.current_stream().wait_event(self.levels[l - 1]._prepared["done"])
                with _nvtx(f"DepthEstimatorLevel/{l + 1}"):
                    est = level(f_pyr_curr[l], d_est, rot, trans, local_camera, new_traj)
                d_est_curr = [est] if d_est_curr is None else d_est_curr + [est]
            d_est_seq.append(d_est_curr[::-1])
        return d_est_seq

    __call__ = call


class M4Depth:
    """m4depth_network.py:325-369: encoder -> decoder -> nearest-neighbour upsampling of the level-1 depth.

    ``use_cuda_graph``: after one eager pass per (new_traj, state parity) variant, single-frame calls are captured
    into CUDA graphs and replayed; inputs are staged into static device buffers first.
    """

    def __init__(self, depth_type="map", nbre_levels=6, is_training=False, ablation_settings=None,
                 use_cuda_graph=True, device=None):
        if is_training:
            raise NotImplementedError("m4depth_b200 implements the inference path only (is_training=False)")
        if not torch.cuda.is_available():
            raise L.M4DError("m4depth_b200 needs a CUDA device (sm_100a): there is no CPU path")
        self.ablation_settings = ablation_settings if ablation_settings is not None else M4depthAblationParameters()
        self.model_settings = {"nbre_lvls": nbre_levels, "is_training": is_training, "ablation": self.ablation_settings}
        self.depth_type = depth_type
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.encoder = FeaturePyramid(self.model_settings, regularizer_weight=0.)
        self.d_estimator = DepthEstimatorPyramid(self.model_settings, regularizer_weight=0.)
        self.step_counter = 0
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        self._seen = set()
        self._static = None
        self._out = None
        self._h2d = None            # side-stream upload state for host (CPU) frames
        self._inputs_event = None   # see inputs_consumed()
        self._fetch = None          # see fetch_depth()

    # ------------------------------------------------------------------------------------------ weights
    def weight_layers(self):
        """{checkpoint key prefix: layer} in the reference's object-graph naming."""
        d = {}
        for i, (c1, c2) in enumerate(zip(self.encoder.conv_layers_s1, self.encoder.conv_layers_s2)):
            d[f"encoder/conv_layers_s1/{i}"] = c1
            d[f"encoder/conv_layers_s2/{i}"] = c2
        for i, lvl in enumerate(self.d_estimator.levels):
            p = f"d_estimator/levels/{i}/disp_refiner"
            for j, cv in enumerate(lvl.disp_refiner.prep_conv_layers):
                d[f"{p}/prep_conv_layers/{j}"] = cv
            for j, cv in enumerate(lvl.disp_refiner.est_d_conv_layers):
                d[f"{p}/est_d_conv_layers/{j}"] = cv
        return d

    def load_weights(self, weights):
        """weights: {"encoder/conv_layers_s1/0/kernel": tensor[3,3,cin,cout], ".../bias": tensor[cout], ...,
        "encoder/dn_layers/0/scale"|"bias": tensor[1,1,1,16]} (torch tensors or numpy arrays).  A checkpoint trained without
        DINL has no dn variables (Keras builds DomainNormalization on its first call only): they are required only when this
        model uses DINL.  Kernel shapes are checked against the layout this model's ablation settings imply."""
        as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
        layers = self.weight_layers()
        need = [k + sfx for k in layers for sfx in ("/kernel", "/bias")]
        if self.encoder.use_dinl:
            need += ["encoder/dn_layers/0/scale", "encoder/dn_layers/0/bias"]
        missing = [k for k in need if k not in weights]
        if missing:
            raise L.M4DError(f"load_weights: {len(missing)} variables missing from the checkpoint, e.g. {missing[:4]}")
        expect = self._expected_cin()
        for prefix, layer in layers.items():
            k = as_t(weights[prefix + "/kernel"])
            if k.dim() == 4 and prefix in expect and k.shape[2] != expect[prefix]:
                raise L.M4DError(f"load_weights: {prefix}/kernel has {k.shape[2]} input channels, this model's layout "
                                 f"(ablation {tuple(self.ablation_settings)}) needs {expect[prefix]}")
            layer.assign(k, as_t(weights[prefix + "/bias"]), self.device)
        dn = self.encoder.dn_layers[0]
        if self.encoder.use_dinl:
            dn.scale = as_t(weights["encoder/dn_layers/0/scale"]).to(self.device, torch.float32).reshape(1, 1, 1, -1).contiguous()
            dn.bias = as_t(weights["encoder/dn_layers/0/bias"]).to(self.device, torch.float32).reshape(1, 1, 1, -1).contiguous()
        self._graphs.clear()
        self._seen.clear()

    def _expected_cin(self):
        """{layer prefix: input channels} of this model's layout (m4depth_network.py:59,102,109,223-242)."""
        ab, d = self.ablation_settings, {}
        cin = 3
        for i, n in enumerate(self.encoder.out_sizes):
            d[f"encoder/conv_layers_s1/{i}"] = cin
            d[f"encoder/conv_layers_s2/{i}"] = n
            cin = n
        for i, _ in enumerate(self.d_estimator.levels):
            cuts = 2 ** ((i + 1) // 2) if ab.subdivide_features else 1
            cin = 9 * cuts + 1 + (4 if ab.level_memory else 0) + (49 * cuts if ab.SNCV else 0) + (1 if ab.time_recurr else 0)
            pfx = f"d_estimator/levels/{i}/disp_refiner"
            for j, n in enumerate((128, 128, 96)):
                d[f"{pfx}/prep_conv_layers/{j}"] = cin
                cin = n
            for j, n in enumerate((64, 32, 16, 5)):
                d[f"{pfx}/est_d_conv_layers/{j}"] = cin
                cin = n
        return d

    def load_checkpoint(self, source, which=None):
        """Load a reference checkpoint: a TF tensor-bundle prefix (``.../cp-0071.ckpt``, what callbacks.py:119-129 saves)
        or an archive such as the reference's ``pretrained_weights.zip`` with ``which`` in ("midair", "kitti").
        Parsed without TensorFlow (m4depth_b200/checkpoint.py)."""
        from .checkpoint import load_reference_weights
        self.load_weights(load_reference_weights(source, which))

    def set_interp(self, mode):
        """Bilinear convention of the PSCV warp for every level (include/m4d.h M4D_INTERP_*)."""
        for lvl in self.d_estimator.levels:
            lvl.interp = mode
        self._graphs.clear()
        self._seen.clear()

    # --------------------------------------------------------------------------------------------- call
    def _forward(self, traj_samples, camera):
        # One frame at a time (the reference encodes all frames first, :358-359): the encoder returns views into per-layer
        # workspaces that the next call overwrites, so a frame's pyramid must be consumed before the next frame is encoded.
        # The recurrent state lives in the levels, so this is the same computation.  Only the last frame's maps stay valid.
        d_maps_pyrs = []
        for s in traj_samples:
            hook = None
            if self.d_estimator.side_stream_prep == 2:
                hook = lambda i, fmap, s=s: self.d_estimator.fork_prepare(i, fmap, s["new_traj"])
            with _nvtx("M4Depth/encoder"):
                pyr = self.encoder(s['RGB_im'], on_level=hook)
            with _nvtx("M4Depth/d_estimator"):
                d_maps_pyrs += self.d_estimator([pyr], [s], camera, False)
        h, w = traj_samples[-1]['RGB_im'].shape[1:3]
        d1 = d_maps_pyrs[-1][0]["depth"]
        b, ih, iw, _ = d1.shape
        if self._out is None or tuple(self._out.shape) != (b, h, w, 1):
            self._out = torch.empty(b, h, w, 1, dtype=torch.float32, device=d1.device)
        L.check(L.lib.m4d_resize_nearest(L.ptr(d1), b, ih, iw, 1, h, w, L.ptr(self._out), L.stream()))    # :368-369
        return d_maps_pyrs

    def _parity(self):
        return self.d_estimator.levels[0]._parity if self.d_estimator.levels[0].shape else 0

    def fetch_depth(self, host_out):
        """Asynchronous read-back of the latest depth map into a pinned host tensor ``[b,H,W,1]``; returns the CUDA event that
        marks the copy done.  The map is first copied into a device staging buffer on the calling stream (a 16 MB device copy:
        microseconds), and the PCIe transfer runs from there on a side stream - so the next ``call`` (which overwrites the
        output buffer) does not wait for the transfer, only the next ``fetch_depth`` does."""
        with torch.cuda.device(self.device):
            if self._out is None:
                raise L.M4DError("fetch_depth: no frame has been processed yet")
            if self._fetch is None or tuple(self._fetch["stage"].shape) != tuple(self._out.shape):
                self._fetch = {"stage": torch.empty_like(self._out), "stream": torch.cuda.Stream(device=self.device),
                               "done": torch.cuda.Event(), "ready": torch.cuda.Event()}
                self._fetch["done"].record()
            f = self._fetch
            main = torch.cuda.current_stream()
            main.wait_event(f["done"])                       # the previous transfer has left the staging buffer
            f["stage"].copy_(self._out, non_blocking=True)
            f["ready"].record()
            f["stream"].wait_event(f["ready"])
            with torch.cuda.stream(f["stream"]):
                host_out.copy_(f["stage"], non_blocking=True)
                f["done"].record()
            return f["done"]

    def inputs_consumed(self):
        """CUDA event recorded after the last asynchronous read of the caller's input tensors by the most recent ``call``.
        With pinned HOST inputs ``call`` returns before the uploads have run: wait for this event (``.synchronize()`` or
        ``stream.wait_event``) before refilling the same host buffers ("Lifetime of input buffers")."""
        return self._inputs_event

    def call(self, data, training=False):
        if training:
            raise NotImplementedError("m4depth_b200 implements the inference path only (training=False)")
        with torch.cuda.device(self.device):            # buffers live on self.device: launch on that device's current stream
            out = self._call(data)
            if self._inputs_event is None:
                self._inputs_event = torch.cuda.Event()
                self._inputs_event.record()
            return out

    def _call(self, data):
        traj_samples, camera = data[0], data[1]
        self.step_counter += 1
        if not (self.use_cuda_graph and len(traj_samples) == 1):
            # eager path (several frames per call, or graphs off): host tensors (e.g. dataloader samples) are uploaded here
            dev = lambda t: t.to(self.device, non_blocking=True) if isinstance(t, torch.Tensor) and t.device.type == "cpu" else t
            traj_samples = [{k: (v if k == "new_traj" else dev(v)) for k, v in s.items()} for s in traj_samples]
            camera = {k: dev(v) for k, v in camera.items()}
            self._forward(traj_samples, camera)
            if self._inputs_event is None:
                self._inputs_event = torch.cuda.Event()
            self._inputs_event.record()
            return {"depth": self._out}

        s = traj_samples[0]
        nt = _new_traj_flag(s["new_traj"])
        rgb = s['RGB_im']
        if self._static is None or tuple(self._static["RGB_im"].shape) != tuple(rgb.shape):
            dev = self.device
            b = rgb.shape[0]
            self._static = {"RGB_im": torch.empty(tuple(rgb.shape), dtype=torch.float32, device=dev),
                            "rot": torch.empty(tuple(s['rot'].shape), dtype=torch.float32, device=dev),
                            "trans": torch.empty(b, 3, dtype=torch.float32, device=dev),
                            "f": torch.empty(b, 2, dtype=torch.float32, device=dev),
                            "c": torch.empty(b, 2, dtype=torch.float32, device=dev)}
            self._graphs.clear()
            self._seen.clear()
        st = self._static
        if rgb.device.type == "cpu":
            # Host frame: upload on a side stream into one of two staging buffers so that the copy of frame t+1 overlaps the
            # graph of frame t (the host runs ahead of the device), then a device-to-device copy into the graph's static input.
            if self._h2d is None or tuple(self._h2d["buf"][0].shape) != tuple(rgb.shape):
                self._h2d = {"stream": torch.cuda.Stream(device=self.device), "idx": 0,
                             "buf": [torch.empty_like(st["RGB_im"]) for _ in range(2)],
                             "up": [torch.cuda.Event() for _ in range(2)], "used": [torch.cuda.Event() for _ in range(2)]}
                for e in self._h2d["used"]:
                    e.record()
            hd = self._h2d
            i = hd["idx"] = 1 - hd["idx"]
            main = torch.cuda.current_stream()
            hd["stream"].wait_event(hd["used"][i])          # the graph input copy that last read this staging buffer
            with torch.cuda.stream(hd["stream"]):
                hd["buf"][i].copy_(rgb, non_blocking=True)
                hd["up"][i].record()
            main.wait_event(hd["up"][i])
            st["RGB_im"].copy_(hd["buf"][i], non_blocking=True)
            hd["used"][i].record()
        else:
            st["RGB_im"].copy_(rgb, non_blocking=True)
        st["rot"].copy_(s['rot'], non_blocking=True)
        st["trans"].copy_(s['trans'], non_blocking=True)
        st["f"].copy_(camera["f"], non_blocking=True)
        st["c"].copy_(camera["c"], non_blocking=True)
        if self._inputs_event is None:
            self._inputs_event = torch.cuda.Event()
        if rgb.device.type == "cpu":
            torch.cuda.current_stream().wait_event(self._h2d["up"][self._h2d["idx"]])
        self._inputs_event.record()                      # every read of the caller's tensors is enqueued before this point
        sample = {"RGB_im": st["RGB_im"], "rot": st["rot"], "trans": st["trans"], "new_traj": [nt]}
        cam = {"f": st["f"], "c": st["c"]}
        key = (nt, self._parity())
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
            for lvl in self.d_estimator.levels:          # host-side mirror of the state ping-pong the graph performs
                lvl._parity ^= 1
        elif key not in self._seen:
            self._seen.add(key)                          # first occurrence: eager (allocations, module loading)
            self._forward([sample], cam)
        else:
            parities = [lvl._parity for lvl in self.d_estimator.levels]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._forward([sample], cam)
            self._graphs[key] = g
            for lvl, p in zip(self.d_estimator.levels, parities):   # capture ran the host code but not the kernels
                lvl._parity = p
            g.replay()
            for lvl in self.d_estimator.levels:
                lvl._parity ^= 1
        return {"depth": self._out}

    __call__ = call

    def predict_step(self, data):
        """m4depth_network.py:476-489 for a single frame dict: {"image", "depth", "new_traj"} like the reference."""
        preds = self.call([[data], data["camera"]], training=False)
        return {"image": data["RGB_im"], "depth": preds["depth"], "new_traj": data["new_traj"]}

    def reset_metrics(self):
        from .metrics import MetricsAccumulator
        self.compiled_metrics = MetricsAccumulator(self.device)

    def test_step(self, data):
        """m4depth_network.py:433-474.  ``data``: one sequence element (``depth`` [b,H,W,1]) or a whole sequence (``depth``
        [b,T,H,W,1], every per-frame entry with the extra axis 1: the KITTI protocol) - then the network runs over the T
        frames and only the LAST frame is scored, unconditionally (:455-457).  Ground truth is clipped to [0, 80], the estimate
        to [0.001, 80] (:465-467); a single frame that starts a trajectory (``new_traj``) is not scored (:469).  Returns the
        running means of the seven metrics (metrics.py), like Keras' ``{m.name: m.result()}``."""
        if not hasattr(self, "compiled_metrics"):
            self.reset_metrics()
        if data["depth"].dim() == 5:
            seq_len = data["depth"].shape[1]
            nt_all = data["new_traj"]
            traj_samples = []
            for i in range(seq_len):
                nt = nt_all[:, i] if hasattr(nt_all, "dim") and nt_all.dim() == 2 else nt_all[i]
                traj_samples.append({"RGB_im": data["RGB_im"][:, i].contiguous(), "rot": data["rot"][:, i].contiguous(),
                                     "trans": data["trans"][:, i].contiguous(), "new_traj": nt})
            preds = self.call([traj_samples, data["camera"]], training=False)
            gt, new_traj = data["depth"][:, -1].contiguous(), False
        else:
            preds = self.call([[data], data["camera"]], training=False)
            gt, nt = data["depth"], data["new_traj"]
            new_traj = _new_traj_flag(nt)
        if not new_traj:
            self.compiled_metrics.update_state(gt.to(self.device), preds["depth"])      # clipping happens in m4d_depth_metrics
        return self.compiled_metrics.result()

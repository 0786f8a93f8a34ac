"""m4depth_b200 - B200-native (sm_100a) implementation of M4Depth's per-frame parallax-inference hot path.

The product is ``libm4d.so`` (hand-written CUDA behind the C ABI in ``include/m4d.h``); this package is the thin
Python host that mirrors the reference's layer / function API on torch CUDA tensors:

    m4depth_b200.m4depth_network   <->  reference m4depth_network.py  (layers, M4Depth)
    m4depth_b200.utils             <->  reference utils/depth_operations.py, utils/dense_image_warp.py
    m4depth_b200.metrics           <->  reference metrics.py

There is no CPU / PyTorch fallback: anything that computes needs the built library (``m4depth_b200._lib`` raises
ImportError without it) and ops fail on non-CUDA tensors.  The library is loaded on first use of a computing name, so the
pure-Python helpers (``m4depth_b200.weights``, ``.checkpoint``, ``.dist``, ``._build``) import without it.
"""
import importlib

_LAZY = {
    "_lib": ("._lib", None), "M4DError": ("._lib", "M4DError"), "INTERP_GATHER": ("._lib", "INTERP_GATHER"),
    "INTERP_BP": ("._lib", "INTERP_BP"), "INTERP_BP_FMA": ("._lib", "INTERP_BP_FMA"), "launch_count": ("._lib", "launch_count"),
    "M4Depth": (".m4depth_network", "M4Depth"), "M4depthAblationParameters": (".m4depth_network", "M4depthAblationParameters"),
    "DomainNormalization": (".m4depth_network", "DomainNormalization"), "FeaturePyramid": (".m4depth_network", "FeaturePyramid"),
    "DispRefiner": (".m4depth_network", "DispRefiner"), "DepthEstimatorLevel": (".m4depth_network", "DepthEstimatorLevel"),
    "DepthEstimatorPyramid": (".m4depth_network", "DepthEstimatorPyramid"), "m4depth_network": (".m4depth_network", None),
    "utils": (".utils", None), "metrics": (".metrics", None),
}

__all__ = ["M4Depth", "M4depthAblationParameters", "DomainNormalization", "FeaturePyramid", "DispRefiner",
           "DepthEstimatorLevel", "DepthEstimatorPyramid", "utils", "metrics", "M4DError", "INTERP_GATHER",
           "INTERP_BP", "INTERP_BP_FMA", "launch_count"]


def __getattr__(name):
    spec = _LAZY.get(name)
    if spec is None:
        raise AttributeError(f"module 'm4depth_b200' has no attribute {name!r}")
    mod = importlib.import_module(spec[0], __name__)
    val = mod if spec[1] is None else getattr(mod, spec[1])
    globals()[name] = val
    return val


def __dir__():
    return sorted(list(globals()) + list(_LAZY))

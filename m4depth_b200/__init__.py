"""m4depth_b200 - B200-native (sm_100a) implementation of M4Depth's per-frame parallax-inference hot path.

The product is ``libm4d.so`` (hand-written CUDA behind the C ABI in ``include/m4d.h``); this package is the thin
Python host that mirrors the reference's layer / function API on torch CUDA tensors:

    m4depth_b200.m4depth_network   <->  reference m4depth_network.py  (layers, M4Depth)
    m4depth_b200.utils             <->  reference utils/depth_operations.py, utils/dense_image_warp.py
    m4depth_b200.metrics           <->  reference metrics.py

There is no CPU / PyTorch fallback: importing fails if the library is not built, ops fail on non-CUDA tensors.
"""
from . import _lib
from ._lib import M4DError, INTERP_GATHER, INTERP_BP, INTERP_BP_FMA, launch_count
from .m4depth_network import (M4Depth, M4depthAblationParameters, DomainNormalization, FeaturePyramid, DispRefiner,
                              DepthEstimatorLevel, DepthEstimatorPyramid)
from . import utils
from . import metrics

__all__ = ["M4Depth", "M4depthAblationParameters", "DomainNormalization", "FeaturePyramid", "DispRefiner",
           "DepthEstimatorLevel", "DepthEstimatorPyramid", "utils", "metrics", "M4DError", "INTERP_GATHER",
           "INTERP_BP", "INTERP_BP_FMA", "launch_count"]

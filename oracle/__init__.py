"""CPU oracle for the M4Depth parallax-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``m4depth_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker / the CPU arm.

What it is: an op-for-op restatement, in torch-CPU fp32 (one rounded fp32 op per reference
TF op, no autograd, no fused multiply-add across ops), of

* ``utils/depth_operations.py:18-68,140-313``   (geometry, PSCV, SNCV)
* ``utils/dense_image_warp.py:61-268``          (bilinear warp, both code paths)
* ``cuda_backproject/backproject_op_gpu.cu.cc:19-79,108-196`` (BackProject forward and gradient)
* ``m4depth_network.py:24-369``                 (DomainNormalization ... M4Depth.call)
* ``metrics.py:1-64``

PARITY PIN.  ``oracle/_ref/libbackproject_ref.so`` (``oracle/ref_binary.py``) is the reference's own BackProject CUDA source
compiled unmodified: the oracle of the BackProject op and of the BP_FMA warp inside the fused PSCV (GPU tests).  For everything
else: the reference ships no tests and no golden vectors, and TensorFlow is not
installed in the build image nor on the GPU box, so the reference cannot be executed as-is.
The pin used instead: ``tools/gen_golden.py`` imports the reference's *own, unmodified*
``utils/depth_operations.py`` / ``utils/dense_image_warp.py`` / ``m4depth_network.py`` from
``/root/reference`` on top of ``tools/tf_shim`` (a small numpy/torch restatement of the ~60 TF
primitives those files call) and records their outputs on seeded inputs into
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this oracle against those
vectors.  The pin therefore covers the reference's Python op sequence exactly; the arithmetic
of the TF primitives themselves (Conv2D, resize_bilinear, fp16 reduce_mean summation order,
batched matmul summation order) is restated from TF 2.7's documented semantics and remains
**unpinned against real TensorFlow** ("Parity pin").
"""

from .geometry import (get_rot_mat, get_coords_2d, parallax2depth, depth2parallax,
                       prev_d2para, pscv_query_points)
from .warp import (interpolate_bilinear, back_project, back_project_grad, dense_image_warp,
                   back_project_index_grids)
from .cost_volumes import get_parallax_sweeping_cv, cost_volume, tile_in_batch
from .network import (conv2d_same, leaky_relu, resize_bilinear_legacy, resize_nearest,
                      group_l2_normalize, DomainNormalization, FeaturePyramid, DispRefiner,
                      DepthEstimatorLevel, DepthEstimatorPyramid, M4Depth,
                      M4depthAblationParameters, init_weights, level_channels, reduction_order)
from .metrics import depth_metrics, METRIC_NAMES

__all__ = [n for n in dir() if not n.startswith("_")]

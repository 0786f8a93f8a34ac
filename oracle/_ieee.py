"""Correctly rounded fp32 square root for the oracle (test infrastructure).

torch.sqrt on CPU goes through MKL VML, which is NOT correctly rounded in fp32 (about 0.7 % of inputs are 1 ulp
off; measured against CUDA's sqrt.rn on the GPU box, tools/gpu_diag2.py).  TensorFlow's CPU kernels (Eigen,
``_mm256_sqrt_ps``) and ``__fsqrt_rn`` on the GPU are IEEE-exact, so the oracle uses numpy's sqrt (hardware
``sqrtps``, exact); tests/test_oracle_golden.py cross-checks it against the float64 route.
"""
import numpy as np
import torch


def sqrt(x):
    a = x.detach().contiguous().numpy()
    return torch.from_numpy(np.sqrt(a)).reshape(x.shape)

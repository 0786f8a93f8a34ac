"""Oracle: the 7 depth metrics (test infrastructure, see oracle/__init__.py).

Restates ``/root/reference/metrics.py:1-64`` and the clipping of ``m4depth_network.py:465-467``.
Each metric is ONE per-batch masked mean (what a single ``update_state`` feeds ``keras.metrics.Mean``);
the run-level value is the mean of these per-batch values.
"""
import torch

METRIC_NAMES = ["AbsRel", "SqRel", "RMSE", "RMSE_log", "Delta1", "Delta2", "Delta3"]   # main.py:124-131 order


def _masked_mean(err, ref):
    mask = (ref > 1e-6).to(torch.float32)                       # metrics.py:4
    e = torch.where(mask != 0, err * mask, torch.zeros_like(err))   # multiply_no_nan
    return e.sum() / torch.clamp(mask.sum(), min=1.0)


def depth_metrics(gt, est, max_d=80.0):
    """gt, est [b,H,W,1] -> tensor[7] in METRIC_NAMES order (clips as test_step does)."""
    gt = torch.clamp(gt, 0.0, max_d)
    est = torch.clamp(est, 0.001, max_d)
    out = []
    out.append(_masked_mean(torch.abs(gt - est) / (gt + 1e-6), gt))
    out.append(_masked_mean((gt - est) ** 2 / (gt + 1e-6), gt))
    out.append(torch.sqrt(_masked_mean((gt - est) ** 2, gt)))
    lg, le = torch.log(gt + 1e-6), torch.log(est + 1e-6)
    out.append(torch.sqrt(_masked_mean((lg - le) ** 2, lg)))   # quirk: mask on log(gt) > 1e-6 (metrics.py:24-28)
    thresh = torch.maximum(gt / est, est / gt)
    for p in (1, 2, 3):
        out.append(_masked_mean((thresh < 1.25 ** p).to(torch.float32), gt))
    return torch.stack(out)

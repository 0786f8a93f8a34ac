"""Mirror of the reference's ``metrics.py`` (7 depth metrics) + the clipping of ``test_step``
(m4depth_network.py:465-467), computed by one libm4d reduction per batch.

``depth_metrics(gt, est)`` returns what ONE ``update_state`` call feeds each ``keras.metrics.Mean``: the masked
per-batch means [AbsRel, SqRel, RMSE, RMSE_log, Delta1, Delta2, Delta3].  ``MetricsAccumulator`` is the run-level
mean of those (and the object whose partial sums are all-gathered across ranks at the end of a run).
"""
import torch

from . import _lib as L

METRIC_NAMES = ["AbsRel", "SqRel", "RMSE", "RMSE_log", "Delta1", "Delta2", "Delta3"]


def depth_metrics(gt, est, max_d=80.0, out=None, ws=None):
    L.f32c(gt, "gt"), L.f32c(est, "est")
    if gt.numel() != est.numel():
        raise L.M4DError("depth_metrics: gt and est must have the same number of elements")
    if ws is None:
        ws = torch.empty(16, dtype=torch.float64, device=gt.device)
    if out is None:
        out = torch.empty(7, dtype=torch.float32, device=gt.device)
    L.check(L.lib.m4d_depth_metrics(L.ptr(gt), L.ptr(est), gt.numel(), float(max_d), L.ptr(ws), L.ptr(out), L.stream()))
    return out


class MetricsAccumulator:
    """keras.metrics.Mean x 7: sum of per-batch values and a count; ``partials()`` is the 14-float record a rank
    contributes to the end-of-run all-gather."""

    def __init__(self, device):
        self.sum = torch.zeros(7, dtype=torch.float64, device=device)
        self.count = 0
        self._ws = torch.empty(16, dtype=torch.float64, device=device)
        self._out = torch.empty(7, dtype=torch.float32, device=device)

    def update_state(self, gt, est, max_d=80.0):
        self.sum += depth_metrics(gt, est, max_d, self._out, self._ws).double()
        self.count += 1

    def partials(self):
        return torch.cat((self.sum, torch.full((7,), float(self.count), dtype=torch.float64, device=self.sum.device)))

    @staticmethod
    def reduce(partials):
        """partials [n_ranks,14] (or [14]) -> dict of run-level means."""
        p = partials.reshape(-1, 14).sum(dim=0)
        vals = (p[:7] / torch.clamp(p[7:], min=1.0)).tolist()
        return dict(zip(METRIC_NAMES, vals))

    def result(self):
        return self.reduce(self.partials())

"""Vector-valued-distance reductions - mirror of sympa/manifolds/metrics.py:6-121.

Inside `manifold.dist` the reduction is fused into the CUDA kernel (selected by `Metric.name`);
`compute_metric(v)` is kept for API compatibility and evaluates the same formula on an already
computed, ascending vector-valued distance with element-wise torch ops."""
from abc import ABC, abstractmethod
from enum import Enum

import torch


class MetricType(Enum):  # metrics.py:6-17
    RIEMANNIAN = "riem"
    FINSLER_ONE = "fone"
    FINSLER_INFINITY = "finf"
    FINSLER_MINIMUM = "fmin"
    WEIGHTED_SUM = "wsum"

    @staticmethod
    def from_str(label):
        return {t.value: t for t in MetricType}[label]


class Metric(ABC):
    name = None

    def __init__(self, dims: int):
        self.dims = dims

    @abstractmethod
    def compute_metric(self, v: torch.Tensor, keepdim=False) -> torch.Tensor:
        raise NotImplementedError

    @classmethod
    def get(cls, type: MetricType, dims: int):  # metrics.py:29-39
        table = {
            MetricType.RIEMANNIAN: RiemannianMetric,
            MetricType.FINSLER_ONE: FinslerOneMetric,
            MetricType.FINSLER_INFINITY: FinslerInfinityMetric,
            MetricType.FINSLER_MINIMUM: FinslerMinimumEntropyMetric,
            MetricType.WEIGHTED_SUM: FinslerWeightedSumMetric,
        }
        return table[type](dims)


class RiemannianMetric(Metric):  # metrics.py:42-52
    name = "riem"

    def compute_metric(self, v, keepdim=False):
        return torch.norm(v, dim=-1, keepdim=keepdim)


class FinslerOneMetric(Metric):  # metrics.py:55-65
    name = "fone"

    def compute_metric(self, v, keepdim=False):
        return torch.sum(v, dim=-1, keepdim=keepdim)


class FinslerInfinityMetric(Metric):  # metrics.py:68-81
    name = "finf"

    def compute_metric(self, v, keepdim=False):
        res = v[:, -1]
        return res.reshape(-1, 1) if keepdim else res


class FinslerMinimumEntropyMetric(Metric):  # metrics.py:84-100
    name = "fmin"

    def __init__(self, dims):
        super().__init__(dims)
        # the code at metrics.py:88-89 evaluates to [0, 2, ..., 2(n-1)] on the ascending vvd
        self.weights = (2 * torch.arange(dims)).unsqueeze(0)

    def compute_metric(self, v, keepdim=False):
        return torch.sum(self.weights.to(v) * v, dim=-1, keepdim=keepdim)


class FinslerWeightedSumMetric(Metric, torch.nn.Module):  # metrics.py:103-121
    name = "wsum"

    def __init__(self, dims):
        torch.nn.Module.__init__(self)
        Metric.__init__(self, dims)
        # float64 like the embedding table and the scale: the reference gets that from the global default dtype it
        # sets in sympa/config.py:17-18; this package does not touch the global default
        self.weights = torch.nn.parameter.Parameter(torch.ones((1, dims), dtype=torch.float64))

    def compute_metric(self, v, keepdim=False):
        return torch.sum(torch.relu(self.weights) * v, dim=-1, keepdim=keepdim)

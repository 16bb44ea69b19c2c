"""Siegel manifolds - drop-in mirror of sympa/manifolds/siegel_manifold.py with `dist` routed to the
CUDA hot path (sympa_b200.ops -> libsympa_b200.so)."""
from abc import ABC
from typing import Optional, Tuple, Union

import torch

from .. import ops
from . import csym as sm
from .base import Manifold
from .metrics import Metric, MetricType


class SiegelManifold(Manifold, ABC):
    """Spaces of complex symmetric matrices stored as real tensors (b, 2, n, n)
    (reference: sympa/manifolds/siegel_manifold.py:11-39)."""

    ndim = 1
    reversible = False
    name = "Siegel Space"
    kind = None  # "upper" | "bounded": selects the kernel entry
    __scaling__ = Manifold.__scaling__.copy()

    def __init__(self, dims=2, ndim=2, metric=MetricType.RIEMANNIAN, use_xitorch=False):
        super().__init__()
        self.dims = dims
        self.ndim = ndim
        self._projected = 0
        self._projected_dev = None
        if isinstance(metric, str):
            metric = MetricType.from_str(metric)
        self.metric = Metric.get(metric, self.dims)

    # ------------------------------------------------------------------ hot path
    def _wsum(self, like):
        if self.metric.name != "wsum":
            return None
        w = self.metric.weights
        if w.device != like.device:
            raise RuntimeError("wsum weights and points are on different devices; call manifold.to(device)")
        return w

    def dist(self, z1: torch.Tensor, z2: torch.Tensor, *, keepdim=False) -> torch.Tensor:
        """Distance between points of shape (b, 2, n, n); returns (b,) - `keepdim` is ignored exactly as
        in the reference (siegel_manifold.py:41,71-72).  One fused CUDA kernel forward, analytic
        backward."""
        d, _ = ops.dist(self.kind, self.metric.name, z1, z2, self._wsum(z1))
        return d

    def vvd(self, z1: torch.Tensor, z2: torch.Tensor) -> torch.Tensor:
        """Ascending vector-valued distance log((1 + d_i) / (1 - d_i)), (b, n).  Addition to the
        reference API (the reference only has it as a local inside dist, siegel_manifold.py:69-70)."""
        with torch.no_grad():
            _, v, _ = ops.forward_raw(self.kind, self.metric.name, z1=z1, z2=z2, wsum_w=self._wsum(z1))
        return v

    def dist_from_table(self, table: torch.Tensor, idx: torch.Tensor, sync_grad: bool = False, accumulator=None) -> torch.Tensor:
        """dist(table[idx[:, 0]], table[idx[:, 1]]) with the gather and its backward fused
        (replaces sympa/embeddings.py:29-34 + model.py:26-38).  sync_grad=True makes the backward average the
        table gradient over the ranks itself (on the packed gradient table): do not all-reduce it again.
        accumulator: a `table_grad_accumulator(table)` shared by the calls of one step (gradient accumulation)."""
        d, _ = ops.table_dist(self.kind, self.metric.name, table, idx, self._wsum(table), sync_grad=sync_grad,
                              accumulator=accumulator)
        return d

    def table_grad_accumulator(self, table: torch.Tensor, scatter_sms=None):
        """see sympa_b200.ops.TableGradAccumulator"""
        return ops.TableGradAccumulator(self.kind, table, scatter_sms)

    def dist_matrix(self, table: torch.Tensor, row_begin: int = 0, row_count=None) -> torch.Tensor:
        """All-pairs distances between the rows of `table` (rows [row_begin, row_begin + row_count) of the
        matrix, zeros on the diagonal, forward only): what Runner.build_distance_matrix
        (sympa/runner.py:142-154) assembles with one dist call per node."""
        return ops.dist_matrix(self.kind, self.metric.name, table, row_begin, row_count, self._wsum(table))

    # ------------------------------------------------------------------ optimizer side
    def retr(self, x: torch.Tensor, u: torch.Tensor) -> torch.Tensor:  # siegel_manifold.py:74-87
        return self.projx(x + u)

    def projx(self, x: torch.Tensor) -> torch.Tensor:  # siegel_manifold.py:130-137
        return sm.to_symmetric(x)

    def proju(self, x: torch.Tensor, u: torch.Tensor) -> torch.Tensor:  # :139-140
        return self.egrad2rgrad(x, u)

    def transp(self, x, y, v):  # :142-154
        return v

    def expmap(self, x, u):
        pass

    def logmap(self, x, y):
        pass

    def _count_projected(self, mask):
        n_proj = (~mask.reshape(-1)).sum()
        self._projected_dev = n_proj if self._projected_dev is None else self._projected_dev + n_proj

    @property
    def projected_points(self):
        """Number of points projx had to move (upper_half.py:64) - accumulated on the device, read
        lazily so that the optimizer step does not synchronise."""
        if self._projected_dev is not None:
            self._projected += int(self._projected_dev.item())
            self._projected_dev = None
        counter = getattr(self, "_projected_counter", None)   # rows moved by the fused optimizer kernel
        if counter is not None:
            self._projected += int(counter.item())
            counter.zero_()
        return self._projected

    @projected_points.setter
    def projected_points(self, value):
        self._projected = value
        self._projected_dev = None

    # ------------------------------------------------------------------ checks
    def _check_shape(self, shape: Tuple[int], name: str) -> Union[Tuple[bool, Optional[str]], bool]:
        ok = shape[-1] == self.dims and shape[-2] == self.dims  # siegel_manifold.py:89-118
        reason = None if ok else "'{}' on the {} requires more than {} dim".format(name, self, self.dims)
        return ok, reason

    def _check_matrices_are_symmetric(self, x: torch.Tensor, *, atol=1e-5, rtol=1e-5):
        return bool(torch.allclose(x, x.transpose(-1, -2), atol=atol, rtol=rtol))

    def _check_vector_on_tangent(self, x, u, *, atol=1e-5, rtol=1e-5):
        pass

    def extra_repr(self):
        return f"dims={self.dims}, metric={self.metric.name}"

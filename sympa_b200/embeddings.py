"""Embedding table of matrix points - mirror of the matrix part of sympa/embeddings.py (the vector
models are out of scope)."""
import torch
import torch.nn as nn

from .manifolds import BoundedDomainManifold, MetricType, SymmetricPositiveDefinite, UpperHalfManifold

INIT_EPS = 1e-3  # sympa/config.py:21

try:  # pragma: no cover - geoopt is not in the build image
    from geoopt import ManifoldParameter  # type: ignore
except Exception:  # noqa: BLE001

    class ManifoldParameter(nn.Parameter):
        """nn.Parameter that remembers its manifold (what geoopt.ManifoldParameter provides to the
        optimizer, used at sympa/embeddings.py:27)."""

        def __new__(cls, data=None, manifold=None, requires_grad=True):
            inst = nn.Parameter.__new__(cls, data, requires_grad)
            inst.manifold = manifold
            return inst


class MatrixEmbeddings(nn.Module):
    """(num_embeddings, 2, n, n) complex symmetric points, or (num_embeddings, n, n) for spd
    (sympa/embeddings.py:54-79)."""

    def __init__(self, num_embeddings, dims, manifold):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.dims = dims
        self.manifold = manifold
        embeds = manifold.random(num_embeddings, dims, dims, from_=-INIT_EPS, to=INIT_EPS, dtype=torch.float64)
        if isinstance(manifold, SymmetricPositiveDefinite):   # embeddings.py:70-72
            embeds = embeds * INIT_EPS + torch.diag_embed(torch.ones(num_embeddings, dims, dtype=torch.float64))
        self.embeds = ManifoldParameter(embeds, manifold=manifold)

    def forward(self, input_index):   # embeddings.py:29-34 (plain gather; the fused path does not call this)
        return self.embeds[input_index]

    def proj_embeds(self):
        with torch.no_grad():
            self.embeds.data = self.manifold.projx(self.embeds.data)

    def check_all_points(self):
        """embeddings.py:41-47 checks the points one by one in a Python loop; the same predicate is evaluated in one
        kernel launch for a table on the GPU (batched torch operations for one on the CPU) and the first offender
        reported."""
        pts = self.embeds.data
        kind = getattr(self.manifold, "kind", None)
        if pts.is_cuda and pts.dtype == torch.float64 and kind in ("upper", "bounded", "spd") and pts.shape[-1] <= 10:
            # one launch over the table (sympa_check_points): same predicate, first offender, reason strings
            from . import ops
            ok, row, reason = ops.check_points(kind, pts)
            return (True, None, None) if ok else (False, pts[row], reason)
        if not torch.allclose(pts, pts.transpose(-1, -2), atol=1e-5, rtol=1e-5):
            for i in range(len(pts)):
                ok, reason = self.manifold.check_point_on_manifold(pts[i], explain=True)
                if not ok:
                    return False, pts[i], reason
        ok, reason = self.manifold.check_point_on_manifold(pts, explain=True)
        if ok:
            return True, None, None
        for i in range(len(pts)):
            ok, reason = self.manifold.check_point_on_manifold(pts[i], explain=True)
            if not ok:
                return False, pts[i], reason
        return True, None, None

    def norm(self):
        return self.embeds.data.reshape(len(self.embeds), -1).norm(dim=-1)


class ManifoldFactory:
    """matrix manifolds of sympa/embeddings.py:131-158"""

    @classmethod
    def get_manifold(cls, manifold_name, metric_name, dims):
        if manifold_name == "spd":
            return SymmetricPositiveDefinite()      # the metric flag is ignored, embeddings.py:153-154
        table = {"upper": UpperHalfManifold, "bounded": BoundedDomainManifold}
        if manifold_name not in table:
            raise ValueError(f"sympa_b200 implements upper / bounded / spd, not {manifold_name!r}")
        return table[manifold_name](dims=dims, metric=MetricType.from_str(metric_name))

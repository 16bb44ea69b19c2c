"""Manifold base class: geoopt's when geoopt is installed (so geoopt optimizers accept these
manifolds unchanged), otherwise the small subset of it that sympa uses
(reference: sympa/manifolds/siegel_manifold.py:4,11-22; sympa/embeddings.py:44)."""
import torch

try:  # pragma: no cover - geoopt is not in the build image
    from geoopt.manifolds.base import Manifold  # type: ignore
except Exception:  # noqa: BLE001

    class Manifold(torch.nn.Module):
        __scaling__ = {}
        name = ""
        ndim = 0
        reversible = False

        def __init__(self, **kwargs):
            super().__init__()

        def _check_shape(self, shape, name):
            return True, None

        def check_point_on_manifold(self, x, *, explain=False, atol=1e-5, rtol=1e-5):
            ok, reason = self._check_shape(x.shape, "x")
            if ok:
                ok, reason = self._check_point_on_manifold(x, atol=atol, rtol=rtol)
            return (ok, reason) if explain else ok

        def extra_repr(self):
            return ""

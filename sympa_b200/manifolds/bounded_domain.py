"""Bounded domain D_n = {Z in Sym(n, C) : I - conj(Z) Z > 0} - mirror of
sympa/manifolds/bounded_domain.py."""
import torch

from . import csym as sm
from .base import Manifold
from .metrics import MetricType
from .siegel_manifold import SiegelManifold
from .upper_half import UpperHalfManifold


class BoundedDomainManifold(SiegelManifold):
    ndim = 1
    reversible = False
    name = "Bounded Domain"
    kind = "bounded"
    __scaling__ = Manifold.__scaling__.copy()

    def __init__(self, dims=2, ndim=2, metric=MetricType.RIEMANNIAN):
        super().__init__(dims=dims, ndim=ndim, metric=metric)

    # dist: inherited - the kernel applies the inverse Cayley transform to both operands itself
    # (bounded_domain.py:27-39)

    def egrad2rgrad(self, z: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
        """A u A with A = I - conj(Z) Z   (bounded_domain.py:41-53, :163-170)."""
        zc = sm.to_complex(z)
        a = torch.eye(zc.shape[-1], dtype=zc.dtype, device=zc.device) - zc.conj() @ zc
        return sm.from_complex(a @ sm.to_complex(u) @ a)

    def projx(self, z: torch.Tensor) -> torch.Tensor:
        """Clamp the Takagi values at 1 - eps for the points that need it (bounded_domain.py:55-84).
        The reference as shipped crashes here (it builds its factoriser without eigenvectors,
        siegel_manifold.py:38 vs bounded_domain.py:70); this implements what its tests describe
        (tests/test_bounded_domain.py:17-66)."""
        z = super().projx(z)
        eps = sm.EPS[z.dtype]
        d, s = sm.takagi(z)
        ok = torch.all(d < 1 - eps, dim=-1)
        dt = torch.diag_embed(d.clamp(max=1 - eps)).to(s.dtype)
        # conj(S) D S^H with S^H = conj(S)^T
        z_t = sm.from_complex(s.conj() @ dt @ s.transpose(-1, -2).conj())
        self._count_projected(ok)
        return torch.where(ok.reshape(-1, 1, 1, 1), z, z_t)

    def inner(self, z, u, v=None, *, keepdim=False):
        """tr[(I - conj(Z) Z)^-1 u (I - Z conj(Z))^-1 conj(v)]   (bounded_domain.py:86-117)."""
        if v is None:
            v = u
        zc = sm.to_complex(z)
        eye = torch.eye(zc.shape[-1], dtype=zc.dtype, device=zc.device)
        left = torch.linalg.inv(eye - zc.conj() @ zc)
        right = torch.linalg.inv(eye - zc @ zc.conj())
        res = left @ sm.to_complex(u) @ right @ sm.to_complex(v).conj()
        tr = torch.diagonal(res.real, dim1=-2, dim2=-1).sum(-1).reshape(-1, 1, 1)
        return sm.stick(tr, tr)

    def _check_point_on_manifold(self, x: torch.Tensor, *, atol=1e-5, rtol=1e-5):
        if not self._check_matrices_are_symmetric(x, atol=atol, rtol=rtol):  # bounded_domain.py:119-150
            return False, "Matrices are not symmetric"
        zc = torch.complex(x[..., 0, :, :], x[..., 1, :, :])
        a = torch.eye(zc.shape[-1], dtype=zc.dtype, device=zc.device) - zc.conj() @ zc
        ok = bool(torch.allclose(a, a.transpose(-1, -2).conj()))
        return ok, (None if ok else "'Id - conj(Z) Z' is not hermitian (is not definite positive)")

    def random(self, *size, dtype=None, device=None, **kwargs) -> torch.Tensor:
        """Cayley transform of upper-half random points (bounded_domain.py:152-160)."""
        pts = UpperHalfManifold(dims=self.dims).random(*size, dtype=torch.float64, **kwargs)
        return sm.cayley_transform(pts).to(device=device, dtype=torch.float64 if dtype is None else dtype)

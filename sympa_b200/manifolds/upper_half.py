"""Upper half space H_n = {Z in Sym(n, C) : Im Z > 0} - mirror of sympa/manifolds/upper_half.py."""
import torch

from . import csym as sm
from .base import Manifold
from .metrics import MetricType
from .siegel_manifold import SiegelManifold


class UpperHalfManifold(SiegelManifold):
    ndim = 1
    reversible = False
    name = "Upper Half Space"
    kind = "upper"
    __scaling__ = Manifold.__scaling__.copy()

    def __init__(self, dims=2, ndim=2, metric=MetricType.RIEMANNIAN):
        super().__init__(dims=dims, ndim=ndim, metric=metric)

    def egrad2rgrad(self, z: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
        """Y G Y on the real and imaginary parts (upper_half.py:25-40)."""
        y = sm.imag(z)
        return sm.stick(y @ sm.real(u) @ y, y @ sm.imag(u) @ y)

    def projx(self, z: torch.Tensor) -> torch.Tensor:
        """Symmetrise, then clamp the eigenvalues of Im Z at eps for the points that need it
        (upper_half.py:42-66, csym_math.py:252-278)."""
        z = super().projx(z)
        y = sm.imag(z)
        eps = sm.EPS[y.dtype]
        lam, s = torch.linalg.eigh(y)
        ok = torch.all(lam > eps, dim=-1, keepdim=True)
        y_t = s @ torch.diag_embed(lam.clamp(min=eps)) @ s.transpose(-1, -2)
        self._count_projected(ok)
        return sm.stick(sm.real(z), torch.where(ok.unsqueeze(-1), y, y_t))

    def inner(self, z, u, v=None, *, keepdim=False):
        """tr[Y^-1 u Y^-1 conj(v)] (upper_half.py:68-91); returns (b, 2, 1, 1)."""
        if v is None:
            v = u
        yi = torch.linalg.inv(sm.imag(z)).to(torch.complex128 if z.dtype == torch.float64 else torch.complex64)
        res = yi @ sm.to_complex(u) @ yi @ sm.to_complex(v).conj()
        tr = torch.diagonal(res.real, dim1=-2, dim2=-1).sum(-1).reshape(-1, 1, 1)
        return sm.stick(tr, tr)

    def _check_point_on_manifold(self, z: torch.Tensor, *, atol=1e-5, rtol=1e-5):
        if not self._check_matrices_are_symmetric(z, atol=atol, rtol=rtol):  # upper_half.py:93-114
            return False, "Matrices are not symmetric"
        ok = bool((torch.det(z[..., 1, :, :]) > 0).all())
        return ok, (None if ok else "'x' determinant is not > 0")

    def random(self, *size, dtype=None, device=None, **kwargs) -> torch.Tensor:
        """X = sym(U(from_, to)), Y = I + sym(U(from_, to))   (upper_half.py:116-131)."""
        from_ = kwargs.get("from_", -0.001)
        to = kwargs.get("to", 0.001)
        n = self.dims
        dtype = torch.float64 if dtype is None else dtype  # the reference's global default (config.py:17-18)
        pert = sm.sym(torch.empty(size[0], n, n, dtype=torch.float64).uniform_(from_, to))
        y = torch.eye(n, dtype=torch.float64).unsqueeze(0) + pert
        x = sm.sym(torch.empty(size[0], n, n, dtype=torch.float64).uniform_(from_, to))
        return sm.stick(x, y).to(device=device, dtype=dtype)

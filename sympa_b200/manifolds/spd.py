"""Symmetric positive definite matrices with the affine-invariant metric.

The reference takes this manifold from geoopt (sympa/embeddings.py:6,142; un-vendored and unpinned,
README.md:40), so its behaviour is restated from geoopt's documented formulas - PARITY UNPINNED
(SURVEY.md F5).  `dist` runs on the CUDA hot path; `--metric` is ignored for spd as in the reference
(embeddings.py:153-154)."""
import torch

from .. import ops
from . import csym as sm
from .base import Manifold


class SymmetricPositiveDefinite(Manifold):
    ndim = 2
    reversible = False
    name = "SymmetricPositiveDefinite"
    kind = "spd"
    __scaling__ = Manifold.__scaling__.copy()

    def __init__(self, dims=None, **kwargs):
        super().__init__()
        self.dims = dims
        self.projected_points = 0

    def dist(self, x: torch.Tensor, y: torch.Tensor, *, keepdim=False) -> torch.Tensor:
        """|| log(X^-1/2 Y X^-1/2) ||_F, points (b, n, n); returns (b,) or (b, 1, 1) with keepdim."""
        d, _ = ops.dist("spd", "riem", x, y)
        return d.reshape(-1, 1, 1) if keepdim else d

    def vvd(self, x, y):
        """log of the eigenvalues of X^-1 Y, ascending (b, n)."""
        with torch.no_grad():
            _, v, _ = ops.forward_raw("spd", "riem", z1=x, z2=y)
        return v

    def dist_from_table(self, table, idx, sync_grad=False, accumulator=None):
        d, _ = ops.table_dist("spd", "riem", table, idx, sync_grad=sync_grad, accumulator=accumulator)
        return d

    def table_grad_accumulator(self, table, scatter_sms=None):
        """see sympa_b200.ops.TableGradAccumulator"""
        return ops.TableGradAccumulator("spd", table, scatter_sms)

    def dist_matrix(self, table, row_begin=0, row_count=None):
        """all-pairs distances between the rows of `table` (see SiegelManifold.dist_matrix)"""
        return ops.dist_matrix("spd", "riem", table, row_begin, row_count)

    def egrad2rgrad(self, x, u):
        return x @ sm.sym(u) @ x.transpose(-1, -2)

    def proju(self, x, u):
        return sm.sym(u)

    def projx(self, x):
        s = sm.sym(x)
        lam, q = torch.linalg.eigh(s)
        return q @ torch.diag_embed(lam.clamp(min=sm.EPS[x.dtype])) @ q.transpose(-1, -2)

    def retr(self, x, u):
        return sm.sym(x + u + 0.5 * u @ torch.linalg.solve(x, u))

    def transp(self, x, y, v):
        return v

    def inner(self, x, u, v=None, *, keepdim=False):
        if v is None:
            v = u
        xi_u = torch.linalg.solve(x, u)
        xi_v = torch.linalg.solve(x, v)
        tr = torch.diagonal(xi_u @ xi_v, dim1=-2, dim2=-1).sum(-1)
        return tr.reshape(-1, 1, 1) if keepdim else tr

    def _check_point_on_manifold(self, x, *, atol=1e-5, rtol=1e-5):
        if not torch.allclose(x, x.transpose(-1, -2), atol=atol, rtol=rtol):
            return False, "`x != x.transpose` with atol={}, rtol={}".format(atol, rtol)
        ok = bool((torch.linalg.eigvalsh(x) > -atol).all())
        return ok, (None if ok else "eigenvalues of x are not all greater than 0.")

    def random(self, *size, dtype=None, device=None, **kwargs):
        t = sm.sym(0.5 * torch.randn(*size, dtype=torch.float64))
        lam, q = torch.linalg.eigh(t)
        return (q @ torch.diag_embed(lam.exp()) @ q.transpose(-1, -2)).to(device=device, dtype=torch.float64 if dtype is None else dtype)

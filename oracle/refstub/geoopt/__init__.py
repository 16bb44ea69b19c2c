"""Minimal stand-in for the un-vendored `geoopt` dependency (TEST INFRASTRUCTURE ONLY).

It exists so that /root/reference (fedelopez77/sympa) can be imported unmodified in the
build container to generate golden vectors (oracle/gen_golden.py).  It is never imported by
the product package `sympa_b200`.  Surface listed in SURVEY.md section 8(c).
"""
import torch
from . import linalg, manifolds, optim  # noqa: F401
from .manifolds.base import Manifold  # noqa: F401


class ManifoldParameter(torch.nn.Parameter):
    def __new__(cls, data=None, manifold=None, requires_grad=True):
        inst = torch.nn.Parameter.__new__(cls, data, requires_grad)
        inst.manifold = manifold
        return inst


class _Vector(Manifold):
    def __init__(self, *a, **k):
        super().__init__()


Euclidean = PoincareBall = Lorentz = Sphere = _Vector


class ProductManifold(_Vector):
    pass

from .base import Manifold


class SymmetricPositiveDefinite(Manifold):
    pass

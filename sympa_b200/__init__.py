"""sympa_b200 - B200-native (sm_100a CUDA, FP64) Siegel / SPD pair-distance hot path of
fedelopez77/sympa behind the reference's `manifold.dist` / metric API.

There is no CPU fallback: every distance goes through libsympa_b200.so (C ABI in
include/sympa_b200.h); importing the ops without the built library raises.
"""
from .manifolds import (  # noqa: F401
    BoundedDomainManifold,
    Metric,
    MetricType,
    SiegelManifold,
    SymmetricPositiveDefinite,
    UpperHalfManifold,
)
from . import ops  # noqa: F401

__version__ = "0.1.0"

from .metrics import (  # noqa: F401
    FinslerInfinityMetric,
    FinslerMinimumEntropyMetric,
    FinslerOneMetric,
    FinslerWeightedSumMetric,
    Metric,
    MetricType,
    RiemannianMetric,
)
from .siegel_manifold import SiegelManifold  # noqa: F401
from .upper_half import UpperHalfManifold  # noqa: F401
from .bounded_domain import BoundedDomainManifold  # noqa: F401
from .spd import SymmetricPositiveDefinite  # noqa: F401

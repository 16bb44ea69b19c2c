from . import base, symmetric_positive_definite  # noqa: F401

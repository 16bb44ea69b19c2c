from . import batch_linalg  # noqa: F401

"""Import the UNMODIFIED reference from /root/reference behind the stubs in oracle/refstub.

TEST INFRASTRUCTURE ONLY - works only where /root/reference exists (the build container).
Nothing that runs on the GPU box may import this module.
"""
import os
import sys
import torch

REFERENCE_ROOT = os.environ.get("SYMPA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sympa"))


def import_reference():
    """Returns the reference's `sympa` package (torch.symeig shimmed onto torch.linalg.eigh)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    stub = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstub")
    for p in (stub, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(torch, "_sympa_symeig_shim"):
        # torch.symeig(y, eigenvectors=True, upper=True) was removed; same LAPACK routine as eigh(UPLO='U')
        def symeig(y, eigenvectors=True, upper=True):
            return torch.linalg.eigh(y, UPLO="U" if upper else "L")
        torch.symeig = symeig
        torch._sympa_symeig_shim = True
    import sympa  # noqa: F401
    import sympa.manifolds  # noqa: F401
    return sys.modules["sympa"]

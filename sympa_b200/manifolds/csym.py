"""Small complex-symmetric helpers on (b, 2, n, n) real tensors used by the NON-hot-path manifold
methods (projx, egrad2rgrad, random, inner).  Torch ops; mirrors the parts of
sympa/math/csym_math.py and sympa/math/cayley_transform.py those methods need."""
import torch

EPS = {torch.float32: 4e-3, torch.float64: 1e-5}  # sympa/config.py:19


def real(z):
    return z[:, 0]


def imag(z):
    return z[:, 1]


def stick(a, b):
    return torch.stack((a, b), dim=1)


def sym(x):
    return 0.5 * (x + x.transpose(-1, -2))


def to_symmetric(z):  # csym_math.py:131-138
    return stick(sym(real(z)), sym(imag(z)))


def to_complex(z):
    return torch.complex(real(z), imag(z))


def from_complex(c):
    return stick(c.real, c.imag)


def cayley_transform(z):  # cayley_transform.py:10-24
    c = to_complex(z)
    eye = torch.eye(c.shape[-1], dtype=c.dtype, device=c.device)
    return from_complex((c - 1j * eye) @ torch.linalg.inv(c + 1j * eye))


def inverse_cayley_transform(z):  # cayley_transform.py:27-40
    c = to_complex(z)
    eye = torch.eye(c.shape[-1], dtype=c.dtype, device=c.device)
    return from_complex(1j * (eye + c) @ torch.linalg.inv(eye - c))


def takagi(z):
    """Takagi factorisation Z = conj(S) D S^H of complex symmetric matrices through the real
    symmetric compound matrix [(A, B), (B, -A)] (takagi_factorization.py:45-64).  Returns the
    values ascending (b, n) and S as a complex tensor (b, n, n)."""
    a, b = real(z), imag(z)
    n = a.shape[-1]
    m = torch.cat((torch.cat((a, b), -1), torch.cat((b, -a), -1)), -2)
    lam, q = torch.linalg.eigh(m)
    right = q[..., :, n:]
    s = torch.complex(right[..., :n, :], -right[..., n:, :])
    return lam[..., n:], s

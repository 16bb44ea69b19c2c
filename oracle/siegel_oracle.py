"""CPU oracle for the Siegel / SPD pair-distance hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-torch (CPU, float64) restatement of the algorithm that
fedelopez77/sympa runs for `manifold.dist(z1, z2)`; every function cites the reference
file:line it follows (paths are relative to /root/reference).  It is the checker for the CUDA
path - only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product package `sympa_b200` never does.

Pinning: `oracle/gen_golden.py` runs the UNMODIFIED reference (imported through
`oracle/refstub`) in the build container and stores its outputs in `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this restatement against those files and against the
reference's own known-answer tests (tests/test_math.py:175-287).  The backward of the
reference is torch autograd through these same ops, so gradients of the oracle are obtained
with autograd as well.

The `spd` distance is geoopt's (un-vendored, unpinned: README.md:40) - restated from its
documented formula; PARITY UNPINNED for that function.

Complex matrices are real tensors (b, 2, n, n): index 0 real part, 1 imaginary part
(sympa/math/csym_math.py:1-8).
"""
import torch

# sympa/config.py:19
EPS = {torch.float32: 4e-3, torch.float64: 1e-5}
METRICS = ("riem", "fone", "finf", "fmin", "wsum")  # sympa/manifolds/metrics.py:6-12


# --------------------------------------------------------------------------- complex algebra
def re(z):  # csym_math.py:15-22
    return z[:, 0]


def im(z):  # csym_math.py:25-31
    return z[:, 1]


def cplx(a, b):  # csym_math.py:34-41 (stick)
    return torch.stack((a, b), dim=1)


def sym(x):  # geoopt.linalg.batch_linalg.sym, used at csym_math.py:12
    return 0.5 * (x + x.transpose(-1, -2))


def to_symmetric(z):  # csym_math.py:131-138
    return cplx(sym(re(z)), sym(im(z)))


def cmatmul(x, y):  # csym_math.py:91-115 (bmm): four real batched products
    a, b = re(x), im(x)
    c, d = re(y), im(y)
    return cplx(a @ c - b @ d, a @ d + b @ c)


def cmatmul3(x, y, z):  # csym_math.py:118-128 (bmm3)
    return cmatmul(cmatmul(x, y), z)


def eye_like(z):  # csym_math.py:334-343 (identity_like): real identity, zero imaginary part
    b, _, n, _ = z.shape
    i = torch.eye(n, dtype=z.dtype, device=z.device).expand(b, n, n)
    return cplx(i, torch.zeros_like(i))


def cinverse(z):
    """csym_math.py:197-249 - Falkenberg's inverse of A + iC by real inversions:
    R = A^-1 C, U = (C R + A)^-1, V = -R U.  The reference special-cases matrices whose real
    (resp. imaginary) part has an all-zero FIRST ROW (:216-217) by swapping in the identity and
    patching the result afterwards (:221-247).  The patch for a zero real part returns
    +i C^-1 where the true inverse is -i C^-1; the restatement keeps that
    behaviour because `dist` only uses singular values, which do not see the sign.
    """
    a, c = re(z), im(z)
    n = a.shape[-1]
    zr = (a[:, 0] == 0).sum(-1) == n
    zi = (c[:, 0] == 0).sum(-1) == n
    any_zr, any_zi = bool(zr.any()), bool(zi.any())
    eye = torch.eye(n, dtype=a.dtype, device=a.device)
    if any_zr:
        c_inv = torch.linalg.inv(c)
        a = a + zr.to(a.dtype).reshape(-1, 1, 1) * eye
    if any_zi:
        a_inv = torch.linalg.inv(a)
        c = c + zi.to(c.dtype).reshape(-1, 1, 1) * eye
    r = torch.linalg.inv(a) @ c
    u = torch.linalg.inv(c @ r + a)
    v = -(r @ u)
    if any_zr:
        m = zr.reshape(-1, 1, 1)
        u = torch.where(m, torch.zeros_like(u), u)
        v = torch.where(m, c_inv, v)
    if any_zi:
        m = zi.reshape(-1, 1, 1)
        u = torch.where(m, a_inv, u)
        v = torch.where(m, torch.zeros_like(v), v)
    return cplx(u, v)


def symeig(y):  # csym_math.py:281-292; torch.symeig(upper=True) == eigh(UPLO='U'), ascending
    return torch.linalg.eigh(y, UPLO="U")


def matrix_sqrt(y):  # csym_math.py:509-520
    lam, s = symeig(y)
    return s @ torch.diag_embed(torch.sqrt(lam)) @ s.transpose(-1, -2)


def compound_symmetric(z):  # csym_math.py:421-436: [(A, B), (B, -A)]
    a, b = re(z), im(z)
    return torch.cat((torch.cat((a, b), -1), torch.cat((b, -a), -1)), -2)


# --------------------------------------------------------------------------- Cayley / Takagi
def cayley_transform(z):  # sympa/math/cayley_transform.py:10-24   (Z - iI)(Z + iI)^-1
    ident = eye_like(z)
    i_ident = cplx(im(ident), re(ident))
    return cmatmul(z - i_ident, cinverse(z + i_ident))


def inverse_cayley_transform(z):  # cayley_transform.py:27-40   i(I + Z)(I - Z)^-1
    ident = eye_like(z)
    s = ident + z
    return cmatmul(cplx(-im(s), re(s)), cinverse(ident - z))  # multiply_by_i: csym_math.py:164-169


def takagi_values(z):
    """sympa/math/takagi_factorization.py:66-75: top-n eigenvalues (ascending) of the 2n x 2n
    real symmetric compound matrix."""
    lam, _ = symeig(compound_symmetric(z))
    return lam[:, z.shape[-1]:]


def takagi_factorize(z):
    """takagi_factorization.py:45-64: values and the unitary S with  Z = conj(S) D S^H."""
    n = z.shape[-1]
    lam, q = symeig(compound_symmetric(z))
    right = q[..., :, n:]
    return lam[:, n:], cplx(right[..., :n, :], -right[..., n:, :])


# --------------------------------------------------------------------------- metrics
def fmin_weights(n, dtype=torch.float64):
    """metrics.py:86-89: 2*(n+1 - arange(n+1, 1, -1)) = [0, 2, ..., 2(n-1)], applied to the
    ASCENDING vector-valued distance."""
    return (2 * (n + 1 - torch.arange(n + 1, 1, -1))).to(dtype)


def compute_metric(v, metric, wsum_weights=None):
    """metrics.py:42-121.  v: (b, n) ascending."""
    if metric == "riem":
        return torch.norm(v, dim=-1)            # :51
    if metric == "fone":
        return v.sum(-1)                        # :64
    if metric == "finf":
        return v[:, -1]                         # :78
    if metric == "fmin":
        return (fmin_weights(v.shape[-1], v.dtype).unsqueeze(0) * v).sum(-1)   # :99
    if metric == "wsum":
        w = torch.relu(wsum_weights.reshape(1, -1))                            # :118
        return (w * v).sum(-1)                                                 # :119-120
    raise ValueError(metric)


# --------------------------------------------------------------------------- distances
def vvd_from_takagi(d):
    """siegel_manifold.py:63-70 (the two asserts are host checks, kept as asserts)."""
    eps = EPS[d.dtype]
    assert torch.all(d >= 0 - eps), "eigenvalues below -eps"
    assert torch.all(d <= 1.01), "eigenvalues above 1.01"
    return torch.log((1 + d) / (1 - d).clamp(min=eps))


def upper_vvd(z1, z2):
    """siegel_manifold.py:41-70: Z3 = Y1^-1/2 (Z2 - X1) Y1^-1/2, W = cayley(Z3), Takagi values
    of W ascending, v = log((1 + d) / (1 - d))."""
    x1, y1 = re(z1), im(z1)
    isq = torch.linalg.inv(matrix_sqrt(y1))                 # :53
    isq = cplx(isq, torch.zeros_like(isq))                  # :54
    z3 = cmatmul3(isq, z2 - cplx(x1, torch.zeros_like(x1)), isq)   # :55-56
    w = cayley_transform(z3)                                # :58
    return vvd_from_takagi(takagi_values(w))                # :60-70


def upper_dist(z1, z2, metric="riem", wsum_weights=None):   # siegel_manifold.py:71
    return compute_metric(upper_vvd(z1, z2), metric, wsum_weights)


def bounded_vvd(z1, z2):  # sympa/manifolds/bounded_domain.py:27-39
    return upper_vvd(inverse_cayley_transform(z1), inverse_cayley_transform(z2))


def bounded_dist(z1, z2, metric="riem", wsum_weights=None):
    return compute_metric(bounded_vvd(z1, z2), metric, wsum_weights)


def spd_vvd(x, y):
    """geoopt SymmetricPositiveDefinite.dist (affine-invariant metric), call site
    sympa/embeddings.py:142 / model.py:38.  Restated from the documented formula
    || log(X^-1/2 Y X^-1/2) ||_F; the vector returned here is log(eig(X^-1/2 Y X^-1/2))
    ascending.  PARITY UNPINNED (no geoopt in the container, no reference test)."""
    lam, s = torch.linalg.eigh(x)
    isq = s @ torch.diag_embed(lam.rsqrt()) @ s.transpose(-1, -2)
    mid = sym(isq @ y @ isq)
    return torch.log(torch.linalg.eigvalsh(mid))


def spd_dist(x, y):
    return torch.norm(spd_vvd(x, y), dim=-1)


def dist(kind, z1, z2, metric="riem", wsum_weights=None):
    if kind == "upper":
        return upper_dist(z1, z2, metric, wsum_weights)
    if kind == "bounded":
        return bounded_dist(z1, z2, metric, wsum_weights)
    if kind == "spd":
        return spd_dist(z1, z2)
    raise ValueError(kind)


def vvd(kind, z1, z2):
    return {"upper": upper_vvd, "bounded": bounded_vvd, "spd": spd_vvd}[kind](z1, z2)


def dist_and_grads(kind, z1, z2, metric="riem", wsum_weights=None, grad_out=None):
    """Forward + autograd backward, as the reference does it (runner.py:101-105).  Returns
    dist, dL/dz1, dL/dz2 (and dL/dw for wsum) for L = sum(grad_out * dist)."""
    z1 = z1.detach().clone().requires_grad_(True)
    z2 = z2.detach().clone().requires_grad_(True)
    w = None
    if wsum_weights is not None:
        w = wsum_weights.detach().clone().requires_grad_(True)
    d = dist(kind, z1, z2, metric, w)
    go = torch.ones_like(d) if grad_out is None else grad_out
    (d * go).sum().backward()
    return d.detach(), z1.grad, z2.grad, (None if w is None else w.grad)


# --------------------------------------------------------------------------- loss / model glue
def distortion_loss(graph_dist, manifold_dist):  # sympa/losses.py:16-19
    return torch.abs(torch.pow(manifold_dist / graph_dist, 2) - 1).sum()


# --------------------------------------------------------------------------- point generators
def upper_random(num, n, from_=-1e-3, to=1e-3, generator=None, dtype=torch.float64):
    """sympa/manifolds/upper_half.py:116-131: X = sym(U(from,to)), Y = I + sym(U(from,to)).
    (The reference draws the imaginary perturbation first.)"""
    pert = sym(torch.empty(num, n, n, dtype=dtype).uniform_(from_, to, generator=generator))
    y = torch.eye(n, dtype=dtype).unsqueeze(0) + pert
    x = sym(torch.empty(num, n, n, dtype=dtype).uniform_(from_, to, generator=generator))
    return cplx(x, y)


def bounded_random(num, n, from_=-1e-3, to=1e-3, generator=None, dtype=torch.float64):
    return cayley_transform(upper_random(num, n, from_, to, generator, dtype))  # bounded_domain.py:152-160


def upper_spread(num, n, generator=None, dtype=torch.float64, scale=0.5):
    """SURVEY.md 8(d) 'spread regime': X = sym(0.5 N(0,1)), Y = C C^T + 0.5 I, C = 0.5 N(0,1)."""
    x = sym(scale * torch.randn(num, n, n, dtype=dtype, generator=generator))
    c = scale * torch.randn(num, n, n, dtype=dtype, generator=generator)
    y = c @ c.transpose(-1, -2) + 0.5 * torch.eye(n, dtype=dtype)
    return cplx(x, sym(y))


def spd_spread(num, n, generator=None, dtype=torch.float64, scale=0.5):
    c = scale * torch.randn(num, n, n, dtype=dtype, generator=generator)
    return sym(c @ c.transpose(-1, -2) + 0.5 * torch.eye(n, dtype=dtype))


# --------------------------------------------------------------------------- optimizer-side ops
def upper_egrad2rgrad(z, u):  # upper_half.py:25-40   Y G Y on real and imaginary parts
    y = im(z)
    return cplx(y @ re(u) @ y, y @ im(u) @ y)


def bounded_egrad2rgrad(z, u):  # bounded_domain.py:41-53, :163-170   A u A, A = I - conj(Z) Z
    zc = cplx(re(z), -im(z))
    a = eye_like(z) - cmatmul(zc, z)
    return cmatmul3(a, u, a)


def positive_conjugate_projection(y):  # csym_math.py:252-278
    eps = EPS[y.dtype]
    lam, s = symeig(y)
    y_t = s @ torch.diag_embed(lam.clamp(min=eps)) @ s.transpose(-1, -2)
    ok = torch.all(lam > eps, dim=-1, keepdim=True)
    return torch.where(ok.unsqueeze(-1), y, y_t), ok


def upper_projx(z):  # upper_half.py:42-66 (+ siegel_manifold.py:130-137)
    z = to_symmetric(z)
    y_t, _ = positive_conjugate_projection(im(z))
    return cplx(re(z), y_t)


def bounded_projx(z):
    """bounded_domain.py:55-84 with the factoriser fix of SURVEY.md F3 (the shipped reference
    builds TakagiFactorization(return_eigenvectors=False) and crashes at :70; the behaviour
    restated here is the one its tests/test_bounded_domain.py:17-66 describe)."""
    eps = EPS[z.dtype]
    z = to_symmetric(z)
    d, s = takagi_factorize(z)
    d_t = d.clamp(max=1 - eps)
    dm = torch.diag_embed(d_t)
    dm = cplx(dm, torch.zeros_like(dm))
    s_conj = cplx(re(s), -im(s))
    s_h = s_conj.transpose(-1, -2)
    z_t = cmatmul3(s_conj, dm, s_h)
    ok = torch.all(d < 1 - eps, dim=-1).reshape(-1, 1, 1, 1)
    return torch.where(ok, z, z_t)


def rsgd_step(kind, p, grad, lr, weight_decay=0.0):
    """geoopt.optim.RiemannianSGD.step without momentum (train.py:68; restated from geoopt's
    documented update, PARITY UNPINNED):  g += wd * p;  r = egrad2rgrad(p, g);
    p <- retr(p, -lr * r) with retr = projx(p + u) (siegel_manifold.py:74-87)."""
    g = grad + weight_decay * p
    if kind == "upper":
        return upper_projx(p - lr * upper_egrad2rgrad(p, g))
    if kind == "bounded":
        return bounded_projx(p - lr * bounded_egrad2rgrad(p, g))
    if kind == "spd":
        # geoopt SymmetricPositiveDefinite (un-vendored; restated from its documented formulas, PARITY UNPINNED):
        # egrad2rgrad(x, u) = x sym(u) x^T;  retr(x, u) = sym(x + u + u x^-1 u / 2)
        u = -lr * (p @ sym(g) @ p.transpose(-1, -2))
        return sym(p + u + 0.5 * u @ torch.linalg.solve(p, u))
    raise ValueError(kind)

"""numpy model of the algorithm the CUDA kernels implement.  TEST / DESIGN INFRASTRUCTURE ONLY.

The reference (oracle/siegel_oracle.py) gets the Takagi values from eigen-decompositions;
the CUDA path uses a different but mathematically equivalent route that suits registers:

  Y1 = L L^T (Cholesky), Li = L^-1
  A = Li (X2 - X1) Li^T,  B = Li Y2 Li^T,  C = B + I        (Z2 moved so that Z1 -> iI)
  N = (A + iC)^-1 = P - iQ,  Q = (C + A C^-1 A)^-1,  P = C^-1 A Q   (two real SPD solves)
  W = I - 2iN = (I - 2Q) - 2iP                               (Cayley transform of A + iB)
  d = singular values of W by one-sided (Hestenes) Jacobi: W V = G, d_k = |G[:,k]|
  v_k = log((1 + d_k) / max(1 - d_k, eps)), metric reductions as in the reference.

and an analytic backward (no autograd): G_W = U diag(dL/dd) V^H, then the chain back
through the inverse, the congruence and the Cholesky factor.  This file states those formulas
once, in numpy, so that they can be checked against the oracle on the CPU
(tests/test_kernel_model.py) independently of any CUDA code.
"""
import numpy as np

EPS64 = 1e-5


def chol_inv(y):
    """lower L with y = L L^T, and L^-1."""
    l = np.linalg.cholesky(y)
    return l, np.linalg.inv(l)


def inv_spd_real(e, f):
    """(E + iF)^-1 = U + iV for E SPD, F symmetric:  U = (E + F E^-1 F)^-1,  V = -E^-1 F U."""
    ef = np.linalg.solve(e, f)
    u = np.linalg.inv(e + f @ ef)
    u = 0.5 * (u + u.T)
    v = -ef @ u
    return u, 0.5 * (v + v.T)


def jacobi_svd(w, max_sweeps=30):
    """One-sided Jacobi on a complex n x n matrix.  Returns G = W V (orthogonal columns), V."""
    n = w.shape[0]
    g = w.astype(np.complex128).copy()
    v = np.eye(n, dtype=np.complex128)
    for sweep in range(max_sweeps):
        worst = 0.0
        for p in range(n - 1):
            for q in range(p + 1, n):
                a = np.vdot(g[:, p], g[:, p]).real
                b = np.vdot(g[:, q], g[:, q]).real
                gam = np.vdot(g[:, p], g[:, q])          # g_p^H g_q
                g2 = gam.real ** 2 + gam.imag ** 2
                if g2 <= 1e-32 * a * b or g2 == 0.0:
                    continue
                worst = max(worst, g2 / (a * b))
                delta = 0.5 * (b - a)
                den = abs(delta) + np.sqrt(delta * delta + g2)
                tc = np.copysign(1.0, delta) * gam / den   # complex tangent
                c = 1.0 / np.sqrt(1.0 + (tc.real ** 2 + tc.imag ** 2))
                s = c * tc
                gp, gq = g[:, p].copy(), g[:, q].copy()
                g[:, p] = c * gp - np.conj(s) * gq
                g[:, q] = s * gp + c * gq
                vp, vq = v[:, p].copy(), v[:, q].copy()
                v[:, p] = c * vp - np.conj(s) * vq
                v[:, q] = s * vp + c * vq
        if worst < 1e-16:      # every rotation of this sweep was tiny: off-diagonals are now ~worst^2
            break
    return g, v, sweep + 1


def metric_weights(metric, n, wsum_w=None):
    """per-RANK weights (rank 0 = smallest v) for the linear metrics."""
    if metric == "fone":
        return np.ones(n)
    if metric == "finf":
        w = np.zeros(n)
        w[-1] = 1.0
        return w
    if metric == "fmin":
        return 2.0 * np.arange(n)
    if metric == "wsum":
        return np.maximum(np.asarray(wsum_w, dtype=np.float64).reshape(-1), 0.0)
    raise ValueError(metric)


def vvd_and_dist(d_sorted, metric, wsum_w=None, eps=EPS64):
    """returns v (ascending), dist, d(dist)/d(d_k) (ascending order)."""
    n = d_sorted.shape[0]
    om = 1.0 - d_sorted
    clamped = om < eps
    v = np.where(clamped, np.log((1.0 + d_sorted) / eps), np.log1p(2.0 * d_sorted / np.where(clamped, 1.0, om)))
    dv = np.where(clamped, 1.0 / (1.0 + d_sorted), 2.0 / ((1.0 - d_sorted) * (1.0 + d_sorted)))
    if metric == "riem":
        dist = np.sqrt(np.sum(v * v))
        gv = v / dist if dist > 0 else np.zeros(n)
    else:
        wt = metric_weights(metric, n, wsum_w)
        dist = np.sum(wt * v)
        gv = wt
    return v, dist, gv * dv


def upper_normal_form(x1, y1, x2, y2):
    l, li = chol_inv(y1)
    t1 = li @ (x2 - x1)
    t2 = li @ y2
    a = t1 @ li.T
    b = t2 @ li.T
    a, b = 0.5 * (a + a.T), 0.5 * (b + b.T)
    return li, t1, t2, a, b


def upper_pair(x1, y1, x2, y2, metric="riem", wsum_w=None, want_grad=True):
    """One pair.  Returns v (ascending), dist and - if want_grad - the gradients of dist with
    respect to X1, Y1, X2, Y2 as SYMMETRIC matrices (the gradient with respect to the (i,j)
    and (j,i) entries of a symmetric parameter, each taken as independent, symmetrised), plus
    d(dist)/d(wsum weight) for wsum."""
    n = x1.shape[0]
    li, t1, t2, a, b = upper_normal_form(x1, y1, x2, y2)
    c = b + np.eye(n)
    q, p = inv_spd_real(c, -a)          # (C - iA)^-1 = Q + iP  ->  (A + iC)^-1 = P - iQ
    w = (np.eye(n) - 2.0 * q) - 2.0j * p
    g, v_, _ = jacobi_svd(w)
    sig = np.sqrt(np.sum(g.real ** 2 + g.imag ** 2, axis=0))
    order = np.argsort(sig, kind="stable")
    d_sorted = sig[order]
    vv, dist, gd = vvd_and_dist(d_sorted, metric, wsum_w)
    if not want_grad:
        return vv, dist
    # dL/d sigma_k in the ORIGINAL column order
    gsig = np.zeros(n)
    gsig[order] = gd
    coef = np.where(sig > 0, gsig / np.where(sig > 0, sig, 1.0), 0.0)
    gw = (g * coef[None, :]) @ v_.conj().T            # G_W = sum_k gsig_k u_k v_k^H  (G_R + i G_S)
    gw = 0.5 * (gw + gw.T)
    gn = 2.0j * gw                                    # W = I - 2iN
    nbar = p + 1.0j * q                               # conj(N), N = P - iQ
    gm = -nbar @ gn @ nbar                            # N = M^-1
    ga, gb = gm.real, gm.imag                         # M = A + i(B + I)
    ga, gb = 0.5 * (ga + ga.T), 0.5 * (gb + gb.T)
    gx2 = li.T @ ga @ li
    gy2 = li.T @ gb @ li
    gli = 2.0 * (ga @ t1 + gb @ t2)
    k = np.tril(gli @ li.T)
    ks = 0.5 * (k + k.T) - 0.5 * np.diag(np.diag(k))
    gy1 = -li.T @ ks @ li
    grads = [-gx2, gy1, gx2, gy2]
    grads = [0.5 * (m + m.T) for m in grads]
    gwts = None
    if metric == "wsum":
        gwts = np.where(np.asarray(wsum_w).reshape(-1) > 0, vv, 0.0)
    return vv, dist, grads, gwts


def bounded_to_upper(zr, zi):
    """Z = i(I + z)(I - z)^-1 = i(2 (I - z)^-1 - I).  Returns X, Y and N = (I - z)^-1."""
    n = zr.shape[0]
    u, v = inv_spd_real(np.eye(n) - zr, -zi)
    return -2.0 * v, 2.0 * u - np.eye(n), u + 1.0j * v


def bounded_pair(z1r, z1i, z2r, z2i, metric="riem", wsum_w=None, want_grad=True):
    x1, y1, n1 = bounded_to_upper(z1r, z1i)
    x2, y2, n2 = bounded_to_upper(z2r, z2i)
    out = upper_pair(x1, y1, x2, y2, metric, wsum_w, want_grad)
    if not want_grad:
        return out
    vv, dist, (gx1, gy1, gx2, gy2), gwts = out
    res = []
    for nn, gx, gy in ((n1, gx1, gy1), (n2, gx2, gy2)):
        gn = 2.0 * gy - 2.0j * gx                     # X = -2 Im N, Y = 2 Re N - I
        gz = nn.conj() @ gn @ nn.conj()               # N = (I - z)^-1
        res += [0.5 * (gz.real + gz.real.T), 0.5 * (gz.imag + gz.imag.T)]
    return vv, dist, res, gwts


def spd_pair(x, y, want_grad=True):
    """Affine-invariant SPD distance sqrt(sum log^2 lambda_i(X^-1 Y)) through two Cholesky factors:
    lambda = sigma^2 of  Gm = Lx^-1 Ly  (one-sided Jacobi on a real matrix)."""
    n = x.shape[0]
    lx, lxi = chol_inv(x)
    ly = np.linalg.cholesky(y)
    gm = lxi @ ly
    g, v_, _ = jacobi_svd(gm)
    g, v_ = g.real, v_.real
    sig = np.sqrt(np.sum(g * g, axis=0))
    order = np.argsort(sig, kind="stable")
    vv = 2.0 * np.log(sig[order])
    dist = np.sqrt(np.sum(vv * vv))
    if not want_grad:
        return vv, dist
    gv = vv / dist if dist > 0 else np.zeros(n)
    gsig = np.zeros(n)
    gsig[order] = gv * 2.0 / sig[order]
    ggm = (g * (gsig / sig)[None, :]) @ v_.T          # dL/dGm = U diag(gsig) V^T
    # Gm = Lxi Ly
    gly = np.tril(lxi.T @ ggm)
    glxi = np.tril(ggm @ ly.T)
    # Ly = chol(Y):  standard Cholesky backward
    def chol_bwd(l, gl):
        ph = np.tril(l.T @ gl)
        ph = ph - 0.5 * np.diag(np.diag(ph))
        linv = np.linalg.inv(l)
        s = linv.T @ ph @ linv
        return 0.5 * (s + s.T)
    gy = chol_bwd(ly, gly)
    k = np.tril(glxi @ lxi.T)
    ks = 0.5 * (k + k.T) - 0.5 * np.diag(np.diag(k))
    gx = -lxi.T @ ks @ lxi
    return vv, dist, [0.5 * (gx + gx.T), gy]

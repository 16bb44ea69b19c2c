"""TEST INFRASTRUCTURE - never imported by the product package.

Multi-precision (mpmath, 40 digits) evaluation of the upper-half-space distance, following the reference's
own op sequence - Z3 = Y1^-1/2 (Z2 - X1) Y1^-1/2 (siegel_manifold.py:53-58), W = (Z3 - iI)(Z3 + iI)^-1
(cayley_transform.py:10-24), Takagi values d = singular values of W (takagi_factorization.py:66-75),
v = log((1 + d) / (1 - d)) (siegel_manifold.py:69-70), Riemannian / Finsler-one reduction (metrics.py) -
and of its gradient by central differences at a step of 1e-15 (40 digits leave ~1e-25 of round-off).

Purpose: the reference's autograd gradient carries 1e-10 .. 3e-6 of relative noise (its
Y^-1/2 goes through an eigendecomposition whose backward divides by eigenvalue gaps), so "parity with the
reference" cannot be tested below that.  This oracle adjudicates: tests/test_mp_adjudication.py checks that the
kernels' gradient agrees with the multi-precision one to ~1e-11 while the reference's own is further away.
"""
import mpmath as mp

mp.mp.dps = 40


def _num(x):
    return x if isinstance(x, mp.mpf) else mp.mpf(float(x))     # float64 inputs are taken exactly


def _sym_from(a, n):
    return mp.matrix([[_num(a[i][j]) for j in range(n)] for i in range(n)])


def _inv_sqrt_spd(y, n):
    lam, q = mp.eigsy(y)
    d = mp.diag([1 / mp.sqrt(lam[i]) for i in range(n)])
    return q * d * q.T


def dist(z1, z2, metric="riem"):
    """z1, z2: nested sequences / arrays (2, n, n) of floats (real part, imaginary part)."""
    n = len(z1[0])
    x1, y1 = _sym_from(z1[0], n), _sym_from(z1[1], n)
    x2, y2 = _sym_from(z2[0], n), _sym_from(z2[1], n)
    r = _inv_sqrt_spd(y1, n)
    eye = mp.eye(n)
    z3 = r * (x2 - x1) * r + mp.mpc(0, 1) * (r * y2 * r)
    w = (z3 - mp.mpc(0, 1) * eye) * mp.inverse(z3 + mp.mpc(0, 1) * eye)
    d = mp.svd_c(w, compute_uv=False)
    v = [mp.log((1 + d[k]) / (1 - d[k])) for k in range(n)]
    if metric == "riem":
        return mp.sqrt(sum(t * t for t in v))
    if metric == "fone":
        return sum(v)
    raise ValueError(metric)


def sym_gradients(z1, z2, metric="riem", h=mp.mpf("1e-15")):
    """sym() of the gradient of dist with respect to every entry of z1 and z2 (entries (i,j) and (j,i) taken as
    independent, then symmetrised - the convention of the kernels and of sym(reference autograd gradient)),
    as two nested lists (2, n, n) of floats."""
    n = len(z1[0])
    base = [[[[mp.mpf(float(z[c][i][j])) for j in range(n)] for i in range(n)] for c in range(2)] for z in (z1, z2)]
    out = [[[[0.0] * n for _ in range(n)] for _ in range(2)] for _ in range(2)]
    for which in range(2):
        for c in range(2):
            for i in range(n):
                for j in range(i + 1):
                    vals = []
                    for sgn in (1, -1):
                        pert = [[[row[:] for row in part] for part in z] for z in base]
                        pert[which][c][i][j] += sgn * h
                        if i != j:
                            pert[which][c][j][i] += sgn * h
                        vals.append(dist(pert[0], pert[1], metric))
                    dd = (vals[0] - vals[1]) / (2 * h)       # = g_ij + g_ji (i != j), g_ii (i == j)
                    g = float(dd / 2) if i != j else float(dd)
                    out[which][c][i][j] = g
                    out[which][c][j][i] = g
    return out

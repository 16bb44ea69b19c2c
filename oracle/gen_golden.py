"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference); the
output files are committed so the GPU box never needs the reference tree.

    python oracle/gen_golden.py            # writes the round-2 additions (n = 1, 7, 8, 9; optim.npz)
    python oracle/gen_golden.py --all      # rewrites all of tests/golden/

What is stored, per (manifold kind, n, input regime):
  z1, z2              inputs (b, 2, n, n) float64
  vvd                 the vector-valued distance the reference feeds to its metric (b, n)
  dist_<metric>       manifold.dist for riem / fone / finf / fmin / wsum
  g1_<metric>, g2_<metric>   autograd gradients of sum(go * dist) wrt z1 / z2 (raw, unsymmetrised)
  gw_wsum, wsum_w     wsum weights used and their gradient
  go                  the upstream gradient used
plus building blocks (cayley, inverse cayley, complex inverse, Takagi values) in blocks.npz and the
optimizer-side manifold functions (egrad2rgrad, projx, retr) in optim.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from import_reference import import_reference  # noqa: E402
import siegel_oracle as so  # noqa: E402  (only for the input generators)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
METRIC_NAMES = ["riem", "fone", "finf", "fmin", "wsum"]


def make_inputs(kind, n, regime, b, seed):
    g = torch.Generator().manual_seed(seed)
    if regime == "init":
        z1 = so.upper_random(b, n, generator=g)
        z2 = so.upper_random(b, n, generator=g)
    elif regime == "spread":
        z1 = so.upper_spread(b, n, generator=g)
        z2 = so.upper_spread(b, n, generator=g)
    elif regime == "mid":      # moderately separated points (d ~ 0.1 .. 0.7)
        z1 = so.upper_spread(b, n, generator=g, scale=0.25)
        z2 = so.upper_spread(b, n, generator=g, scale=0.25)
    else:
        raise ValueError(regime)
    return z1, z2


def gen_pairs(ns, seed, sympa_mods, b=12):
    """dist / vvd / gradient goldens for upper + bounded at the matrix sizes `ns`; seeds seed+1, seed+2, ... in
    (kind, n, regime) order."""
    UpperHalfManifold, BoundedDomainManifold, MetricType, sm, cayley_transform = sympa_mods
    for kind in ("upper", "bounded"):
        cls = UpperHalfManifold if kind == "upper" else BoundedDomainManifold
        for n in ns:
            for regime in ("init", "mid", "spread"):
                seed += 1
                z1, z2 = make_inputs(kind, n, regime, b, seed)
                if kind == "bounded":
                    z1, z2 = cayley_transform(z1), cayley_transform(z2)
                    z1, z2 = sm.to_symmetric(z1), sm.to_symmetric(z2)
                go = torch.empty(b).uniform_(0.5, 1.5, generator=torch.Generator().manual_seed(seed))
                rec = {"z1": z1.numpy(), "z2": z2.numpy(), "go": go.numpy()}
                for mname in METRIC_NAMES:
                    man = cls(dims=n, metric=MetricType.from_str(mname))
                    if mname == "wsum":
                        w = torch.linspace(-0.4, 1.3, n).reshape(1, n) if n > 1 else torch.tensor([[0.7]])
                        if n >= 3:
                            w[0, 1] = 0.0      # relu'(0) = 0 in torch
                        man.metric.weights.data = w.clone()
                        rec["wsum_w"] = w.numpy()
                    captured = {}
                    orig = man.metric.compute_metric

                    def spy(v, *a, _orig=orig, _c=captured, **k):
                        _c["v"] = v.detach().clone()
                        return _orig(v, *a, **k)

                    man.metric.compute_metric = spy
                    a1 = z1.clone().requires_grad_(True)
                    a2 = z2.clone().requires_grad_(True)
                    d = man.dist(a1, a2)
                    (d * go).sum().backward()
                    rec[f"dist_{mname}"] = d.detach().numpy()
                    rec[f"g1_{mname}"] = a1.grad.numpy()
                    rec[f"g2_{mname}"] = a2.grad.numpy()
                    rec["vvd"] = captured["v"].numpy()
                    if mname == "wsum":
                        rec["gw_wsum"] = man.metric.weights.grad.numpy()
                np.savez_compressed(os.path.join(OUT, f"{kind}_n{n}_{regime}.npz"), **rec)
                print(kind, n, regime, "dist_riem[:3] =", rec["dist_riem"][:3])
    return seed


def gen_optim(sympa_mods):
    """Optimizer-side functions of the unmodified reference (the RiemannianSGD row update calls them:
    geoopt RSGD -> manifold.egrad2rgrad, manifold.retr -> projx): UpperHalfManifold.egrad2rgrad / projx / retr
    (sympa/manifolds/upper_half.py:25-66, siegel_manifold.py:74-87) and BoundedDomainManifold.egrad2rgrad
    (bounded_domain.py:41-53).  BoundedDomainManifold.projx / retr crash in the reference as shipped
    and cannot be pinned.  Inputs include rows whose update leaves the manifold (projection active)."""
    UpperHalfManifold, BoundedDomainManifold, MetricType, sm, cayley_transform = sympa_mods
    rec = {}
    g = torch.Generator().manual_seed(4242)
    for n in (2, 3, 4, 6, 10):
        up = UpperHalfManifold(dims=n, metric=MetricType.from_str("riem"))
        bd = BoundedDomainManifold(dims=n, metric=MetricType.from_str("riem"))
        z = so.upper_spread(10, n, generator=g, scale=0.4)
        u = torch.randn(10, 2, n, n, generator=g)                      # raw (unsymmetrised) Euclidean gradient
        us = sm.to_symmetric(u)
        rec[f"upper_z_n{n}"] = z.numpy()
        rec[f"upper_u_n{n}"] = u.numpy()
        rec[f"upper_egrad2rgrad_n{n}"] = up.egrad2rgrad(z, u).numpy()
        rec[f"upper_egrad2rgrad_sym_n{n}"] = up.egrad2rgrad(z, us).numpy()
        # a big step: some imaginary parts lose positive definiteness, projx clamps their spectrum
        step = -2.5 * up.egrad2rgrad(z, us)
        rec[f"upper_step_n{n}"] = step.numpy()
        rec[f"upper_retr_n{n}"] = up.retr(z, step).numpy()
        rec[f"upper_projx_n{n}"] = up.projx(z + step).numpy()
        small = -1e-3 * up.egrad2rgrad(z, us)
        rec[f"upper_retr_small_n{n}"] = up.retr(z, small).numpy()
        zb = sm.to_symmetric(cayley_transform(z))
        rec[f"bounded_z_n{n}"] = zb.numpy()
        rec[f"bounded_egrad2rgrad_n{n}"] = bd.egrad2rgrad(zb, u).numpy()
        rec[f"bounded_egrad2rgrad_sym_n{n}"] = bd.egrad2rgrad(zb, us).numpy()
    np.savez_compressed(os.path.join(OUT, "optim.npz"), **rec)
    print("wrote optim.npz")


def main():
    import_reference()
    from sympa.manifolds import UpperHalfManifold, BoundedDomainManifold
    from sympa.manifolds.metrics import MetricType
    from sympa.math import csym_math as sm
    from sympa.math.cayley_transform import cayley_transform, inverse_cayley_transform
    from sympa.math.takagi_factorization import TakagiFactorization

    torch.set_default_dtype(torch.float64)
    os.makedirs(OUT, exist_ok=True)
    mods = (UpperHalfManifold, BoundedDomainManifold, MetricType, sm, cayley_transform)
    if "--all" in sys.argv:                      # the round-1 set (bit-identical when regenerated)
        gen_pairs((2, 3, 4, 5, 6, 10), 1000, mods)
    gen_pairs((1, 7, 8, 9), 2000, mods)          # round 2: the sizes the first set skipped
    gen_optim(mods)
    if "--all" not in sys.argv:
        return

    # building blocks
    g = torch.Generator().manual_seed(7)
    blocks = {}
    for n in (2, 3, 4):
        z = so.upper_spread(6, n, generator=g)
        w = cayley_transform(z)
        blocks[f"z_n{n}"] = z.numpy()
        blocks[f"cayley_n{n}"] = w.numpy()
        blocks[f"icayley_n{n}"] = inverse_cayley_transform(sm.to_symmetric(w)).numpy()
        blocks[f"inverse_n{n}"] = sm.inverse(z).numpy()
        tv = TakagiFactorization(n, return_eigenvectors=False).factorize(sm.to_symmetric(w))
        blocks[f"takagi_n{n}"] = tv.numpy()
        tv2, s = TakagiFactorization(n, return_eigenvectors=True).factorize(sm.to_symmetric(w))
        blocks[f"takagi_vec_n{n}"] = s.numpy()
        blocks[f"msqrt_n{n}"] = sm.matrix_sqrt(sm.imag(z)).numpy()
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **blocks)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

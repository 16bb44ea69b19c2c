"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference); the
output files are committed so the GPU box never needs the reference tree.

    python oracle/gen_golden.py            # rewrites tests/golden/

What is stored, per (manifold kind, n, input regime):
  z1, z2              inputs (b, 2, n, n) float64
  vvd                 the vector-valued distance the reference feeds to its metric (b, n)
  dist_<metric>       manifold.dist for riem / fone / finf / fmin / wsum
  g1_<metric>, g2_<metric>   autograd gradients of sum(go * dist) wrt z1 / z2 (raw, unsymmetrised)
  gw_wsum, wsum_w     wsum weights used and their gradient
  go                  the upstream gradient used
plus building blocks (cayley, inverse cayley, complex inverse, Takagi values) in blocks.npz.
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


def main():
    import_reference()
    from sympa.manifolds import UpperHalfManifold, BoundedDomainManifold
    from sympa.manifolds.metrics import MetricType
    from sympa.math import csym_math as sm
    from sympa.math.cayley_transform import cayley_transform, inverse_cayley_transform
    from sympa.math.takagi_factorization import TakagiFactorization

    torch.set_default_dtype(torch.float64)
    os.makedirs(OUT, exist_ok=True)
    b = 12
    seed = 1000
    for kind in ("upper", "bounded"):
        cls = UpperHalfManifold if kind == "upper" else BoundedDomainManifold
        for n in (2, 3, 4, 5, 6, 10):
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
                        w = torch.linspace(-0.4, 1.3, n).reshape(1, n)
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

"""Riemannian SGD as the reference uses it (geoopt.optim.RiemannianSGD, train.py:66-71: no momentum,
stabilize=None):   g += wd * p;  r = egrad2rgrad(p, g);  p <- retr(p, -lr * r) = projx(p - lr * r).

geoopt is un-vendored (README.md:40) - restated from its documented update, PARITY UNPINNED.

`sparse_rows=True` applies the update only to the table rows whose gradient is non-zero.  That is
exactly equivalent to the dense step: an untouched row has zero gradient, and projx leaves points that
are already on the manifold bit-for-bit unchanged (masks at csym_math.py:275-278 and
bounded_domain.py:79-84) - but it turns the O(N) eigendecompositions of projx into O(batch)
(605 ms of a 636 ms step at N = 100 000)."""
import torch


class RiemannianSGD(torch.optim.Optimizer):
    def __init__(self, params, lr, weight_decay=0.0, sparse_rows=False, fused=False):
        """fused=True: CUDA tables on the upper / bounded / spd manifolds are updated by the one-launch kernel
        behind sympa_rsgd_step (rows with a zero gradient untouched); everything else takes the torch
        path below."""
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay, sparse_rows=sparse_rows, fused=fused))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            lr, wd = group["lr"], group["weight_decay"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                manifold = getattr(p, "manifold", None)
                g = p.grad
                if manifold is None:        # Euclidean parameters: scale, wsum weights
                    if wd:
                        g = g + wd * p
                    p.add_(g, alpha=-lr)
                    continue
                kind = getattr(manifold, "kind", None)
                if group["fused"] and wd == 0.0 and p.is_cuda and kind in ("upper", "bounded", "spd") and p.dtype == torch.float64:
                    from . import ops
                    counter = None
                    if kind in ("upper", "bounded"):
                        counter = getattr(manifold, "_projected_counter", None)
                        if counter is None or counter.device != p.device:
                            counter = torch.zeros(1, dtype=torch.int64, device=p.device)
                            manifold._projected_counter = counter
                    ops.rsgd_step(kind, p.data, g, lr, projected=counter)
                    continue
                if group["sparse_rows"] and wd == 0.0 and p.dim() >= 3:
                    rows = torch.nonzero(g.reshape(g.shape[0], -1).abs().amax(dim=1) > 0).reshape(-1)
                    if rows.numel() == 0:
                        continue
                    pr, gr = p[rows], g[rows]
                    p[rows] = manifold.retr(pr, -lr * manifold.egrad2rgrad(pr, gr))
                    continue
                if wd:
                    g = g + wd * p
                p.copy_(manifold.retr(p, -lr * manifold.egrad2rgrad(p, g)))
        return loss

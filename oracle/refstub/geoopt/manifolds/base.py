import torch


class Manifold(torch.nn.Module):
    __scaling__ = {}
    name = "stub"
    ndim = 0
    reversible = False

    def __init__(self, **kwargs):
        super().__init__()

    def check_point_on_manifold(self, x, *, explain=False, atol=1e-5, rtol=1e-5):
        ok, reason = self._check_shape(x.shape, "x")
        if ok:
            ok, reason = self._check_point_on_manifold(x, atol=atol, rtol=rtol)
        if explain:
            return ok, reason
        return ok

"""Mirror of sympa/losses.py.  On CUDA float64 vectors the loss and its backward are one kernel each
(sympa_distortion_loss_forward / _backward); any other input takes the reference's torch expression."""
import torch

from . import _lib


class _DistortionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, graph_distances, manifold_distances):
        g = graph_distances.detach().reshape(-1).contiguous()
        d = manifold_distances.detach().reshape(-1).contiguous()
        loss = torch.zeros((), dtype=torch.float64, device=d.device)
        with torch.cuda.device(d.device):
            _lib.check(_lib.load().sympa_distortion_loss_forward(d.numel(), g.data_ptr(), d.data_ptr(), loss.data_ptr(),
                                                                 torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(g, d)
        ctx.dshape = manifold_distances.shape
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        g, d = ctx.saved_tensors
        grad_loss = grad_loss.detach().reshape(1).contiguous()
        grad_d = torch.empty_like(d)
        with torch.cuda.device(d.device):
            _lib.check(_lib.load().sympa_distortion_loss_backward(d.numel(), g.data_ptr(), d.data_ptr(), grad_loss.data_ptr(),
                                                                  grad_d.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return None, grad_d.reshape(ctx.dshape)


class AverageDistortionLoss:
    """sum | (d_manifold / d_graph)^2 - 1 |   (sympa/losses.py:10-19)"""

    def calculate_loss(self, graph_distances, manifold_distances):
        if (manifold_distances.is_cuda and graph_distances.is_cuda and manifold_distances.dtype == torch.float64
                and graph_distances.dtype == torch.float64 and graph_distances.numel() == manifold_distances.numel()
                and not graph_distances.requires_grad):
            return _DistortionLossFn.apply(graph_distances, manifold_distances)
        loss = torch.pow(manifold_distances / graph_distances, 2)
        loss = torch.abs(loss - 1)
        return loss.sum()

"""Mirror of sympa/losses.py."""
import torch


class AverageDistortionLoss:
    """sum | (d_manifold / d_graph)^2 - 1 |   (sympa/losses.py:10-19)"""

    def calculate_loss(self, graph_distances, manifold_distances):
        loss = torch.pow(manifold_distances / graph_distances, 2)
        loss = torch.abs(loss - 1)
        return loss.sum()

"""Synthetic graphs of the BASELINE configs as (src, dst, distance) triplets, the format
preprocess.py:118-126 writes (all pairs i < j with their shortest-path distance).  Grid and balanced
tree distances are computed directly (no networkx / networkit needed)."""
import itertools

import torch


def grid_triplets(side, dims=2):
    """`side`^dims grid graph (preprocess.py: nx.grid_graph); distance = Manhattan distance."""
    coords = torch.tensor(list(itertools.product(range(side), repeat=dims)), dtype=torch.int64)
    n = coords.shape[0]
    i, j = torch.triu_indices(n, n, offset=1)
    dist = (coords[i] - coords[j]).abs().sum(-1)
    return torch.stack((i, j), 1).contiguous(), dist.double(), n


def balanced_tree_triplets(branching, height):
    """nx.balanced_tree(branching, height): node k has parent (k - 1) // branching."""
    n = (branching ** (height + 1) - 1) // (branching - 1)
    depth = torch.zeros(n, dtype=torch.int64)
    parent = torch.zeros(n, dtype=torch.int64)
    for k in range(1, n):
        parent[k] = (k - 1) // branching
        depth[k] = depth[parent[k]] + 1
    i, j = torch.triu_indices(n, n, offset=1)
    a, b = i.clone(), j.clone()
    d = torch.zeros_like(a)
    # climb the deeper node until both meet
    for _ in range(2 * height + 1):
        deeper_a = depth[a] > depth[b]
        deeper_b = depth[b] > depth[a]
        same = (~deeper_a) & (~deeper_b) & (a != b)
        d += deeper_a.long() + deeper_b.long() + 2 * same.long()
        a = torch.where(deeper_a | same, parent[a], a)
        b = torch.where(deeper_b | same, parent[b], b)
    return torch.stack((i, j), 1).contiguous(), d.double(), n

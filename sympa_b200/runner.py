"""The training step / epoch of sympa/runner.py:90-122 on the fused path (host logic only; the
tensorboard / checkpoint / evaluation plumbing of the reference Runner is out of scope)."""
import torch
from torch.nn.utils import clip_grad_norm_

from . import distributed as sd
from . import ops
from .losses import AverageDistortionLoss


def train_epoch(model, optimizer, src_dst_ids, graph_distances, batch_size, max_grad_norm=50.0, grad_accum_steps=1,
                world_size=1, rank=0, epoch=0, shuffle=True, sync_stats=True):
    """One pass over the training triplets (runner.py:90-122).

    src_dst_ids (T, 2) int64 and graph_distances (T,) live on the device (train.py:95-96).  Each rank
    takes the DistributedSampler shard of the triplets (train.py:105-110) in batches of
    batch_size // world_size; gradients of all parameters are averaged over the ranks before clipping
    (what DistributedDataParallel does, train.py:59) - the one collective of the path.
    Returns the mean loss per step (as runner.py:122).  sync_stats=False skips the per-step
    `.item()` host synchronisations of runner.py:108-110 and reads the loss once at the end.
    """
    loss_fn = AverageDistortionLoss()
    dev = src_dst_ids.device
    order = sd.shard_indices(src_dst_ids.shape[0], rank, world_size, epoch=epoch, shuffle=shuffle).to(dev)
    per_rank = max(batch_size // world_size, 1)
    ok, point, reason = model.check_all_points()        # runner.py:91
    if not ok:
        raise AssertionError(f"Point outside manifold. Reason: {reason}\n{point}")
    model.train()
    optimizer.zero_grad(set_to_none=False)
    total = torch.zeros((), dtype=torch.float64, device=dev)
    steps = 0
    params = [p for p in model.parameters() if p.requires_grad]
    for step, start in enumerate(range(0, order.numel(), per_rank)):
        sel = order[start:start + per_rank]
        dist_m = model(src_dst_ids[sel])
        loss = loss_fn.calculate_loss(graph_distances[sel], dist_m) / grad_accum_steps
        loss.backward()
        if sync_stats:
            total += loss.item()                         # runner.py:108 (host sync, as the reference)
        else:
            total += loss.detach()
        steps += 1
        if (step + 1) % grad_accum_steps == 0:
            if world_size > 1:
                sd.allreduce_gradients([p.grad for p in params], average=True)
            clip_grad_norm_(params, max_grad_norm)       # runner.py:115
            optimizer.step()
            optimizer.zero_grad(set_to_none=False)
    ops.check_status(dev)       # the reference's per-call asserts, once per epoch
    return float(total.item()) / max(steps, 1)

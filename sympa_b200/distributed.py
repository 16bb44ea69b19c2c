"""Data-parallel step over the pairs of one batch: one process per GPU, pairs sharded the way the
reference's DistributedSampler does (train.py:105-110), the embedding table replicated, and ONE
collective - the all-reduce of the dense table gradient that DistributedDataParallel performs in the
reference (train.py:59; averaged, with lr scaled by the number of processes at train.py:136)."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """env:// rendezvous (train.py:133); NCCL on GPUs, gloo on CPU.  Selects cuda:LOCAL_RANK - the
    reference pins every rank to cuda:0 (sympa/config.py:11-14)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://")
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_indices(num_items, rank, world, epoch=0, shuffle=True, seed=0, drop_last=False):
    """Same partition as torch.utils.data.DistributedSampler(num_replicas=world, rank=rank)
    (train.py:108): a seeded permutation, padded by wrapping so every rank gets ceil(T / world)
    items, then strided rank::world."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(num_items, generator=g)
    else:
        order = torch.arange(num_items)
    if drop_last and num_items % world:
        per = num_items // world
        order = order[: per * world]
    else:
        per = -(-num_items // world)
        pad = per * world - num_items
        if pad:
            reps = -(-pad // max(num_items, 1))
            order = torch.cat([order, order.repeat(reps)[:pad]])
    return order[rank::world]


def allreduce_gradients(tensors, average=True):
    """The single collective of the path: sum (or average, as DDP does) the dense gradients over all
    ranks.  On GPUs this is NCCL over NVLink on the current stream; tensors are reduced in place."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensors
    world = dist.get_world_size()
    # NCCL averages inside the collective (no extra read + write pass over the 268 MB table gradient);
    # gloo (CPU tests) has no AVG: sum, then divide
    fused_avg = average and dist.get_backend() == "nccl"
    for t in tensors:
        if t is None:
            continue
        dist.all_reduce(t, op=dist.ReduceOp.AVG if fused_avg else dist.ReduceOp.SUM)
        if average and not fused_avg:
            t.div_(world)
    return tensors

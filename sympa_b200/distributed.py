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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index=None):
    """Run this process on the CPUs of the NUMA node its GPU hangs off (Linux sysfs), so that the pinned host
    buffers it allocates afterwards are first-touched there: with one process per GPU on a two-socket box the
    ranks on the far socket otherwise read their batches across the inter-socket link.  Returns
    {"node": k, "cpus": count} or None when the topology cannot be read (nothing is changed then)."""
    try:
        if not torch.cuda.is_available():
            return None
        index = torch.cuda.current_device() if device_index is None else int(device_index)
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        cpus &= os.sched_getaffinity(0)        # never widen what the launcher (cgroup, taskset) allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError):
        return None


def shard_indices(num_items, rank, world, epoch=0, shuffle=True, seed=0, drop_last=False, device=None):
    """Same partition as torch.utils.data.DistributedSampler(num_replicas=world, rank=rank)
    (train.py:108): a seeded permutation, padded by wrapping so every rank gets ceil(T / world)
    items, then strided rank::world.
    device: draw the permutation ON that device (same seed -> the same permutation on every rank's GPU, but not
    the CPU generator's, i.e. not DistributedSampler's order).  The CPU permutation of the 12.5 M pairs of
    BASELINE config 3 costs ~0.2 s per epoch - as much as the epoch's whole GPU work."""
    if shuffle and device is not None and torch.device(device).type == "cuda":
        g = torch.Generator(device=device)
        g.manual_seed(seed + epoch)
        order = torch.randperm(num_items, generator=g, device=device)
    elif shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(num_items, generator=g)
    else:
        order = torch.arange(num_items, device=device)
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


def row_shard(num_rows, rank, world):
    """Rows [begin, end) of the table owned by `rank`: equal blocks of ceil(N / world) rows (the last
    block may be shorter or empty)."""
    per = -(-num_rows // world)
    begin = min(rank * per, num_rows)
    return begin, min(begin + per, num_rows), per


def sharded_rsgd_step(param, lr, update_rows, average=True):
    """Owner-computes variant of `allreduce_gradients` + optimizer step for the embedding table
    (the reference all-reduces the dense gradient through DDP, train.py:59, and
    every rank then runs the O(N) Riemannian update and projection on the whole table):

      1. reduce-scatter the dense table gradient by row owner (each rank receives the reduced gradient
         of its ceil(N / world) rows);
      2. the owner updates its rows:  update_rows(rows_view, grad_rows, lr)  in place - the fused
         `sympa_b200.ops.rsgd_step` on a GPU, any Riemannian update otherwise;
      3. all-gather the updated rows into every rank's table.

    Bytes on the wire equal those of the all-reduce (reduce-scatter + all-gather of the same size); what
    changes is that the update and projection of the table is done once, split over the ranks, instead
    of world times.  NCCL runs both collectives on the current stream; backends without
    reduce_scatter_tensor (gloo, CPU tests) take an all-reduce and keep their slice.  Every rank ends
    with the same table, bit-identical to the replicated step given the same reduced gradient.
    Returns the (begin, end) rows this rank owns."""
    table, grad = param.data, param.grad
    if not dist.is_initialized() or dist.get_world_size() == 1:
        update_rows(table, grad, lr)
        return 0, table.shape[0]
    world, rank = dist.get_world_size(), dist.get_rank()
    n_rows = table.shape[0]
    begin, end, per = row_shard(n_rows, rank, world)
    row_elems = table[0].numel()
    padded = per * world
    nccl = dist.get_backend() == "nccl"
    if nccl:
        src = grad.reshape(n_rows, row_elems)
        if padded != n_rows:   # reduce_scatter_tensor wants world equal blocks
            src = torch.cat([src, src.new_zeros(padded - n_rows, row_elems)])
        mine = torch.empty(per, row_elems, dtype=grad.dtype, device=grad.device)
        dist.reduce_scatter_tensor(mine, src, op=dist.ReduceOp.AVG if average else dist.ReduceOp.SUM)
        mine = mine[: end - begin]
    else:
        full = grad.clone()
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
        if average:
            full.div_(world)
        mine = full.reshape(n_rows, row_elems)[begin:end]
    if end > begin:
        update_rows(table[begin:end], mine.reshape((end - begin,) + tuple(table.shape[1:])), lr)
    # all-gather the updated rows
    block = torch.zeros(per, row_elems, dtype=table.dtype, device=table.device)
    if end > begin:
        block[: end - begin] = table[begin:end].reshape(end - begin, row_elems)
    if nccl:
        gathered = torch.empty(padded, row_elems, dtype=table.dtype, device=table.device)
        dist.all_gather_into_tensor(gathered, block)
    else:
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block)
        gathered = torch.cat(parts)
    table.copy_(gathered[:n_rows].reshape(table.shape))
    return begin, end

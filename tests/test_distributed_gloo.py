"""World-size-2 gloo test of the host-side data-parallel logic (CPU): the pair shards of
`shard_indices` partition the batch the way DistributedSampler does, and summing the per-shard table
gradients with `allreduce_gradients` reproduces the full-batch gradient.  The per-shard gradient
itself comes from the oracle here (no GPU in this test); the CUDA kernels are exercised by -m gpu."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siegel_oracle as so
    from sympa_b200 import distributed as sd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = sd.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(0)
    rows, pairs, n = 23, 101, 3
    table = so.upper_spread(rows, n, generator=g, scale=0.3)
    src = torch.randint(0, rows, (pairs,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (pairs,), generator=g)) % rows
    gd = torch.randint(1, 20, (pairs,), generator=g).double()

    def grad_of(sel):
        t = table.clone().requires_grad_(True)
        d = so.upper_dist(t[src[sel]], t[dst[sel]], "fone")
        so.distortion_loss(gd[sel], d).backward()
        return t.grad

    shard = sd.shard_indices(pairs, rank, world, epoch=3, shuffle=True, seed=0, drop_last=True)
    # same partition as the reference's DistributedSampler (train.py:108)
    from torch.utils.data.distributed import DistributedSampler
    ds = DistributedSampler(range(pairs), num_replicas=world, rank=rank, shuffle=True, seed=0, drop_last=True)
    ds.set_epoch(3)
    assert list(ds) == shard.tolist()
    local = grad_of(shard)
    (total,) = sd.allreduce_gradients([local.clone()], average=False)
    (avg,) = sd.allreduce_gradients([local.clone()], average=True)
    # every rank must hold the gradient of the union of the shards
    all_sel = torch.cat([sd.shard_indices(pairs, k, world, epoch=3, shuffle=True, seed=0, drop_last=True)
                         for k in range(world)])
    assert len(set(all_sel.tolist())) == len(all_sel) == (pairs // world) * world
    full = grad_of(all_sel)
    ok = torch.allclose(total, full, rtol=1e-10, atol=1e-12) and torch.allclose(avg * world, full, rtol=1e-10, atol=1e-12)
    # padded (drop_last=False) variant matches DistributedSampler too
    ds2 = DistributedSampler(range(pairs), num_replicas=world, rank=rank, shuffle=False)
    ok = ok and list(ds2) == sd.shard_indices(pairs, rank, world, shuffle=False).tolist()
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    world = 2
    port = 29600 + os.getpid() % 300
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _worker_sharded(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siegel_oracle as so
    from sympa_b200 import UpperHalfManifold
    from sympa_b200 import distributed as sd
    from sympa_b200.embeddings import ManifoldParameter

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sd.init_process_group(backend="gloo")
    g = torch.Generator().manual_seed(1)
    rows, n, lr = 23, 3, 0.3                      # 23 rows over 2 ranks: blocks of 12 and 11
    table = so.upper_spread(rows, n, generator=g, scale=0.3)
    grads = [so.sym(torch.randn(rows, 2, n, n, dtype=torch.float64, generator=g)) for _ in range(world)]
    man = UpperHalfManifold(dims=n)

    def update_rows(t, gr, step):              # the Riemannian update of the host optimizer, in place
        t.copy_(man.retr(t, -step * man.egrad2rgrad(t, gr)))

    # replicated step: all-reduce (average) then update the whole table on every rank
    rep = table.clone()
    (avg,) = sd.allreduce_gradients([grads[rank].clone()], average=True)
    update_rows(rep, avg, lr)
    # owner-computes step
    p = ManifoldParameter(table.clone(), manifold=man)
    p.grad = grads[rank].clone()
    begin, end = sd.sharded_rsgd_step(p, lr, update_rows, average=True)
    ok = (begin, end) == sd.row_shard(rows, rank, world)[:2]
    ok = ok and torch.equal(p.data, rep)
    ok = ok and torch.allclose(rep, so.rsgd_step("upper", table, sum(grads) / world, lr), rtol=1e-10, atol=1e-12)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_optimizer_step_equals_replicated_step():
    world = 2
    port = 29900 + os.getpid() % 90
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker_sharded, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}

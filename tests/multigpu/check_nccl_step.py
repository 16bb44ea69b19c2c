"""Multi-GPU (NCCL) check of the data-parallel step; run under torchrun on >= 2 GPUs, e.g.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/check_nccl_step.py
(not collected by pytest: the -m gpu suite runs on one GPU; the same logic is covered on CPU with gloo in
tests/test_distributed_gloo.py).  Checks, on every rank:
  * pairs sharded over the ranks + the averaged NCCL all-reduce of the table gradient == the full-batch
    gradient of one GPU / world; the all-reduce issued inside the backward on the packed gradient table
    (dist_from_table(..., sync_grad=True)) gives the same gradient;
  * the same through a TableGradAccumulator (three calls, one all-reduce in finish(sync_grad=True));
  * the owner-computes optimizer step (reduce-scatter, fused sympa_rsgd_step on the owned rows,
    all-gather) leaves every rank with the table of the replicated step."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import siegel_oracle as so
    from sympa_b200 import MetricType, UpperHalfManifold, ops
    from sympa_b200 import distributed as sd
    from sympa_b200.embeddings import ManifoldParameter

    rank, world, local = sd.init_process_group()
    dev = torch.device("cuda", local)
    g = torch.Generator().manual_seed(0)
    rows, pairs, n = 1001, 40000, 4
    table = so.upper_spread(rows, n, generator=g, scale=0.3).to(dev)
    src = torch.randint(0, rows, (pairs,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (pairs,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).to(dev)
    gd = torch.randint(1, 20, (pairs,), generator=g).double().to(dev)
    man = UpperHalfManifold(dims=n, metric=MetricType.from_str("riem")).to(dev)

    def grad_of(sel):
        t = table.clone().requires_grad_(True)
        d = man.dist_from_table(t, idx[sel].contiguous())
        (torch.abs((d / gd[sel]) ** 2 - 1)).sum().backward()
        return t.grad

    shard = sd.shard_indices(pairs, rank, world, shuffle=True, seed=0, drop_last=True).to(dev)
    local_grad = grad_of(shard)
    (avg,) = sd.allreduce_gradients([local_grad.clone()], average=True)
    all_sel = torch.cat([sd.shard_indices(pairs, k, world, shuffle=True, seed=0, drop_last=True) for k in range(world)]).to(dev)
    full = grad_of(all_sel)
    torch.testing.assert_close(avg * world, full, rtol=1e-9, atol=1e-9 * full.abs().max().item())

    # the same all-reduce issued inside the backward, on the packed gradient table (sync_grad=True)
    t = table.clone().requires_grad_(True)
    d = man.dist_from_table(t, idx[shard].contiguous(), sync_grad=True)
    (torch.abs((d / gd[shard]) ** 2 - 1)).sum().backward()
    torch.testing.assert_close(t.grad, avg, rtol=1e-10, atol=1e-12 * full.abs().max().item())

    # bounded domain, by-rows route (inverse Cayley per table row) with the in-backward all-reduce
    from sympa_b200 import BoundedDomainManifold
    btable = so.to_symmetric(so.cayley_transform(table.cpu())).to(dev)
    bman = BoundedDomainManifold(dims=n, metric=MetricType.from_str("fone")).to(dev)
    grads = []
    for sync in (False, True):
        t = btable.clone().requires_grad_(True)
        d = bman.dist_from_table(t, idx[shard].contiguous(), sync_grad=sync)
        (torch.abs((d / gd[shard]) ** 2 - 1)).sum().backward()
        if not sync:
            sd.allreduce_gradients([t.grad], average=True)
        grads.append(t.grad.clone())
    torch.testing.assert_close(grads[1], grads[0], rtol=1e-10, atol=1e-12 * grads[0].abs().max().item())

    # gradient accumulation: the shard in three calls into one packed table (scatter on the accumulator's side stream),
    # one all-reduce of the packed table in finish() - upper and bounded (by rows)
    for m, tb, want in ((man, table, avg), (bman, btable, grads[0])):
        t = tb.clone().requires_grad_(True)
        acc = m.table_grad_accumulator(t)
        cuts = [0, len(shard) // 2, len(shard) // 2 + 1, len(shard)]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            sel = shard[lo:hi]
            d = m.dist_from_table(t, idx[sel].contiguous(), accumulator=acc)
            (torch.abs((d / gd[sel]) ** 2 - 1)).sum().backward()
        acc.finish(sync_grad=True)
        torch.testing.assert_close(t.grad, want, rtol=1e-10, atol=1e-12 * want.abs().max().item())

    lr = 0.05
    rep = table.clone()
    ops.rsgd_step("upper", rep, avg, lr)
    p = ManifoldParameter(table.clone(), manifold=man)
    p.grad = local_grad.clone()
    begin, end = sd.sharded_rsgd_step(p, lr, lambda t, gr, step: ops.rsgd_step("upper", t, gr.contiguous(), step))
    assert (begin, end) == sd.row_shard(rows, rank, world)[:2]
    # the reduce-scatter and the all-reduce may sum in different orders: equal up to rounding
    torch.testing.assert_close(p.data, rep, rtol=1e-12, atol=1e-13)
    ops.check_status(dev)
    dist.barrier()
    if rank == 0:
        print(f"nccl step check ok on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

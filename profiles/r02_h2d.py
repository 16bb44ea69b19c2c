"""host -> device copy bandwidth per rank with all ranks copying at once (torchrun), and rank 0 alone:
what the e2e leg's PairFeeder can get out of the host at N GPUs."""
import os, sys, time, torch
import torch.distributed as dist
sys.path.insert(0, ".")
from sympa_b200 import distributed as sd
rank, world, local = sd.init_process_group()
dev = torch.device("cuda", local)
nbytes = 75 * (1 << 20)
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
def run(k=40):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return nbytes * k / (e0.elapsed_time(e1) * 1e-3) / 1e9
run(5)
if world > 1: dist.barrier()
together = run()
t = torch.tensor([together], dtype=torch.float64, device=dev)
if world > 1:
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    dist.barrier()
alone = run() if rank == 0 else 0.0
if world > 1: dist.barrier()
if rank == 0:
    vals = [round(float(x), 1) for x in lst] if world > 1 else [round(together, 1)]
    print("H2D GB/s per rank, all %d ranks at once: %s  sum %.1f;  rank 0 alone: %.1f" % (world, vals, sum(vals), alone))
if world > 1: dist.destroy_process_group()

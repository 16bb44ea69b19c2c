import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import siegel_oracle as so
import sympa_b200 as sb
n, rows, b = 4, 3000, 200_000
g = torch.Generator().manual_seed(5)
table = so.upper_spread(rows, n, generator=g, scale=0.3).cuda()
src = torch.randint(0, rows, (b,), generator=g)
dst = (src + 1 + torch.randint(0, rows - 1, (b,), generator=g)) % rows
idx = torch.stack((src, dst), 1).cuda()
gd = torch.randint(1, 20, (b,), generator=g).double().cuda()
man = sb.UpperHalfManifold(dims=n, metric=sb.MetricType.from_str("riem")).cuda()
for sms in (0, 48):
  for use_acc in (True, False):
    t = table.clone().requires_grad_(True)
    acc = man.table_grad_accumulator(t, scatter_sms=sms) if use_acc else None
    for slow in (False, True):
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        if slow:
            torch.cuda._sleep(int(2e9))      # the device is ~1 s behind the host
        peaks = []
        for step in range(8):
            t.grad = None
            for lo, hi in ((0, 70_000), (70_000, 140_000), (140_000, b)):
                so.distortion_loss(gd[lo:hi], man.dist_from_table(t, idx[lo:hi], accumulator=acc)).backward()
            if acc is not None:
                acc.finish()
            peaks.append((torch.cuda.max_memory_allocated() - base) >> 20)
        torch.cuda.synchronize()
        print("sms", sms, "acc", use_acc, "device behind" if slow else "device keeps up", peaks)

"""time of the packed-table scatter alone as a function of the SM partition (sympa_table_grad_scatter_add max_sms)"""
import sys, torch
sys.path.insert(0, ".")
from sympa_b200 import _lib, ops
lib = _lib.load()
n, rows, b = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 1 << 20, 1 << 23
per_s = n * (n + 1)
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
idx = torch.randint(0, rows, (b, 2), device=dev, generator=g)
gd = torch.rand(b, device=dev, dtype=torch.float64)
saved = torch.rand(2 * b * per_s, device=dev, dtype=torch.float64)
nbytes = lib.sympa_backward_workspace_bytes(0, n, rows)
ws = torch.zeros(nbytes // 8, device=dev, dtype=torch.float64)
tickets = torch.zeros(4, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for sms in (0, 148, 96, 64, 48, 32, 24, 16, 8):
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.sympa_table_grad_scatter_add(0, n, b, gd.data_ptr(), saved.data_ptr(), rows, idx.data_ptr(), ws.data_ptr(), nbytes,
                                                    sms, tickets.data_ptr() if sms else None, st))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[2:])[len(ts[2:]) // 2]
    print("n %d max_sms %4d  scatter %.3f ms   SM-time %.1f SM*ms" % (n, sms, t, t * (sms if sms else 148)))

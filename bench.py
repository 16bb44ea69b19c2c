#!/usr/bin/env python
"""Benchmark of the Siegel pair-distance hot path (BASELINE.json metric: pairs/s forward+backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 4] [--kind upper]

Workload (config.workload): BASELINE.json configs[4] 'distance microbench' - a table of 2^20 random
points, B index pairs per GPU per step, manifold `--kind` with n x n matrices, metric riem, forward
+ backward through the distortion loss (sympa/losses.py:16-19).  One step = gather -> dist ->
loss -> backward -> scatter-add into the dense table gradient (+ NCCL all-reduce of that gradient
when N > 1).  `value` is measured with everything resident in HBM; `e2e` runs the same step through
the public manifold API with the step's inputs (index pairs, graph distances) copied from pinned
host memory and the loss read back, inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
    # keep NCCL's version banner (printed to STDOUT at VERSION and WARN level) out of the way: rank 0 prints
    # ONE JSON line.  An unrecognised level means "no logging" to NCCL; INFO / TRACE set by the user are kept.
    os.environ["NCCL_DEBUG"] = "NONE"

import torch  # noqa: E402


def default_pairs(n):
    """pairs per GPU per step: BASELINE.json's 64M-pair batch over 8 GPUs (2^23 per GPU) where the saved unit
    gradients fit comfortably, smaller for the large matrix sizes (a step is still >= 25 ms there)."""
    return (1 << 23) if n <= 4 else ((1 << 20) if n <= 6 else (1 << 18))


def algorithmic_bytes_per_pair(kind, n):
    """read two rows, accumulate two gradient rows, two int64 indices, vvd out, dist
    out, upstream gradient in."""
    if kind == "spd":
        return 32 * n * n + 8 * n + 32
    return 64 * n * n + 8 * n + 32


def lean_flops_per_pair(n):
    return 100 * n ** 3   # SURVEY.md 8(d) yardstick F_lean


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(dev):
    """attainable FP64 FMA rate of this GPU in TFLOP/s (sympa_probe_fp64, best of 5, CUDA events)."""
    from sympa_b200 import _lib
    lib = _lib.load()
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    iters = 1 << 16
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = lib.sympa_probe_fp64(iters, out.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        if flops > 0:
            best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def make_table(kind, n, rows, device, seed):
    """'spread' regime of SURVEY.md 8(d), scale 0.3: X = sym(0.3 N(0,1)), Y = C C^T + 0.5 I."""
    g = torch.Generator(device=device).manual_seed(seed)
    c = 0.3 * torch.randn(rows, n, n, dtype=torch.float64, device=device, generator=g)
    y = c @ c.transpose(-1, -2) + 0.5 * torch.eye(n, dtype=torch.float64, device=device)
    y = 0.5 * (y + y.transpose(-1, -2))
    if kind == "spd":
        return y.contiguous()
    x = 0.3 * torch.randn(rows, n, n, dtype=torch.float64, device=device, generator=g)
    x = 0.5 * (x + x.transpose(-1, -2))
    z = torch.stack((x, y), 1).contiguous()
    if kind == "bounded":
        from sympa_b200.manifolds import csym
        out = torch.empty_like(z)
        for s in range(0, rows, 1 << 16):
            out[s:s + (1 << 16)] = csym.to_symmetric(csym.cayley_transform(z[s:s + (1 << 16)]))
        z = out
    return z


def make_pairs(rows, b, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, rows, (b,), device=device, generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (b,), device=device, generator=g)) % rows   # src != dst
    gd = torch.randint(1, 21, (b,), device=device, generator=g).double()
    return torch.stack((src, dst), 1).contiguous(), gd


def distortion_loss(graph_dist, manifold_dist):
    """sympa/losses.py:16-19.  GPU legs: the package's AverageDistortionLoss (same expression, one kernel forward
    and one backward on CUDA float64 vectors); CPU legs (baseline / reference arm): the reference's torch ops."""
    if manifold_dist.is_cuda:
        from sympa_b200.losses import AverageDistortionLoss
        return AverageDistortionLoss().calculate_loss(graph_dist, manifold_dist)
    return torch.abs(torch.pow(manifold_dist / graph_dist, 2) - 1).sum()


def run_ours(args):
    import torch.distributed as dist
    from sympa_b200 import BoundedDomainManifold, MetricType, SymmetricPositiveDefinite, UpperHalfManifold, ops
    from sympa_b200 import distributed as sd

    rank, world, local = sd.init_process_group()
    assert world == args.gpus, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    dev = torch.device("cuda", local)
    kind, n = args.kind, args.n
    rows = args.rows
    b = args.pairs if args.pairs else default_pairs(n)
    if kind == "spd":
        man = SymmetricPositiveDefinite().to(dev)
    else:
        man = {"upper": UpperHalfManifold, "bounded": BoundedDomainManifold}[kind](
            dims=n, metric=MetricType.from_str(args.metric)).to(dev)
    table = make_table(kind, n, rows, dev, seed=1).requires_grad_(True)     # replicated on every rank
    idx, gd = make_pairs(rows, b, dev, seed=100 + rank)                       # each rank its own shard
    scale = 1.0

    def step_resident():
        table.grad = None
        d = man.dist_from_table(table, idx, sync_grad=world > 1)   # the all-reduce (average) runs inside the backward,
        loss = distortion_loss(gd, d * scale)                       # on the packed gradient table
        loss.backward()
        return loss

    # host-resident inputs for the e2e leg
    idx_h = idx.to(torch.int32).cpu().pin_memory()    # node ids as int32 on the host, widened on the device by the feeder
    gd_h = gd.cpu().pin_memory()
    loss_h = torch.empty(1, dtype=torch.float64).pin_memory()

    from sympa_b200.feeder import PairFeeder
    feeder = PairFeeder(dev)

    def step_e2e(more):
        """One step through the public API with HOST inputs: this step's batch was submitted (pinned ->
        device, side stream) before the previous step's compute was enqueued, the next one is submitted
        here, the loss is read back and the host waits for it."""
        idx_b, gd_b = feeder.next()
        if more:
            feeder.submit(idx_h, gd_h)
        table.grad = None
        d = man.dist_from_table(table, idx_b, sync_grad=world > 1)
        loss = distortion_loss(gd_b, d * scale)
        loss.backward()
        feeder.done()
        loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_h[0])

    def run_e2e(steps):
        feeder.submit(idx_h, gd_h)            # first batch: its copy is exposed, and inside the timed region
        for s in range(steps):
            step_e2e(s + 1 < steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, fwd_events=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ops.check_status(dev)

    # dominant kernel (forward + unit gradients) timed alone inside the same loop shape
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = world * b * args.steps / (ms_total * 1e-3)

    # per-kernel durations with CUDA events on the launching (current) stream
    fwd_ms, bwd_ms = [], []
    for _ in range(min(args.steps, 10)):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.no_grad():
            ev[0].record()
            dd, vv, saved = ops.forward_raw(kind, args.metric if kind != "spd" else "riem", table=table.detach(), idx=idx,
                                            wsum_w=None, want_grad=True)
            ev[1].record()
        g = torch.ones_like(dd)
        gt = torch.empty_like(table)
        from sympa_b200 import _lib
        lib = _lib.load()
        ws, ws_bytes = ops.backward_workspace_for(kind, n, rows, dev)
        ev[2].record()   # the backward of the table path as the autograd Function issues it (workspace, overwrite)
        _lib.check(lib.sympa_dist_backward_table(_lib.KIND[kind], n, _lib.METRIC["riem"], b, g.data_ptr(), saved.data_ptr(),
                                                 gt.data_ptr(), rows, idx.data_ptr(), None, None, None,
                                                 None if ws is None else ws.data_ptr(), ws_bytes, 1,
                                                 torch.cuda.current_stream().cuda_stream))
        ev[3].record()
        torch.cuda.synchronize()
        fwd_ms.append(ev[0].elapsed_time(ev[1]))
        bwd_ms.append(ev[2].elapsed_time(ev[3]))
        del saved, dd, vv, gt
    fwd = statistics.mean(fwd_ms)
    bwd = statistics.mean(bwd_ms)

    # fused single-launch step (extension: on-device loss), for comparison
    gt = torch.zeros_like(table)
    loss_out = torch.zeros(1, dtype=torch.float64, device=dev)

    def step_fused():
        gt.zero_()
        loss_out.zero_()
        ops.distortion_step(kind, args.metric if kind != "spd" else "riem", table.detach(), idx, gd, scale, gt,
                            loss_out=loss_out)
    for _ in range(3):
        step_fused()
    ms_fused = timed(step_fused, args.steps) / args.steps
    del gt

    # e2e
    run_e2e(3)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    run_e2e(args.steps)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e_value = world * b * args.steps / (ms_e2e * 1e-3)
    ops.check_status(dev)

    peaks, peak_kind = load_peaks()
    fp64_peak = measure_fp64_peak(dev)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(f"{kind}_n{n}_fwdsave")
        if rec:
            traffic = rec["bytes_per_pair"] * b     # ncu dram bytes per pair x pairs of one launch
    scatter = None
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(f"{kind}_n{n}_backward")
        if rec:   # the backward of a step: its DRAM traffic (ncu) over its live-timed duration
            sbytes = rec["scatter_bytes_per_pair"] * b + rec["expand_bytes_per_row"] * rows
            scatter = {"kernel": rec["kernel"], "bound": "hbm", "traffic": sbytes,
                       "achieved": round(sbytes / (bwd * 1e-3) / 1e9, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": round(sbytes / (bwd * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                       "note": "DRAM bytes measured by ncu at 2^20 pairs / 2^20 rows (per pair: random read-modify-write of "
                               "packed table-gradient rows that do not fit L2 + the packed saved state; per row: memset, "
                               "read and dense write of the expansion), scaled to this run's pairs and rows - not "
                               "algorithmic bytes"}
    abytes = algorithmic_bytes_per_pair(kind, n) * b
    achieved = abytes / (fwd * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": f"pair_kernel<{n},{kind},fwd+unit-grad>", "achieved": round(achieved, 2),
        "peak": peaks["hbm_gbs"], "peak_source": peak_kind, "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
        "traffic": traffic, "traffic_source": "ncu --set full capture, see profiles/r01_traffic.json",
        "algorithmic_bytes": abytes, "kernel_ms": round(fwd, 4), "scatter_kernel_ms": round(bwd, 4),
        "fp64": {"achieved_lean_tflops": round(lean_flops_per_pair(n) * b / (fwd * 1e-3) / 1e12, 3),
                 "peak_measured_tflops": round(fp64_peak, 2), "peak_nominal_tflops": 37.2,
                 "frac_of_measured": round(lean_flops_per_pair(n) * b / (fwd * 1e-3) / 1e12 / max(fp64_peak, 1e-9), 4),
                 "how": "algorithmic flops 100 n^3 per pair (SURVEY 8(d) F_lean) / kernel time; peak = sympa_probe_fp64 "
                        "(pure DFMA chains) timed with CUDA events in this run"},
        "note": "n >= 3 is bound by the FP64 pipe (Jacobi sweeps), not HBM: see the fp64 object and profiles/; the "
                "HBM fraction uses algorithmic bytes 64n^2+8n+32 per pair",
    }
    if scatter is not None:
        roofline["scatter"] = scatter
    out = {
        "metric": "Siegel dist pairs/s fwd+bwd", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"distance microbench: {kind} n={n} metric={args.metric}, table 2^{rows.bit_length() - 1} rows, "
                               f"{b} pairs/GPU/step, fwd+bwd through AverageDistortionLoss",
                   "pairs_per_gpu_per_step": b, "rows": rows, "n": n, "kind": kind, "metric": args.metric,
                   "l2": "table + saved unit gradients + indices exceed L2 every step (no explicit flush needed)",
                   "parallelism": f"dp{world} pairs sharded, table replicated, one NCCL all-reduce (average) of the packed "
                                  "table gradient inside the backward"},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(idx_h.numel() * idx_h.element_size() + gd_h.numel() * gd_h.element_size()),
                "d2h_bytes_per_step": 8,
                "how": "manifold.dist_from_table + AverageDistortionLoss + backward per step; every step's index pairs (int32 "
                       "on the host, widened on the device) and graph distances (float64) are copied from pinned host memory (sympa_b200.feeder.PairFeeder: the copy of step "
                       "k+1 overlaps the compute of step k on a side stream; the first copy is exposed), the loss is read "
                       "back and the host waits for it every step"},
        # our kernels per step: forward+unit-gradient kernel, then either the direct scatter (1) or the packed scatter +
        # expansion (2) - torch's loss / fill kernels are not counted
        "gpu_launches": (1 + (2 if (ops.backward_workspace_for(kind, n, rows, dev)[1] > 0 and 2 * b >= rows) else 1)) * args.steps,
        "fused_step": {"pairs_per_s": world * b / (ms_fused * 1e-3), "ms_per_step": ms_fused, "launches_per_step": 1},
        "clocks": clocks,
        "roofline": roofline,
    }
    if not args.no_extras:
        out["also"] = extras(args, world, rank, dev)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(kind, n, args.metric, budget_s=args.cpu_budget)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def quick_pairs_per_s(kind, n, metric, pairs, rows, dev, world, steps=5):
    """resident fwd+bwd pairs/s of another configuration (same step shape as the main measurement)."""
    import torch.distributed as dist
    from sympa_b200 import BoundedDomainManifold, MetricType, SymmetricPositiveDefinite, UpperHalfManifold
    from sympa_b200 import distributed as sd
    man = (SymmetricPositiveDefinite() if kind == "spd" else
           {"upper": UpperHalfManifold, "bounded": BoundedDomainManifold}[kind](dims=n, metric=MetricType.from_str(metric))).to(dev)
    table = make_table(kind, n, rows, dev, seed=2).requires_grad_(True)
    idx, gd = make_pairs(rows, pairs, dev, seed=300)

    def step():
        table.grad = None
        d = man.dist_from_table(table, idx, sync_grad=world > 1)
        distortion_loss(gd, d).backward()

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return world * pairs * steps / (ms * 1e-3)


def epoch_seconds(dev, world, rank, epochs=3, fused=True, sync_stats=True, config=1):
    """Seconds per training epoch (runner.py:90-122 semantics, per-step loss.item() kept), batch 2048, RiemannianSGD.
    config 1 (BASELINE configs[0]): grid 20x20 (400 nodes, 79 800 pairs), upper / riem / n=2;
    config 2 (BASELINE configs[1]): balanced tree branching 3 height 5 (364 nodes, 66 066 pairs), bounded / fone / n=3."""
    from types import SimpleNamespace
    from sympa_b200.graphs import balanced_tree_triplets, grid_triplets
    from sympa_b200.model import Model
    from sympa_b200.optim import RiemannianSGD
    from sympa_b200.runner import train_epoch
    torch.manual_seed(0)
    if config == 1:
        idx, gd, nodes = grid_triplets(20, 2)
        args = SimpleNamespace(manifold="upper", metric="riem", dims=2, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                               train_scale=False)
    else:
        idx, gd, nodes = balanced_tree_triplets(3, 5)
        args = SimpleNamespace(manifold="bounded", metric="fone", dims=3, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                               train_scale=False)
    model = Model(args).to(dev)
    opt = RiemannianSGD(model.parameters(), lr=1e-2 * world, fused=fused)
    idx, gd = idx.to(dev), gd.to(dev)
    train_epoch(model, opt, idx, gd, 2048, world_size=world, rank=rank, epoch=0, sync_stats=sync_stats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ep in range(1, epochs + 1):
        loss = train_epoch(model, opt, idx, gd, 2048, world_size=world, rank=rank, epoch=ep, sync_stats=sync_stats)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / epochs, loss


def fused_epoch_seconds(dev, world, rank, epochs=5, config=1):
    """The same epochs on sympa_b200.runner.FusedEpochRunner: two launches per step (fused distortion step, fused
    optimizer row kernel with clipping and zero_grad), the epoch replayed from a CUDA graph on one GPU, the loss
    read back once per epoch."""
    from types import SimpleNamespace
    from sympa_b200.graphs import balanced_tree_triplets, grid_triplets
    from sympa_b200.model import Model
    from sympa_b200.runner import FusedEpochRunner
    torch.manual_seed(0)
    if config == 1:
        idx, gd, nodes = grid_triplets(20, 2)
        args = SimpleNamespace(manifold="upper", metric="riem", dims=2, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                               train_scale=False)
    else:
        idx, gd, nodes = balanced_tree_triplets(3, 5)
        args = SimpleNamespace(manifold="bounded", metric="fone", dims=3, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                               train_scale=False)
    model = Model(args).to(dev)
    runner = FusedEpochRunner(model, 1e-2 * world, idx.to(dev), gd.to(dev), 2048, world_size=world, rank=rank)
    runner.run_epoch(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ep in range(1, epochs + 1):
        loss = runner.run_epoch(ep)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / epochs, loss


def extras(args, world, rank, dev):
    """The other sizes BASELINE.json's metric names (n = 10) and its epoch-time leg, measured briefly."""
    out = {}
    try:
        out["upper_n10_fmin_pairs_per_s"] = quick_pairs_per_s("upper", 10, "fmin", 1 << 18, 1 << 18, dev, world)
        out["upper_n2_riem_pairs_per_s"] = quick_pairs_per_s("upper", 2, "riem", 1 << 22, 1 << 20, dev, world)
        out["spd_n10_pairs_per_s"] = quick_pairs_per_s("spd", 10, "riem", 1 << 18, 1 << 18, dev, world)
        out["bounded_n3_fone_pairs_per_s"] = quick_pairs_per_s("bounded", 3, "fone", 1 << 22, 1 << 20, dev, world)
        sec, loss = epoch_seconds(dev, world, rank, fused=True, sync_stats=True)
        out["train_epoch_sec_config1_grid400_upper_riem_n2_b2048"] = sec
        out["train_epoch_final_loss"] = loss
        sec_h, _ = epoch_seconds(dev, world, rank, fused=False, sync_stats=True)
        out["train_epoch_sec_config1_host_optimizer"] = sec_h
        sec_n, _ = epoch_seconds(dev, world, rank, fused=True, sync_stats=False)
        out["train_epoch_sec_config1_no_per_step_item_sync"] = sec_n
        for cfg, key in ((1, "train_epoch_sec_config1_fused_graph"), (2, "train_epoch_sec_config2_fused_graph")):
            sec_g, loss_g = fused_epoch_seconds(dev, world, rank, config=cfg)
            out[key] = sec_g
            out[key.replace("_sec_", "_final_loss_")] = loss_g
        sec2, loss2 = epoch_seconds(dev, world, rank, fused=True, sync_stats=True, config=2)
        out["train_epoch_sec_config2_tree_b3h5_bounded_fone_n3_b2048"] = sec2
        out["train_epoch_config2_final_loss"] = loss2
    except Exception as e:  # noqa: BLE001 - extras must never take the headline line down
        out["error"] = repr(e)
    return out


def cpu_baseline(kind, n, metric, budget_s=15.0, threads=None):
    """The oracle port (the reference's own torch op sequence, oracle/siegel_oracle.py) timed on the
    host cores: gather -> dist -> distortion loss -> autograd backward, float64."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siegel_oracle as so

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    rows = 1 << 14
    table = make_table(kind, n, rows, torch.device("cpu"), seed=1).requires_grad_(True)
    b = 1 << 14
    idx, gd = make_pairs(rows, b, torch.device("cpu"), seed=100)

    def step():
        table.grad = None
        d = so.dist(kind, table[idx[:, 0]], table[idx[:, 1]], metric if kind != "spd" else "riem")
        distortion_loss(gd, d).backward()

    step()
    t0 = time.perf_counter()
    reps = 0
    while True:
        step()
        reps += 1
        el = time.perf_counter() - t0
        if el > budget_s or reps >= 50:
            break
    return {"value": b * reps / el, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"{reps} x {b} pairs of the same workload ({kind} n={n}), torch CPU float64, {el:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, n = args.kind, args.n
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import siegel_oracle as so
    torch.set_num_threads(threads)
    rows = 1 << 14
    b = 1 << 13
    table = make_table(kind, n, rows, torch.device("cpu"), seed=1).requires_grad_(True)
    idx, gd = make_pairs(rows, b, torch.device("cpu"), seed=100)

    def step():
        table.grad = None
        d = so.dist(kind, table[idx[:, 0]], table[idx[:, 1]], args.metric if kind != "spd" else "riem")
        distortion_loss(gd, d).backward()

    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = b * args.steps / el
    cfg_b = args.pairs if args.pairs else default_pairs(n)
    out = {
        "impl": "reference", "metric": "Siegel dist pairs/s fwd+bwd", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"distance microbench: {kind} n={n} metric={args.metric}, table 2^{args.rows.bit_length() - 1} rows, "
                               f"{cfg_b} pairs/GPU/step, fwd+bwd through AverageDistortionLoss",
                   "n": n, "kind": kind, "metric": args.metric},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": f"each step = {b} pairs of the workload (bounded sample), table 2^14 rows, torch CPU float64"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kind", default="upper", choices=["upper", "bounded", "spd"])
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--metric", default="riem")
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default by n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the n=10 / epoch-time side measurements")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

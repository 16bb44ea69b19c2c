"""Parity of the CUDA path (through the C ABI / the manifold API) with the oracle and with the golden
vectors of the unmodified reference.  Tolerances (BASELINE.json north_star):
  dist / vvd:  1e-9 relative (float64)
  gradients:   sym(grad) within 1e-6 max|g| for n <= 6, 1e-5 max|g| for n = 10 (the reference's own
               autograd noise is 1e-10 .. 3e-6 there)."""
import os

import numpy as np
import pytest
import torch

import siegel_oracle as so
from conftest import METRICS, golden_files, grad_tolerance, load_golden, sym

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-9, 1e-12


@pytest.fixture(scope="module")
def sb():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    import sympa_b200
    from sympa_b200 import _lib
    _lib.load()      # a missing libsympa_b200.so is an error on a GPU box, never a silent fallback
    return sympa_b200


def make_manifold(sb, kind, n, metric, w=None):
    cls = {"upper": sb.UpperHalfManifold, "bounded": sb.BoundedDomainManifold}[kind]
    man = cls(dims=n, metric=sb.MetricType.from_str(metric)).cuda()
    if metric == "wsum":
        man.metric.weights.data = torch.tensor(w, dtype=torch.float64, device="cuda").reshape(1, n)
    return man


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_manifold_dist_matches_reference_golden(sb, path):
    kind, n, regime, r = load_golden(path)
    z1 = torch.tensor(r["z1"], device="cuda")
    z2 = torch.tensor(r["z2"], device="cuda")
    go = torch.tensor(r["go"], device="cuda")
    for m in METRICS:
        man = make_manifold(sb, kind, n, m, r.get("wsum_w"))
        a1, a2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        d = man.dist(a1, a2)
        assert d.shape == (z1.shape[0],)
        (d * go).sum().backward()
        np.testing.assert_allclose(d.detach().cpu().numpy(), r["dist_" + m], rtol=RTOL)
        np.testing.assert_allclose(man.vvd(z1, z2).cpu().numpy(), r["vvd"], rtol=RTOL, atol=ATOL)
        gmax = max(np.abs(r["g1_" + m]).max(), np.abs(r["g2_" + m]).max())
        assert np.abs(a1.grad.cpu().numpy() - sym(r["g1_" + m])).max() <= grad_tolerance(n) * gmax
        assert np.abs(a2.grad.cpu().numpy() - sym(r["g2_" + m])).max() <= grad_tolerance(n) * gmax
        if m == "wsum":
            np.testing.assert_allclose(man.metric.weights.grad.cpu().numpy(), r["gw_wsum"], rtol=1e-9)
        with torch.no_grad():
            d0 = man.dist(z1, z2)
        # forward-only kernel vs forward+grad kernel: two instantiations of the same templates, equal up
        # to the compiler's FMA contraction choices
        torch.testing.assert_close(d0, d.detach(), rtol=1e-13, atol=1e-15)
    sb.ops.check_status()


@pytest.mark.parametrize("kind,n,metric", [("upper", 2, "riem"), ("upper", 4, "fmin"), ("bounded", 3, "fone"),
                                           ("upper", 6, "finf"), ("bounded", 5, "wsum"), ("upper", 10, "fmin"),
                                           ("upper", 1, "riem"), ("bounded", 1, "fone"), ("upper", 3, "wsum"), ("upper", 5, "fone"),
                                           ("upper", 7, "riem"), ("upper", 8, "fmin"), ("upper", 9, "wsum"),
                                           ("bounded", 2, "riem"), ("bounded", 4, "fmin"), ("bounded", 6, "finf"),
                                           ("bounded", 7, "fone"), ("bounded", 8, "riem"), ("bounded", 9, "fmin"),
                                           ("bounded", 10, "wsum"),
                                           ("spd", 1, "riem"), ("spd", 3, "riem"), ("spd", 4, "riem"), ("spd", 7, "riem"),
                                           ("spd", 10, "riem")])
def test_table_path_and_fused_step_match_oracle(sb, kind, n, metric):
    """fused gather + scatter-add backward, and the one-launch distortion step, against the oracle
    run the way the reference runs it: gather -> dist -> AverageDistortionLoss -> autograd."""
    g = torch.Generator().manual_seed(100 + n)
    rows, b = 37, 200
    if kind == "spd":
        table = so.spd_spread(rows, n, generator=g)
    else:
        table = so.upper_spread(rows, n, generator=g, scale=0.3)
        if kind == "bounded":
            table = so.to_symmetric(so.cayley_transform(table))
    src = torch.randint(0, rows, (b,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (b,), generator=g)) % rows
    idx = torch.stack((src, dst), 1)
    gdist = torch.randint(1, 20, (b,), generator=g).double()
    scale = 1.7
    w = torch.linspace(-0.2, 1.1, n, dtype=torch.float64).reshape(1, n) if metric == "wsum" else None

    # oracle
    tab_o = table.clone().requires_grad_(True)
    w_o = None if w is None else w.clone().requires_grad_(True)
    d_o = so.dist(kind, tab_o[idx[:, 0]], tab_o[idx[:, 1]], metric, w_o)
    loss_o = so.distortion_loss(gdist, d_o * scale)
    loss_o.backward()
    gt_o = sym(tab_o.grad.numpy())
    tol = grad_tolerance(n) * np.abs(gt_o).max()

    # CUDA, autograd path through the table
    man = (sb.SymmetricPositiveDefinite().cuda() if kind == "spd" else make_manifold(sb, kind, n, metric, None if w is None else w.numpy()))
    tab_c = table.cuda().requires_grad_(True)
    d_c = man.dist_from_table(tab_c, idx.cuda())
    loss_c = so.distortion_loss(gdist.cuda(), d_c * scale)
    loss_c.backward()
    np.testing.assert_allclose(d_c.detach().cpu().numpy(), d_o.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(loss_c.item(), loss_o.item(), rtol=1e-9)
    assert np.abs(tab_c.grad.cpu().numpy() - gt_o).max() <= tol
    if w is not None:
        np.testing.assert_allclose(man.metric.weights.grad.cpu().numpy(), w_o.grad.numpy(), rtol=1e-8, atol=1e-10)

    # CUDA, fused one-launch step
    gt = torch.zeros_like(tab_c)
    gs = torch.zeros(1, dtype=torch.float64, device="cuda")
    gw = torch.zeros(n, dtype=torch.float64, device="cuda") if w is not None else None
    dist_out = torch.empty(b, dtype=torch.float64, device="cuda")
    loss = sb.ops.distortion_step(kind, metric, tab_c.detach(), idx.cuda(), gdist.cuda(), scale, gt,
                                  wsum_w=None if w is None else w.cuda(), grad_wsum_w=gw, grad_scale=gs,
                                  dist_out=dist_out)
    np.testing.assert_allclose(loss.item(), loss_o.item(), rtol=1e-9)
    np.testing.assert_allclose(dist_out.cpu().numpy(), d_o.detach().numpy(), rtol=RTOL)
    assert np.abs(gt.cpu().numpy() - gt_o).max() <= tol
    if w is not None:
        np.testing.assert_allclose(gw.cpu().numpy(), w_o.grad.numpy().reshape(-1), rtol=1e-8, atol=1e-10)
    # d loss / d scale = sum sign * 2 r * d / g
    s_o = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
    so.distortion_loss(gdist, d_o.detach() * s_o).backward()
    np.testing.assert_allclose(gs.item(), s_o.grad.item(), rtol=1e-9)
    sb.ops.check_status()


@pytest.mark.parametrize("kind", ["upper", "bounded"])
@pytest.mark.parametrize("n", [2, 3, 4, 6])
def test_reference_metamorphic_properties(sb, kind, n):
    """tests/test_upper_half.py:109-186 and tests/test_bounded_domain.py:95-127 of the reference."""
    man = make_manifold(sb, kind, n, "riem")
    x, y = man.random(10).cuda(), man.random(10).cuda()
    torch.testing.assert_close(man.dist(x, y), man.dist(y, x), rtol=1e-5, atol=1e-8)
    dxx = man.dist(x, x)
    torch.testing.assert_close(dxx, torch.zeros_like(dxx), rtol=0, atol=1e-8)
    eye = torch.eye(n, dtype=torch.bool, device="cuda").expand(10, 2, n, n)
    xd, yd = torch.where(eye, x, torch.zeros_like(x)), torch.where(eye, y, torch.zeros_like(y))
    torch.testing.assert_close(man.dist(xd, yd), man.dist(yd, xd), rtol=1e-5, atol=1e-8)
    ref = so.dist(kind, xd.cpu(), yd.cpu())
    torch.testing.assert_close(man.dist(xd, yd).cpu(), ref, rtol=1e-9, atol=1e-12)
    if kind == "upper":
        xi = torch.stack((torch.zeros_like(x[:, 0]), x[:, 1]), 1)
        yi = torch.stack((torch.zeros_like(y[:, 0]), y[:, 1]), 1)
        torch.testing.assert_close(man.dist(xi, yi), man.dist(yi, xi), rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(man.dist(xi, yi).cpu(), so.dist(kind, xi.cpu(), yi.cpu()), rtol=1e-9, atol=1e-12)
    sb.ops.check_status()


def test_edge_cases(sb):
    man = make_manifold(sb, "upper", 3, "riem")
    # empty batch
    e = torch.empty(0, 2, 3, 3, dtype=torch.float64, device="cuda")
    assert man.dist(e, e).shape == (0,)
    # point outside the manifold -> status bit, raised lazily
    x = man.random(4).cuda()
    bad = x.clone()
    bad[:, 1] = -bad[:, 1]
    man.dist(bad, x)
    with pytest.raises(AssertionError, match="outside the manifold"):
        sb.ops.check_status()
    sb.ops.check_status()   # reset
    # out-of-range index -> status bit, pair skipped
    table = man.random(5).cuda()
    idx = torch.tensor([[0, 1], [2, 7]], device="cuda")
    d = man.dist_from_table(table, idx)
    assert d[1].item() == 0.0
    with pytest.raises(AssertionError, match="index out of range"):
        sb.ops.check_status()
    # wrong dtype / shape / n
    with pytest.raises(TypeError):
        man.dist(x.float(), x.float())
    with pytest.raises(ValueError):
        man.dist(x[:, 0], x[:, 0])
    big = torch.zeros(1, 2, 11, 11, dtype=torch.float64, device="cuda")
    with pytest.raises(NotImplementedError):
        man.dist(big, big)
    # repeated rows inside one batch accumulate (atomics)
    table = man.random(3).cuda().requires_grad_(True)
    idx = torch.tensor([[0, 1]] * 64, device="cuda")
    man.dist_from_table(table, idx).sum().backward()
    one = man.random(3).cuda()
    t1 = table.detach().clone().requires_grad_(True)
    man.dist_from_table(t1, idx[:1]).sum().backward()
    torch.testing.assert_close(table.grad, 64 * t1.grad, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("kind,n", [("upper", 2), ("upper", 4), ("bounded", 3), ("upper", 10)])
def test_full_size_properties(sb, kind, n):
    """Size-independent properties at a BASELINE-sized batch: symmetry d(a,b) = d(b,a), d(a,a) = 0,
    riem^2 = sum vvd^2, fone = sum vvd, finf = max vvd, ascending vvd, and conservation of the
    scatter-add (sum of the table gradient == sum of per-pair gradients)."""
    rows = 1 << 16
    b = (1 << 20) if n <= 4 else (1 << 17)
    g = torch.Generator().manual_seed(9)
    table = so.upper_spread(rows, n, generator=g, scale=0.3)
    if kind == "bounded":
        table = so.to_symmetric(so.cayley_transform(table))
    table = table.cuda()
    src = torch.randint(0, rows, (b,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (b,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).cuda()
    out = {}
    for m in ("riem", "fone", "finf"):
        man = make_manifold(sb, kind, n, m)
        with torch.no_grad():
            out[m] = man.dist_from_table(table, idx)
            rev = man.dist_from_table(table, idx.flip(1))
        torch.testing.assert_close(out[m], rev, rtol=1e-9, atol=1e-12)
    v = man.vvd(table[idx[:, 0]], table[idx[:, 1]])
    assert bool((v[:, 1:] >= v[:, :-1]).all())
    torch.testing.assert_close(out["riem"], v.norm(dim=-1), rtol=1e-12, atol=0)
    torch.testing.assert_close(out["fone"], v.sum(-1), rtol=1e-12, atol=0)
    # (dist on the table path and vvd on materialised operands may run different kernel instantiations - for the
    # bounded domain the by-rows route against the per-pair kernel - so equal up to FMA contraction, not bitwise)
    torch.testing.assert_close(out["finf"], v[:, -1], rtol=1e-12, atol=0)
    same = torch.stack((idx[:, 0], idx[:, 0]), 1)
    with torch.no_grad():
        dxx = man.dist_from_table(table, same)
    assert dxx.abs().max().item() < 1e-7
    t = table.clone().requires_grad_(True)
    man = make_manifold(sb, kind, n, "riem")
    man.dist_from_table(t, idx).sum().backward()
    z1 = table[idx[:, 0]].requires_grad_(True)
    z2 = table[idx[:, 1]].requires_grad_(True)
    man.dist(z1, z2).sum().backward()
    total = z1.grad.sum(0) + z2.grad.sum(0)
    torch.testing.assert_close(t.grad.sum(0), total, rtol=1e-8, atol=1e-8)
    sb.ops.check_status()


@pytest.mark.parametrize("n", [8, 9, 10])       # the cooperative sizes (n = 7 became a register kernel in round 2)
@pytest.mark.parametrize("metric", ["riem", "wsum"])
def test_split_path_equals_single_kernel_path(sb, n, metric):
    """upper, n > 6 (the cooperative kernels): the three-kernel path (state parked in scratch; used when the batch is large enough
    and scratch is given) must reproduce the single-kernel path bit for bit - forward, saved unit
    gradients and the fused distortion step."""
    from sympa_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(21 + n)
    rows, b = 500, 5000       # > the 2048-pair threshold, not a multiple of the chunk or warp sizes
    table = so.upper_spread(rows, n, generator=g, scale=0.3).cuda()
    src = torch.randint(0, rows, (b,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (b,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).cuda()
    gd = torch.randint(1, 20, (b,), generator=g).double().cuda()
    w = torch.linspace(-0.2, 1.1, n, dtype=torch.float64).cuda() if metric == "wsum" else None
    stream = torch.cuda.current_stream().cuda_stream
    status = sb.ops.status_word(table.device)

    def fwd(use_scratch):
        dist = torch.empty(b, dtype=torch.float64, device="cuda")
        vvd = torch.empty(b, n, dtype=torch.float64, device="cuda")
        saved = torch.empty(lib.sympa_workspace_bytes(0, n, b) // 8, dtype=torch.float64, device="cuda")
        scratch, nbytes = sb.ops.scratch_for("upper", n, b, table.device) if use_scratch else (None, 0)
        _lib.check(lib.sympa_dist_forward(0, n, _lib.METRIC[metric], b, None, None, table.data_ptr(), rows,
                                          idx.data_ptr(), None if w is None else w.data_ptr(), dist.data_ptr(),
                                          vvd.data_ptr(), saved.data_ptr(),
                                          None if scratch is None else scratch.data_ptr(), nbytes,
                                          status.data_ptr(), stream))
        return dist, vvd, saved

    def step(use_scratch):
        gt = torch.zeros_like(table)
        gs = torch.zeros(1, dtype=torch.float64, device="cuda")
        gw = torch.zeros(n, dtype=torch.float64, device="cuda")
        loss = torch.zeros(1, dtype=torch.float64, device="cuda")
        scratch, nbytes = sb.ops.scratch_for("upper", n, b, table.device) if use_scratch else (None, 0)
        _lib.check(lib.sympa_distortion_step(0, n, _lib.METRIC[metric], b, table.data_ptr(), rows, idx.data_ptr(),
                                             gd.data_ptr(), 1.3, None if w is None else w.data_ptr(), gt.data_ptr(),
                                             gw.data_ptr() if w is not None else None, gs.data_ptr(), loss.data_ptr(),
                                             None, None if scratch is None else scratch.data_ptr(), nbytes,
                                             status.data_ptr(), stream))
        return gt, gs, gw, loss

    d0, v0, s0 = fwd(False)
    a = step(False)
    assert lib.sympa_set_option(_lib.OPT_SPLIT_PATH, 1) == 0
    try:
        assert lib.sympa_scratch_bytes(0, n, b) > 0
        d1, v1, s1 = fwd(True)
        b_ = step(True)
    finally:
        lib.sympa_set_option(_lib.OPT_SPLIT_PATH, 0)
    assert torch.equal(d0, d1) and torch.equal(v0, v1) and torch.equal(s0, s1)
    for x, y in zip(a, b_):
        torch.testing.assert_close(x, y, rtol=1e-12, atol=1e-12)   # atomics: order differs, values do not
    sb.ops.check_status()



@pytest.mark.parametrize("kind,metric,n,graph", [("upper", "riem", 2, "grid5"), ("bounded", "fone", 3, "grid5"),
                                                 ("upper", "riem", 2, "config1"), ("bounded", "fone", 3, "config2"),
                                                 ("upper", "finf", 6, "product")])
def test_train_epoch_matches_oracle_loop(sb, kind, metric, n, graph):
    """One epoch of the training loop (runner.py:90-122: forward, distortion loss, backward, clip,
    RiemannianSGD) against the same loop driven by the oracle on the CPU: on a small grid graph, and at the
    full size of BASELINE.json configs[0] (grid 20x20: 400 nodes, 79 800 pairs, upper / riem / n = 2, batch 2048)
    and configs[1] (balanced tree 3/5: 364 nodes, 66 066 pairs, bounded / fone / n = 3, batch 2048)."""
    from types import SimpleNamespace
    from torch.nn.utils import clip_grad_norm_
    from sympa_b200.graphs import balanced_tree_triplets, grid_triplets, product_cartesian_triplets
    from sympa_b200.model import Model
    from sympa_b200.optim import RiemannianSGD
    from sympa_b200.runner import train_epoch

    torch.manual_seed(0)
    if graph == "grid5":
        (idx, gd, nodes), bs = grid_triplets(5, 2), 64
    elif graph == "product":     # BASELINE configs[2] in small: tree(2,3) x grid 3x3 = 135 nodes, 9045 pairs, upper / finf / n = 6
        (idx, gd, nodes), bs = product_cartesian_triplets(2, 3, 3, 2), 1024
    elif graph == "config1":
        (idx, gd, nodes), bs = grid_triplets(20, 2), 2048
    else:
        (idx, gd, nodes), bs = balanced_tree_triplets(3, 5), 2048
    args = SimpleNamespace(manifold=kind, metric=metric, dims=n, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                           train_scale=False)
    model = Model(args)
    table0 = model.embeddings.embeds.detach().clone()
    model = model.cuda()
    lr = 1e-2
    opt = RiemannianSGD(model.parameters(), lr=lr)
    mean_loss = train_epoch(model, opt, idx.cuda(), gd.cuda(), bs, max_grad_norm=50.0, shuffle=False)

    table = table0.clone()
    total, steps = 0.0, 0
    for s in range(0, idx.shape[0], bs):
        sel = slice(s, s + bs)
        t = table.clone().requires_grad_(True)
        d = so.dist(kind, t[idx[sel, 0]], t[idx[sel, 1]], metric)
        loss = so.distortion_loss(gd[sel], d)
        loss.backward()
        g = 0.5 * (t.grad + t.grad.transpose(-1, -2))
        holder = torch.nn.Parameter(torch.zeros_like(g))
        holder.grad = g
        clip_grad_norm_([holder], 50.0)
        table = so.rsgd_step(kind, table, holder.grad, lr)
        total += loss.item()
        steps += 1
    np.testing.assert_allclose(mean_loss, total / steps, rtol=1e-7)
    # the 33-39 steps of a full-size epoch accumulate the reference's own autograd gradient noise (SURVEY.md F7:
    # up to 2e-10 relative per step at n = 3): measured 2.2e-9 on one entry of 6552 at config 2
    atol = 1e-9 if graph == "grid5" else 5e-9
    torch.testing.assert_close(model.embeddings.embeds.detach().cpu(), table, rtol=1e-6, atol=atol)


@pytest.mark.parametrize("kind,n", [("upper", 2), ("upper", 3), ("upper", 4), ("upper", 6), ("upper", 10), ("spd", 3), ("spd", 6),
                                    ("bounded", 2), ("bounded", 3), ("bounded", 4), ("bounded", 7), ("bounded", 10)])
@pytest.mark.parametrize("lr", [1e-2, 5.0])
def test_fused_rsgd_step_matches_host_optimizer(sb, kind, n, lr):
    """sympa_rsgd_step (one launch: egrad2rgrad + retr + projx per row) against the torch implementation
    of the same update and against the oracle's restatement of the reference (upper_half.py:25-66)."""
    from sympa_b200.embeddings import ManifoldParameter
    from sympa_b200.optim import RiemannianSGD
    g = torch.Generator().manual_seed(5 + n)
    rows = 300
    if kind == "spd":
        table = so.spd_spread(rows, n, generator=g)
        man = sb.SymmetricPositiveDefinite()
        lr = min(lr, 0.05)
    else:
        table = so.upper_spread(rows, n, generator=g, scale=0.3)
        man = sb.UpperHalfManifold(dims=n)
        if kind == "bounded":
            table = so.to_symmetric(so.cayley_transform(table))
            man = sb.BoundedDomainManifold(dims=n)
    grad = torch.zeros_like(table)
    touched = torch.randperm(rows, generator=g)[: rows // 3]
    gr = torch.randn((len(touched),) + tuple(table.shape[1:]), dtype=torch.float64, generator=g)
    grad[touched] = 0.5 * (gr + gr.transpose(-1, -2))
    out = {}
    for fused in (False, True):
        p = ManifoldParameter(table.clone().cuda(), manifold=man)
        p.grad = grad.clone().cuda()
        RiemannianSGD([p], lr=lr, fused=fused).step()
        out[fused] = p.detach().cpu()
    torch.testing.assert_close(out[True], out[False], rtol=1e-9, atol=1e-11)
    untouched = torch.ones(rows, dtype=torch.bool)
    untouched[touched] = False
    assert torch.equal(out[True][untouched], table[untouched])
    # the oracle's restatement of the update (spd: geoopt's documented retraction, parity with geoopt itself unpinned)
    torch.testing.assert_close(out[True], so.rsgd_step(kind, table, grad, lr), rtol=1e-9, atol=1e-11)
    if kind in ("upper", "bounded") and lr > 1:
        assert man.projected_points > 0


def test_training_with_fused_optimizer_equals_host_optimizer(sb):
    from types import SimpleNamespace
    from sympa_b200.graphs import balanced_tree_triplets
    from sympa_b200.model import Model
    from sympa_b200.optim import RiemannianSGD
    from sympa_b200.runner import train_epoch
    idx, gd, nodes = balanced_tree_triplets(3, 3)
    args = SimpleNamespace(manifold="upper", metric="fone", dims=3, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                           train_scale=False)
    res = {}
    for fused in (False, True):
        torch.manual_seed(1)
        model = Model(args).cuda()
        opt = RiemannianSGD(model.parameters(), lr=5e-2, fused=fused)
        losses = [train_epoch(model, opt, idx.cuda(), gd.cuda(), 128, epoch=e, sync_stats=not fused) for e in range(3)]
        res[fused] = (losses, model.embeddings.embeds.detach().cpu())
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=1e-8)
    torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-7, atol=1e-10)
    assert res[True][0][-1] < res[True][0][0]      # it trains


@pytest.mark.parametrize("n", [2, 3, 4, 6, 10])
def test_far_apart_and_nearly_coincident_pairs(sb, n):
    """the clamp of siegel_manifold.py:69 (1 - d < eps, points very far apart: clustered singular values
    with derivative weights spanning orders of magnitude) and nearly coincident points (d ~ 1e-9)"""
    g = torch.Generator().manual_seed(77)
    base = so.upper_spread(6, n, generator=g, scale=0.2)
    far = base.clone()
    sc = torch.tensor([1e5, 3e5, 1e6, 1e7, 4e5, 2e6]).sqrt().reshape(-1, 1) * (1.0 + 0.3 * torch.arange(n)).reshape(1, -1)
    far[:, 1] = sc.unsqueeze(-1) * far[:, 1] * sc.unsqueeze(-2)
    near = base.clone()
    near[:, 0] = near[:, 0] + 1e-9 * so.sym(torch.randn(6, n, n, dtype=torch.float64, generator=g))
    near[:, 1] = near[:, 1] * (1 + 1e-9)
    for metric in ("riem", "fone"):
        man = make_manifold(sb, "upper", n, metric)
        d_ref, g1_ref, g2_ref, _ = so.dist_and_grads("upper", base, far, metric)
        a1, a2 = base.clone().cuda().requires_grad_(True), far.clone().cuda().requires_grad_(True)
        d = man.dist(a1, a2)
        d.sum().backward()
        np.testing.assert_allclose(d.detach().cpu().numpy(), d_ref.numpy(), rtol=1e-7)
        tol = 1e-6 if n <= 6 else 1e-5
        for b in range(6):
            for got, ref in ((a1.grad[b], g1_ref[b]), (a2.grad[b], g2_ref[b])):
                ref = 0.5 * (ref + ref.transpose(-1, -2))
                assert (got.cpu() - ref).abs().max() <= tol * ref.abs().max()
        d_ref = so.dist("upper", base, near, metric)
        a1 = base.clone().cuda().requires_grad_(True)
        d = man.dist(a1, near.cuda())
        d.sum().backward()
        np.testing.assert_allclose(d.detach().cpu().numpy(), d_ref.numpy(), rtol=1e-4, atol=1e-13)
        assert torch.isfinite(a1.grad).all()
    sb.ops.check_status()


@pytest.mark.parametrize("kind,n,metric", [("upper", 3, "riem"), ("bounded", 4, "fone"), ("upper", 7, "fmin"), ("spd", 4, "riem")])
def test_distance_matrix_matches_oracle_and_reference_semantics(sb, kind, n, metric):
    """Model.build_distance_matrix / manifold.dist_matrix against the oracle evaluated
    the way Runner.build_distance_matrix does (sympa/runner.py:142-154): row by row, exact zeros on the
    diagonal; the chunking (index workspace smaller than the matrix) and row blocks must not matter."""
    rows = 37
    g = torch.Generator().manual_seed(11)
    if kind == "spd":
        table = so.spd_spread(rows, n, generator=g)
        man = sb.SymmetricPositiveDefinite().cuda()
    else:
        table = so.upper_spread(rows, n, generator=g, scale=0.3)
        if kind == "bounded":
            table = so.to_symmetric(so.cayley_transform(table))
        man = make_manifold(sb, kind, n, metric)
    ref = torch.zeros(rows, rows, dtype=torch.float64)
    allr = torch.arange(rows)
    for i in range(rows):
        src = torch.full((rows,), i)
        src[i] = (i + 1) % rows
        d = so.dist(kind, table[src], table[allr], metric)
        d[i] = 0
        ref[i] = d
    t = table.cuda()
    full = man.dist_matrix(t)
    assert full.shape == (rows, rows)
    torch.testing.assert_close(full.cpu(), ref, rtol=RTOL, atol=ATOL)
    assert torch.all(torch.diagonal(full) == 0)
    # small index workspace (several chunks per call) and a row block
    small = sb.ops.dist_matrix(kind, metric if kind != "spd" else "riem", t, chunk_pairs=100)
    assert torch.equal(small, full)
    block = man.dist_matrix(t, 5, 9)
    assert torch.equal(block, full[5:14])
    assert man.dist_matrix(t, 3, 0).shape == (0, rows)
    sb.ops.check_status()


@pytest.mark.parametrize("kind,n", [("upper", 4), ("bounded", 3), ("spd", 5), ("upper", 8)])
@pytest.mark.parametrize("pairs", [7, 400])
def test_table_backward_with_workspace_equals_direct_scatter(sb, kind, n, pairs):
    """sympa_dist_backward_table (packed gradient table in a workspace + one expansion pass when the batch
    covers the table densely; overwrite / accumulate) against the direct scatter of sympa_dist_backward."""
    from sympa_b200 import _lib
    lib = _lib.load()
    rows = 60
    g = torch.Generator().manual_seed(3 + n)
    if kind == "spd":
        table = so.spd_spread(rows, n, generator=g).cuda()
    else:
        table = so.upper_spread(rows, n, generator=g, scale=0.3)
        if kind == "bounded":
            table = so.to_symmetric(so.cayley_transform(table))
        table = table.cuda()
    src = torch.randint(0, rows, (pairs,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (pairs,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).cuda()
    idx[min(3, pairs - 1), 1] = rows + 5         # one rejected pair: contributes nothing
    gdist = torch.randn(pairs, generator=g, dtype=torch.float64).cuda()
    _, _, saved = sb.ops.forward_raw(kind, "riem", table=table, idx=idx, want_grad=True)
    sb.ops.status_word(table.device).zero_()     # the rejected pair raised the bad-index bit on purpose
    k, stream = _lib.KIND[kind], torch.cuda.current_stream().cuda_stream
    ref = torch.zeros_like(table)
    _lib.check(lib.sympa_dist_backward(k, n, 0, pairs, gdist.data_ptr(), saved.data_ptr(), None, None, ref.data_ptr(),
                                       rows, idx.data_ptr(), None, None, None, stream))
    ws, ws_bytes = sb.ops.backward_workspace_for(kind, n, rows, table.device)
    assert (ws is None) == (lib.sympa_backward_workspace_bytes(k, n, rows) == 0)
    out = torch.full_like(table, float("nan"))   # overwrite mode must not read it
    _lib.check(lib.sympa_dist_backward_table(k, n, 0, pairs, gdist.data_ptr(), saved.data_ptr(), out.data_ptr(), rows,
                                             idx.data_ptr(), None, None, None, None if ws is None else ws.data_ptr(),
                                             ws_bytes, 1, stream))
    tol = 1e-12 * ref.abs().max().item()
    torch.testing.assert_close(out, ref, rtol=1e-12, atol=tol)
    acc = ref.clone()                            # accumulate mode: 2 x ref
    _lib.check(lib.sympa_dist_backward_table(k, n, 0, pairs, gdist.data_ptr(), saved.data_ptr(), acc.data_ptr(), rows,
                                             idx.data_ptr(), None, None, None, None if ws is None else ws.data_ptr(),
                                             ws_bytes, 0, stream))
    torch.testing.assert_close(acc, 2 * ref, rtol=1e-12, atol=2 * tol)
    # symmetric rows: exactly for the packed route (both triangles come from one accumulator), up to the
    # order of the atomics for the direct scatter
    torch.testing.assert_close(out, out.transpose(-1, -2), rtol=0, atol=tol)
    if ws is not None and 2 * pairs >= rows:
        assert torch.equal(out, out.transpose(-1, -2))


@pytest.mark.parametrize("manifold,metric,n,use_graph", [("upper", "riem", 2, True), ("bounded", "fone", 3, True),
                                                         ("upper", "wsum", 3, False), ("spd", "riem", 3, True)])
def test_fused_epoch_runner_equals_train_epoch(sb, manifold, metric, n, use_graph):
    """FusedEpochRunner (two launches per step, whole epoch replayed from a CUDA graph) against the step-by-step
    mirror of runner.py:90-122 with the same permutations, clipping and fused optimizer."""
    from types import SimpleNamespace
    from sympa_b200.graphs import balanced_tree_triplets
    from sympa_b200.model import Model
    from sympa_b200.optim import RiemannianSGD
    from sympa_b200.runner import FusedEpochRunner, train_epoch
    idx, gd, nodes = balanced_tree_triplets(3, 3)
    idx, gd = idx.cuda(), gd.cuda()
    args = SimpleNamespace(manifold=manifold, metric=metric, dims=n, num_points=nodes, scale_init=1.0, scale_coef=1.0,
                           train_scale=False)
    lr, batch, clip = 5e-2, 128, 3.0           # clip low enough to be active in the first steps
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)     # as the reference does (sympa/config.py:17-18): float64 wsum weights
    try:
        torch.manual_seed(2)
        ref = Model(args).cuda()
        torch.manual_seed(2)
        model = Model(args).cuda()
    finally:
        torch.set_default_dtype(prev)
    opt = RiemannianSGD(ref.parameters(), lr=lr, fused=True)
    ref_losses = [train_epoch(ref, opt, idx, gd, batch, max_grad_norm=clip, epoch=e) for e in range(3)]
    runner = FusedEpochRunner(model, lr, idx, gd, batch, max_grad_norm=clip, use_graph=use_graph)
    losses = [runner.run_epoch(e) for e in range(3)]
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-8)
    torch.testing.assert_close(model.embeddings.embeds.detach(), ref.embeddings.embeds.detach(), rtol=1e-7, atol=1e-10)
    if metric == "wsum":
        torch.testing.assert_close(model.manifold.metric.weights.detach(), ref.manifold.metric.weights.detach(),
                                   rtol=1e-7, atol=1e-10)
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("kind,n", [("upper", 4), ("bounded", 3), ("upper", 8)])
def test_sync_grad_backward_route_equals_plain_backward_on_one_rank(sb, kind, n):
    """dist_from_table(..., sync_grad=True): scatter into the packed gradient table / [all-reduce] / expansion as
    separate ABI calls.  In a single process the collective is a no-op and the gradient must equal the plain
    backward's (the multi-rank behaviour is checked by tests/multigpu/check_nccl_step.py under torchrun)."""
    rows, pairs = 50, 300
    g = torch.Generator().manual_seed(9)
    table = so.upper_spread(rows, n, generator=g, scale=0.3)
    if kind == "bounded":
        table = so.to_symmetric(so.cayley_transform(table))
    src = torch.randint(0, rows, (pairs,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (pairs,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).cuda()
    go = torch.randn(pairs, generator=g, dtype=torch.float64).cuda()
    man = make_manifold(sb, kind, n, "wsum", np.linspace(0.3, 1.2, n))
    grads = {}
    for sync in (False, True):
        t = table.clone().cuda().requires_grad_(True)
        man.metric.weights.grad = None
        d = man.dist_from_table(t, idx, sync_grad=sync)
        (d * go).sum().backward()
        grads[sync] = (t.grad.clone(), man.metric.weights.grad.clone())
    tol = 1e-12 * grads[False][0].abs().max().item()
    torch.testing.assert_close(grads[True][0], grads[False][0], rtol=1e-12, atol=tol)
    torch.testing.assert_close(grads[True][1], grads[False][1], rtol=1e-12, atol=1e-12)
    sb.ops.check_status()


@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_pair_feeder_double_buffering_delivers_batches_in_order(sb, idx_dtype):
    """PairFeeder: batches submitted from pinned host memory come out in order, as int64 indices and float64
    distances on the device, while later batches are already on the wire (slots are reused only after done())."""
    from sympa_b200.feeder import PairFeeder
    feeder = PairFeeder(torch.device("cuda", torch.cuda.current_device()))
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randint(0, 1000, (257, 2), generator=g).to(idx_dtype).pin_memory(),
                torch.rand(257, generator=g, dtype=torch.float64).pin_memory()) for _ in range(5)]
    feeder.submit(*batches[0])
    for k in range(5):
        idx_d, gd_d = feeder.next()
        if k + 1 < 5:
            feeder.submit(*batches[k + 1])
        assert idx_d.dtype == torch.int64 and idx_d.is_cuda and gd_d.dtype == torch.float64
        got_idx, got_gd = idx_d.clone(), gd_d.clone()     # "compute" on the current stream
        feeder.done()
        assert torch.equal(got_idx.cpu(), batches[k][0].to(torch.int64)) and torch.equal(got_gd.cpu(), batches[k][1])
    with pytest.raises(AssertionError):
        feeder.next()                                     # nothing left in flight


@pytest.mark.parametrize("shape", [(1000,), (257, 1), (0,)])
def test_fused_distortion_loss_equals_reference_expression(sb, shape):
    """sympa_b200.losses.AverageDistortionLoss on CUDA vectors (one kernel each way) against the reference's torch
    expression (sympa/losses.py:16-19), values and gradient, including the sign(0) = 0 convention of torch.abs."""
    from sympa_b200.losses import AverageDistortionLoss
    g = torch.Generator().manual_seed(4)
    gd = torch.randint(1, 20, shape, generator=g).double().cuda()
    d0 = (gd * (0.5 + torch.rand(shape, generator=g, dtype=torch.float64).cuda()))
    if d0.numel() > 3:
        d0.view(-1)[3] = gd.view(-1)[3]               # (d / g)^2 - 1 == 0 exactly: gradient 0
    outs = []
    for fused in (True, False):
        d = d0.clone().requires_grad_(True)
        if fused:
            loss = AverageDistortionLoss().calculate_loss(gd, d)
        else:
            loss = torch.abs(torch.pow(d / gd, 2) - 1).sum()
        (3.0 * loss).backward()
        outs.append((loss.detach(), d.grad.clone()))
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=1e-12, atol=1e-14)
    assert outs[0][1].shape == d0.shape


@pytest.mark.parametrize("n,metric", [(2, "finf"), (3, "fone"), (4, "riem"), (5, "fmin"), (6, "wsum"), (7, "riem")])
def test_bounded_by_rows_equals_per_pair_bounded_kernels_and_oracle(sb, n, metric):
    """Bounded domain on the table path with a dense batch: inverse Cayley transform once per table row
    (sympa_bounded_rows_to_upper), upper-half kernels on the pairs, gradient mapped back per row
    (sympa_bounded_rows_backward) - against the per-pair bounded kernels and the oracle (bounded_domain.py:27-39)."""
    rows, pairs = 40, 200
    g = torch.Generator().manual_seed(21 + n)
    table = so.to_symmetric(so.cayley_transform(so.upper_spread(rows, n, generator=g, scale=0.3)))
    src = torch.randint(0, rows, (pairs,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (pairs,), generator=g)) % rows
    idx = torch.stack((src, dst), 1)
    go = torch.rand(pairs, generator=g, dtype=torch.float64) + 0.5
    w = np.linspace(0.3, 1.2, n) if metric == "wsum" else None
    man = make_manifold(sb, "bounded", n, metric, w)
    res = {}
    for by_rows in (True, False):
        sb.ops.BOUNDED_BY_ROWS = by_rows
        try:
            t = table.clone().cuda().requires_grad_(True)
            if metric == "wsum":
                man.metric.weights.grad = None
            d = man.dist_from_table(t, idx.cuda())
            (d * go.cuda()).sum().backward()
            res[by_rows] = (d.detach().cpu(), t.grad.cpu(), None if metric != "wsum" else man.metric.weights.grad.cpu().clone())
        finally:
            sb.ops.BOUNDED_BY_ROWS = True
    gmax = res[False][1].abs().max().item()
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-11, atol=1e-13)
    assert (res[True][1] - res[False][1]).abs().max().item() <= 1e-9 * gmax
    if metric == "wsum":
        torch.testing.assert_close(res[True][2], res[False][2], rtol=1e-9, atol=1e-12)
    to = table.clone().requires_grad_(True)
    do = so.dist("bounded", to[idx[:, 0]], to[idx[:, 1]], metric, None if w is None else torch.tensor(w))
    (do * go).sum().backward()
    torch.testing.assert_close(res[True][0], do.detach(), rtol=RTOL, atol=ATOL)
    assert (res[True][1] - so.sym(to.grad)).abs().max().item() <= grad_tolerance(n) * gmax
    sb.ops.check_status()


@pytest.mark.parametrize("n", [7, 10])
def test_bounded_row_kernels_large_n(sb, n):
    """sympa_bounded_rows_to_upper / _backward for the sizes the dispatcher does not route through them (rolled row
    kernels): still exact against the oracle's inverse Cayley transform and its autograd backward."""
    from sympa_b200 import _lib
    lib = _lib.load()
    rows = 33
    g = torch.Generator().manual_seed(n)
    up = so.upper_spread(rows, n, generator=g, scale=0.3)
    z = so.to_symmetric(so.cayley_transform(up))
    zc = z.cuda()
    out = torch.empty_like(zc)
    st = sb.ops.status_word(zc.device)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.sympa_bounded_rows_to_upper(n, rows, zc.data_ptr(), out.data_ptr(), st.data_ptr(), stream))
    torch.testing.assert_close(out.cpu(), up, rtol=1e-9, atol=1e-11)
    gu = so.sym(torch.randn(rows, 2, n, n, generator=g, dtype=torch.float64))
    zr = z.clone().requires_grad_(True)
    (so.inverse_cayley_transform(zr) * gu).sum().backward()
    gz = torch.full_like(zc, float("nan"))
    _lib.check(lib.sympa_bounded_rows_backward(n, rows, zc.data_ptr(), gu.cuda().data_ptr(), gz.data_ptr(), 1, stream))
    ref = so.sym(zr.grad)
    assert (gz.cpu() - ref).abs().max().item() <= 1e-9 * ref.abs().max().item()
    sb.ops.check_status()


@pytest.mark.parametrize("kind,n", [("upper", 4), ("upper", 10), ("spd", 5)])
def test_table_grad_scatter_rows_equals_full_scatter(sb, kind, n):
    """sympa_table_grad_scatter_rows over a partition of the rows (what a pipelined data-parallel backward calls,
    one range at a time) fills the packed gradient table exactly like sympa_table_grad_scatter; an empty batch
    only zeroes its range."""
    from sympa_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(31 + n)
    rows, pairs = 97, 700
    table = (so.spd_spread(rows, n, generator=g) if kind == "spd" else so.upper_spread(rows, n, generator=g, scale=0.3)).cuda()
    src = torch.randint(0, rows, (pairs,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (pairs,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).cuda()
    gdist = torch.rand(pairs, dtype=torch.float64, generator=g).cuda() + 0.5
    _, _, saved = sb.ops.forward_raw(kind, "riem", table=table, idx=idx, want_grad=True)
    k = _lib.KIND[kind]
    need = lib.sympa_backward_workspace_bytes(k, n, rows)
    assert need == rows * (1 if kind == "spd" else 2) * (n * (n + 1) // 2) * 8
    stream = torch.cuda.current_stream().cuda_stream
    full = torch.full((need // 8,), 7.0, dtype=torch.float64, device="cuda")
    _lib.check(lib.sympa_table_grad_scatter(k, n, 0, pairs, gdist.data_ptr(), saved.data_ptr(), rows, idx.data_ptr(), None, None,
                                            None, full.data_ptr(), need, stream))
    parts = torch.full((need // 8,), -3.0, dtype=torch.float64, device="cuda")
    for lo, hi in ((0, 10), (10, 64), (64, 64), (64, rows)):
        _lib.check(lib.sympa_table_grad_scatter_rows(k, n, pairs, gdist.data_ptr(), saved.data_ptr(), rows, idx.data_ptr(), lo, hi,
                                                     parts.data_ptr(), need, stream))
    torch.testing.assert_close(parts, full, rtol=1e-12, atol=1e-12 * full.abs().max().item())   # atomics: order differs
    per_s = need // 8 // rows
    _lib.check(lib.sympa_table_grad_scatter_rows(k, n, 0, None, None, rows, None, 5, 9, parts.data_ptr(), need, stream))
    assert float(parts[5 * per_s: 9 * per_s].abs().max()) == 0.0
    torch.testing.assert_close(parts[9 * per_s:], full[9 * per_s:], rtol=1e-12, atol=1e-12 * full.abs().max().item())
    assert lib.sympa_table_grad_scatter_rows(k, n, pairs, gdist.data_ptr(), saved.data_ptr(), rows, idx.data_ptr(), 5, rows + 1,
                                             parts.data_ptr(), need, stream) == 1      # range outside the table
    sb.ops.check_status()


@pytest.mark.parametrize("kind,n,metric", [("upper", 4, "riem"), ("bounded", 3, "fone"), ("upper", 10, "wsum"), ("spd", 5, "riem"),
                                           ("bounded", 8, "fmin")])
def test_table_grad_accumulator_equals_one_call(sb, kind, n, metric):
    """Gradient accumulation over the chunks of a step (TableGradAccumulator, runner.py:104 grad_accum_steps): three
    dist_from_table calls sharing an accumulator + finish() give the table gradient (and dL/dw) of one call on the
    whole batch, and the oracle's; a second step starts from a clean accumulator; the cached transformed table of
    the bounded domain follows in-place updates of the table."""
    g = torch.Generator().manual_seed(77 + n)
    rows, b = 41, 300
    if kind == "spd":
        table = so.spd_spread(rows, n, generator=g)
    else:
        table = so.upper_spread(rows, n, generator=g, scale=0.3)
        if kind == "bounded":
            table = so.to_symmetric(so.cayley_transform(table))
    src = torch.randint(0, rows, (b,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (b,), generator=g)) % rows
    idx = torch.stack((src, dst), 1)
    gdist = torch.randint(1, 20, (b,), generator=g).double()
    w = torch.linspace(-0.2, 1.1, n, dtype=torch.float64).reshape(1, n) if metric == "wsum" else None
    man = (sb.SymmetricPositiveDefinite().cuda() if kind == "spd" else make_manifold(sb, kind, n, metric, None if w is None else w.numpy()))
    idx_c, gd_c = idx.cuda(), gdist.cuda()

    def one_call(t):
        d = man.dist_from_table(t, idx_c)
        so.distortion_loss(gd_c, d).backward()

    ref = table.cuda().requires_grad_(True)
    one_call(ref)
    gw_ref = None if w is None else man.metric.weights.grad.clone()
    if w is not None:
        man.metric.weights.grad = None

    t = table.cuda().requires_grad_(True)
    acc = man.table_grad_accumulator(t)
    for step in range(2):
        t.grad = None
        for lo, hi in ((0, 100), (100, 101), (101, b)):
            d = man.dist_from_table(t, idx_c[lo:hi], accumulator=acc)
            so.distortion_loss(gd_c[lo:hi], d).backward()
        assert t.grad is None                      # nothing dense until the step is finished
        acc.finish()
        torch.testing.assert_close(t.grad, ref.grad, rtol=1e-11, atol=1e-12 * ref.grad.abs().max().item())
        if w is not None:
            torch.testing.assert_close(man.metric.weights.grad, gw_ref, rtol=1e-11, atol=1e-13)
            man.metric.weights.grad = None
    # against the oracle
    to = table.clone().requires_grad_(True)
    wo = None if w is None else w.clone().requires_grad_(True)
    so.distortion_loss(gdist, so.dist(kind, to[idx[:, 0]], to[idx[:, 1]], metric, wo)).backward()
    go = so.sym(to.grad)
    assert (t.grad.cpu() - go).abs().max().item() <= grad_tolerance(n) * go.abs().max().item()
    # finish() accumulates into an existing gradient
    before = t.grad.clone()
    d = man.dist_from_table(t, idx_c, accumulator=acc)
    so.distortion_loss(gd_c, d).backward()
    acc.finish()
    torch.testing.assert_close(t.grad, 2 * before, rtol=1e-11, atol=1e-12 * before.abs().max().item())
    # in-place update of the table (an optimizer step): the next step must see the new points
    with torch.no_grad():
        t.copy_(torch.roll(t, 1, 0))
    t.grad = None
    d_new = man.dist_from_table(t, idx_c, accumulator=acc)
    torch.testing.assert_close(d_new.detach(), man.dist_from_table(t.detach(), idx_c), rtol=1e-13, atol=1e-15)
    assert not torch.allclose(d_new.detach(), d.detach())
    sb.ops.check_status()


@pytest.mark.parametrize("scatter_sms", [0, 1, 48, 1 << 20])
def test_accumulator_scatter_on_side_stream_sm_partition(sb, scatter_sms):
    """The accumulator's scatter on its side stream, confined to `scatter_sms` SMs (ticket-pulled work, last CTA
    drains): every work item exactly once wherever the CTAs land - 1 SM, a partition, more SMs than there are -
    and the streams joined by finish().  Against one call on the whole batch."""
    n, rows, b = 4, 3000, 200_000
    g = torch.Generator().manual_seed(5)
    table = so.upper_spread(rows, n, generator=g, scale=0.3).cuda()
    src = torch.randint(0, rows, (b,), generator=g)
    dst = (src + 1 + torch.randint(0, rows - 1, (b,), generator=g)) % rows
    idx = torch.stack((src, dst), 1).cuda()
    gd = torch.randint(1, 20, (b,), generator=g).double().cuda()
    man = make_manifold(sb, "upper", n, "riem")
    ref = table.clone().requires_grad_(True)
    so.distortion_loss(gd, man.dist_from_table(ref, idx)).backward()
    t = table.clone().requires_grad_(True)
    acc = man.table_grad_accumulator(t, scatter_sms=scatter_sms)
    for step in range(2):
        t.grad = None
        for lo, hi in ((0, 70_000), (70_000, 70_001), (70_001, b)):
            if step == 1 and hi == b:
                acc.last_chunk()        # hint: whole GPU for the scatter that nothing follows
            so.distortion_loss(gd[lo:hi], man.dist_from_table(t, idx[lo:hi], accumulator=acc)).backward()
        acc.finish()
        torch.testing.assert_close(t.grad, ref.grad, rtol=1e-10, atol=1e-12 * ref.grad.abs().max().item())
    # a host that enqueues many steps ahead of the device must not make the allocator grow (the state a side-stream
    # scatter reads is held until the caller's stream has joined it, then reused in stream order)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    peaks = []
    for step in range(12):
        t.grad = None
        for lo, hi in ((0, 70_000), (70_000, 140_000), (140_000, b)):
            so.distortion_loss(gd[lo:hi], man.dist_from_table(t, idx[lo:hi], accumulator=acc)).backward()
        acc.finish()
        peaks.append(torch.cuda.max_memory_allocated() - base)
    torch.cuda.synchronize()
    if not os.environ.get("SYMPA_UNDER_SANITIZER"):     # (compute-sanitizer keeps freed blocks accounted: measured there, not here)
        assert peaks[-1] <= 1.05 * peaks[1], peaks
    torch.testing.assert_close(t.grad, ref.grad, rtol=1e-10, atol=1e-12 * ref.grad.abs().max().item())
    sb.ops.check_status()


@pytest.mark.parametrize("kind,n", [("upper", 2), ("upper", 4), ("upper", 10), ("bounded", 3), ("bounded", 10), ("spd", 5), ("spd", 10)])
def test_check_points_kernel_equals_reference_predicate(sb, kind, n):
    """sympa_check_points (Embeddings.check_all_points, sympa/embeddings.py:41-47, in one launch) against the manifold's
    own check_point_on_manifold evaluated point by point as the reference does: verdict, FIRST offending row, reason."""
    g = torch.Generator().manual_seed(90 + n)
    rows = 257
    if kind == "spd":
        table = so.spd_spread(rows, n, generator=g)
        man = sb.SymmetricPositiveDefinite()
    else:
        table = so.upper_spread(rows, n, generator=g, scale=0.3)
        if kind == "bounded":
            table = so.to_symmetric(so.cayley_transform(table))
        man = make_manifold(sb, kind, n, "riem")

    def reference(t):
        for i in range(len(t)):
            ok, reason = man.check_point_on_manifold(t[i], explain=True)
            if not ok:
                return False, i, reason
        return True, None, None

    def both(t):
        got = sb.ops.check_points(kind, t.cuda())
        want = reference(t)
        assert got == want, (got, want)
        return got

    assert both(table) == (True, None, None)
    # an asymmetric entry beyond the tolerance in row 200 and a smaller one (inside the tolerance) in row 100
    t = table.clone()
    t[100, ..., 0, 1] += 5e-6 if n > 1 else 0.0
    t[200, ..., 1, 0] += 1e-3
    # (bounded: the in-tolerance asymmetry of row 100 already breaks the 1e-8 Hermitian test of I - conj(z) z)
    assert both(t)[:2] == (False, 200 if kind != "bounded" else 100)
    # the kind's own predicate fails in row 150 (and the later row 200 stays asymmetric: the FIRST one is reported)
    if kind == "upper":
        t[150, 1] = torch.diag(torch.tensor([-1.0] + [1.0 + 0.1 * k for k in range(n - 1)], dtype=torch.float64))   # det < 0
    elif kind == "spd":
        t[150] = t[150] - (torch.linalg.eigvalsh(t[150]).min() + 0.5) * torch.eye(n, dtype=torch.float64)
    if kind != "bounded":          # (I - conj(z) z is Hermitian for every symmetric z: that predicate cannot fail alone)
        got = both(t)
        assert got[:2] == (False, 150) and got[2] == sb.ops.CHECK_REASONS[2][kind]
    # NaN anywhere: not on the manifold
    t = table.clone()
    t[7, ..., 0, 0] = float("nan")
    assert both(t)[:2] == (False, 7)
    # the model-level entry point uses the kernel for a table on the GPU
    from sympa_b200.embeddings import MatrixEmbeddings
    emb = MatrixEmbeddings(rows, n, man).cuda()
    with torch.no_grad():
        emb.embeds.copy_(table)
    assert emb.check_all_points() == (True, None, None)
    with torch.no_grad():
        emb.embeds[33, ..., 0, n - 1] += 1.0
    ok, point, reason = emb.check_all_points()
    assert not ok and reason == sb.ops.CHECK_REASONS[1][kind].format(atol=1e-5, rtol=1e-5) and torch.equal(point, emb.embeds.data[33])

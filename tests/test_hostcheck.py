"""The per-pair C++ templates of sympa_b200/csrc/pair_math*.{cuh,inc} - the exact code the CUDA
kernels instantiate - compiled for the host and checked against the reference golden vectors and
the reference's metamorphic properties.  (CPU only; the same checks run through the C ABI on the
GPU in tests/test_parity_gpu.py.)"""
import numpy as np
import pytest
import torch

import siegel_oracle as so
from conftest import METRICS, golden_files, grad_tolerance, load_golden, sym

VVD_RTOL = 1e-9   # north_star: vvd and distances within 1e-9 relative in float64
VVD_ATOL = 1e-12


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_templates_match_reference_golden(hostcheck, path):
    kind, n, regime, r = load_golden(path)
    variants = (0, 1) if n <= 6 else (1,)
    for variant in variants:
        for m in METRICS:
            w = r["wsum_w"] if m == "wsum" else None
            d, v, g1, g2, st = hostcheck(variant, kind, n, m, r["z1"], r["z2"], w)
            assert st == 0
            np.testing.assert_allclose(v, r["vvd"], rtol=VVD_RTOL, atol=VVD_ATOL)
            np.testing.assert_allclose(d, r["dist_" + m], rtol=VVD_RTOL)
            go = r["go"][:, None, None, None]
            gmax = max(np.abs(r["g1_" + m]).max(), np.abs(r["g2_" + m]).max())
            assert np.abs(go * g1 - sym(r["g1_" + m])).max() <= grad_tolerance(n) * gmax
            assert np.abs(go * g2 - sym(r["g2_" + m])).max() <= grad_tolerance(n) * gmax
            d0, v0, _, _, _ = hostcheck(variant, kind, n, m, r["z1"], r["z2"], w, grad=False)
            assert np.array_equal(d0, d) and np.array_equal(v0, v)   # forward-only == forward+grad


@pytest.mark.parametrize("n", [2, 3, 4, 7, 10])
def test_spd_against_oracle(hostcheck, n):
    g = torch.Generator().manual_seed(5 + n)
    x, y = so.spd_spread(16, n, generator=g), so.spd_spread(16, n, generator=g)
    go = torch.rand(16, generator=g, dtype=torch.float64) + 0.5
    d, g1, g2, _ = so.dist_and_grads("spd", x, y, grad_out=go)
    for variant in ((0, 1, 2) if n <= 6 else (1, 2)):     # 2 = cooperative kernel templates
        dd, vv, h1, h2, st = hostcheck(variant, "spd", n, "riem", x.numpy(), y.numpy())
        assert st == 0
        np.testing.assert_allclose(dd, d.numpy(), rtol=1e-9)
        np.testing.assert_allclose(vv, so.spd_vvd(x, y).numpy(), rtol=1e-9, atol=1e-12)
        gmax = g1.abs().max().item()
        assert np.abs(go.numpy()[:, None, None] * h1 - sym(g1.numpy())).max() <= 1e-9 * gmax
        assert np.abs(go.numpy()[:, None, None] * h2 - sym(g2.numpy())).max() <= 1e-9 * gmax


@pytest.mark.parametrize("kind", ["upper", "bounded"])
@pytest.mark.parametrize("n", [2, 3, 4])
def test_reference_properties(hostcheck, kind, n):
    """symmetry under the sign flips / diagonal / pure-imaginary inputs and dist(x, x) = 0
    (reference tests/test_upper_half.py:109-186, tests/test_bounded_domain.py:95-127)."""
    g = torch.Generator().manual_seed(42)
    x, y = so.upper_random(10, n, generator=g), so.upper_random(10, n, generator=g)
    cases = {"plain": (x, y), "neg_real": (so.cplx(-so.re(x), so.im(x)), y)}
    eye = torch.eye(n, dtype=torch.bool).expand(10, 2, n, n)
    cases["diagonal"] = (torch.where(eye, x, torch.zeros_like(x)), torch.where(eye, y, torch.zeros_like(y)))
    cases["pure_imag"] = (so.cplx(torch.zeros_like(so.re(x)), so.im(x)), so.cplx(torch.zeros_like(so.re(y)), so.im(y)))
    xs = x.clone()
    xs[:, 0] *= 1.001
    cases["small_perturbation"] = (x, xs)
    for name, (a, b) in cases.items():
        if kind == "bounded":
            a, b = so.to_symmetric(so.cayley_transform(a)), so.to_symmetric(so.cayley_transform(b))
        for m in METRICS:
            w = np.linspace(0.2, 1.0, n) if m == "wsum" else None
            dab, _, _, _, st1 = hostcheck(0, kind, n, m, a.numpy(), b.numpy(), w)
            dba, _, _, _, st2 = hostcheck(0, kind, n, m, b.numpy(), a.numpy(), w)
            assert st1 == 0 and st2 == 0, name
            np.testing.assert_allclose(dab, dba, rtol=1e-5, atol=1e-8, err_msg=name)
            ref = so.dist(kind, a, b, m, None if w is None else torch.tensor(w))
            np.testing.assert_allclose(dab, ref.numpy(), rtol=1e-9, atol=1e-12, err_msg=name)
        daa, vaa, g1, g2, st = hostcheck(0, kind, n, "riem", a.numpy(), a.numpy())
        assert st == 0 and np.all(np.abs(daa) < 1e-8) and np.all(np.isfinite(g1)) and np.all(np.isfinite(g2))


def test_status_flags(hostcheck):
    z = so.upper_random(2, 3, generator=torch.Generator().manual_seed(0)).numpy()
    bad = z.copy()
    bad[:, 1] = -bad[:, 1]          # Im Z negative definite: outside the manifold
    _, _, _, _, st = hostcheck(0, "upper", 3, "riem", bad, z)
    assert st & 1


@pytest.mark.parametrize("path", golden_files("*_n*.npz"), ids=lambda p: p.split("/")[-1][:-4])
def test_cooperative_templates_match_reference_golden(hostcheck, path):
    """namespace sympa::coop (the warp-cooperative shared-memory kernel, n >= 5 on the GPU) with the
    lanes of a group emulated one after the other; shared memory is poisoned before every pair."""
    kind, n, regime, r = load_golden(path)
    if n < 2:
        pytest.skip("the cooperative templates start at n = 2 (the library never uses them below n = 7)")
    for m in METRICS:
        w = r["wsum_w"] if m == "wsum" else None
        d, v, g1, g2, st = hostcheck(2, kind, n, m, r["z1"], r["z2"], w)
        assert st == 0
        if kind == "upper":
            # the three-kernel split path (state parked in scratch) must give bit-identical results
            ds, vs_, g1s, g2s, sts = hostcheck(3, kind, n, m, r["z1"], r["z2"], w)
            assert sts == 0 and np.array_equal(ds, d) and np.array_equal(vs_, v)
            assert np.array_equal(g1s, g1) and np.array_equal(g2s, g2)
        np.testing.assert_allclose(v, r["vvd"], rtol=VVD_RTOL, atol=VVD_ATOL)
        np.testing.assert_allclose(d, r["dist_" + m], rtol=VVD_RTOL)
        go = r["go"][:, None, None, None]
        gmax = max(np.abs(r["g1_" + m]).max(), np.abs(r["g2_" + m]).max())
        assert np.abs(go * g1 - sym(r["g1_" + m])).max() <= grad_tolerance(n) * gmax
        assert np.abs(go * g2 - sym(r["g2_" + m])).max() <= grad_tolerance(n) * gmax
        d0, v0, _, _, _ = hostcheck(2, kind, n, m, r["z1"], r["z2"], w, grad=False)
        assert np.array_equal(d0, d) and np.array_equal(v0, v)


def _extreme_pairs(n):
    """pairs that exercise the clamp of siegel_manifold.py:69 (1 - d < eps) and nearly coincident points"""
    g = torch.Generator().manual_seed(77)
    base = so.upper_spread(6, n, generator=g, scale=0.2)
    far = base.clone()
    # Y2 = S Y1 S with S = diag(scale * (1, 1.3, 1.6, ..)): distinct, very large vector-valued distances
    # (v ~ 11 .. 17, so the clamp at log(2 / eps) = 12.2 is active for some of them)
    sc = torch.tensor([1e5, 3e5, 1e6, 1e7, 4e5, 2e6]).sqrt().reshape(-1, 1) * (1.0 + 0.3 * torch.arange(n)).reshape(1, -1)
    far[:, 1] = sc.unsqueeze(-1) * far[:, 1] * sc.unsqueeze(-2)
    near = base.clone()
    near[:, 0] = near[:, 0] + 1e-9 * so.sym(torch.randn(6, n, n, dtype=torch.float64, generator=g))
    near[:, 1] = near[:, 1] * (1 + 1e-9)
    return base, far, near


@pytest.mark.parametrize("n", [2, 3, 4, 6])
def test_clamped_and_nearly_coincident_pairs(hostcheck, n):
    base, far, near = _extreme_pairs(n)
    variants = (0, 2) if n <= 6 else (1, 2)
    for metric in ("riem", "fone", "finf"):
        for other, name in ((far, "far"), (near, "near")):
            if name == "far" and metric == "finf":
                # with EVERY value clamped the singular values agree to ~1e-7 relative and the choice of
                # "the largest" is ill-conditioned (in the reference too): order-dependent metrics are
                # only compared where the values are separated
                continue
            d_ref, g1_ref, g2_ref, _ = so.dist_and_grads("upper", base, other, metric)
            if name == "far":
                assert (so.upper_vvd(base, other) > 12.2).any()       # the clamp is active for some values
            for variant in variants:
                d, v, g1, g2, st = hostcheck(variant, "upper", n, metric, base.numpy(), other.numpy())
                assert st == 0
                if name == "far":
                    np.testing.assert_allclose(d, d_ref.numpy(), rtol=1e-7)
                    for b in range(g1.shape[0]):     # per pair: gradients span 1e-8 .. 1 across this batch
                        gmax = g1_ref[b].abs().max().item()
                        assert np.abs(g1[b] - sym(g1_ref[b].numpy())).max() <= 1e-6 * gmax
                else:
                    # d ~ 1e-9: absolute agreement at the level of the reference's own round-off
                    np.testing.assert_allclose(d, d_ref.numpy(), rtol=1e-4, atol=1e-13)
                    assert np.all(np.isfinite(g1)) and np.all(np.isfinite(g2))


@pytest.mark.parametrize("G", [2, 3, 4, 5])
def test_one_directional_ring_schedule_meets_every_pair_once_per_sweep(hostcheck, G):
    """coop::ring_send_masks (the schedule WarpExec::jacobi runs on the GPU for n = 2G - 1, 2G): G lanes hold two columns,
    after every round each lane passes ONE of them to its right neighbour.  Simulated here from the packed table the
    kernels use: in the N - 1 rounds of a sweep all N (N - 1) / 2 column pairs meet exactly once, from any starting
    arrangement, sweep after sweep, and no pair meets in two consecutive rounds across a sweep boundary."""
    packed = hostcheck.ring_send_masks(G)
    N = 2 * G
    rng = np.random.default_rng(G)
    for trial in range(5):
        labels = rng.permutation(N)
        kept = [int(labels[2 * g]) for g in range(G)]        # slot 0 of lane g ("top")
        recv = [int(labels[2 * g + 1]) for g in range(G)]    # slot 1 ("bottom": the column received last)
        last_round = set()
        for sweep in range(4):
            met = set()
            for r in range(N - 1):
                pairs = {(min(a, b), max(a, b)) for a, b in zip(kept, recv)}
                assert len(pairs) == G and not (pairs & met), (G, sweep, r)
                if r == 0:
                    assert not (pairs & last_round)
                met |= pairs
                last_round = pairs
                bits = [(packed >> (r * G + g)) & 1 for g in range(G)]
                sent = [recv[g] if bits[g] else kept[g] for g in range(G)]
                kept = [kept[g] if bits[g] else recv[g] for g in range(G)]
                recv = [sent[(g - 1) % G] for g in range(G)]
            assert len(met) == N * (N - 1) // 2
    assert packed >> (G * (N - 1)) == 0      # nothing beyond the N - 1 masks

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
    variants = (0, 1) if n <= 4 else (1,)
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
    for variant in ((0, 1, 2) if n <= 4 else (1, 2)):     # 2 = cooperative kernel templates
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

"""Optimizer-side functions against outputs of the UNMODIFIED reference (tests/golden/optim.npz, written by
oracle/gen_golden.py): UpperHalfManifold.egrad2rgrad / projx / retr (sympa/manifolds/upper_half.py:25-66,
siegel_manifold.py:74-87) and BoundedDomainManifold.egrad2rgrad (bounded_domain.py:41-53).  Checked here:

  * the oracle's restatements (CPU);
  * the package's torch mirrors of the manifold API (CPU);
  * the row-update templates the rsgd kernel instantiates, compiled for the host (CPU);
  * the fused kernel itself through the C ABI (GPU).

BoundedDomainManifold.projx / retr crash in the reference as shipped and cannot be pinned: the
bounded row is pinned through egrad2rgrad with a step small enough that the projection is the identity."""
import os

import numpy as np
import pytest
import torch

import siegel_oracle as so
from conftest import GOLDEN, sym

NS = (2, 3, 4, 6, 10)
T = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731


@pytest.fixture(scope="module")
def gold():
    r = np.load(os.path.join(GOLDEN, "optim.npz"))
    return {k: r[k] for k in r.files}


@pytest.mark.parametrize("n", NS)
def test_oracle_optimizer_functions_match_reference(gold, n):
    z, u = T(gold[f"upper_z_n{n}"]), T(gold[f"upper_u_n{n}"])
    np.testing.assert_allclose(so.upper_egrad2rgrad(z, u).numpy(), gold[f"upper_egrad2rgrad_n{n}"], rtol=1e-12, atol=1e-13)
    step = T(gold[f"upper_step_n{n}"])
    np.testing.assert_allclose(so.upper_projx(z + step).numpy(), gold[f"upper_projx_n{n}"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(so.upper_projx(z + step).numpy(), gold[f"upper_retr_n{n}"], rtol=1e-11, atol=1e-13)
    # the big step really leaves the manifold for some rows (otherwise this pins nothing about the clamp)
    lam = torch.linalg.eigvalsh(so.im(so.to_symmetric(z + step)))
    assert (lam.min(-1).values <= so.EPS[torch.float64]).any()
    zb = T(gold[f"bounded_z_n{n}"])
    np.testing.assert_allclose(so.bounded_egrad2rgrad(zb, u).numpy(), gold[f"bounded_egrad2rgrad_n{n}"], rtol=1e-12,
                               atol=1e-13)
    # the whole step of the oracle (what the GPU optimizer tests compare with)
    np.testing.assert_allclose(so.rsgd_step("upper", z, u, 2.5).numpy(), gold[f"upper_retr_n{n}"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(so.rsgd_step("upper", z, u, 1e-3).numpy(), gold[f"upper_retr_small_n{n}"], rtol=1e-12,
                               atol=1e-14)


@pytest.mark.parametrize("n", NS)
def test_manifold_api_mirrors_match_reference(gold, n):
    from sympa_b200 import BoundedDomainManifold, UpperHalfManifold
    up, bd = UpperHalfManifold(dims=n), BoundedDomainManifold(dims=n)
    z, u = T(gold[f"upper_z_n{n}"]), T(gold[f"upper_u_n{n}"])
    np.testing.assert_allclose(up.egrad2rgrad(z, u).numpy(), gold[f"upper_egrad2rgrad_n{n}"], rtol=1e-12, atol=1e-13)
    step = T(gold[f"upper_step_n{n}"])
    np.testing.assert_allclose(up.projx(z + step).numpy(), gold[f"upper_projx_n{n}"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(up.retr(z, step).numpy(), gold[f"upper_retr_n{n}"], rtol=1e-11, atol=1e-13)
    zb = T(gold[f"bounded_z_n{n}"])
    np.testing.assert_allclose(bd.egrad2rgrad(zb, u).numpy(), gold[f"bounded_egrad2rgrad_n{n}"], rtol=1e-12, atol=1e-13)


def _bounded_small_step(gold, n, lr):
    """reference update of a bounded row while the projection is the identity: sym(z - lr A u A)"""
    zb = gold[f"bounded_z_n{n}"]
    return sym(zb) - lr * sym(gold[f"bounded_egrad2rgrad_n{n}"])


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n", NS)
def test_row_update_templates_match_reference(hostcheck, gold, n, variant):
    if variant == 0 and n > 6:
        pytest.skip("the unrolled namespace is instantiated up to n = 6 in the host harness")
    z, u = gold[f"upper_z_n{n}"], gold[f"upper_u_n{n}"]
    out, projected = hostcheck.rsgd(variant, "upper", n, z, u, 2.5)
    assert projected > 0
    np.testing.assert_allclose(out, gold[f"upper_retr_n{n}"], rtol=1e-9, atol=1e-11)
    out, projected = hostcheck.rsgd(variant, "upper", n, z, u, 1e-3)
    assert projected == 0
    np.testing.assert_allclose(out, gold[f"upper_retr_small_n{n}"], rtol=1e-12, atol=1e-13)
    # bounded: the gradient is NOT symmetric here - sym(A u A) differs from sym(A sym(u) A)
    lr = 1e-4
    out, projected = hostcheck.rsgd(variant, "bounded", n, gold[f"bounded_z_n{n}"], u, lr)
    assert projected == 0
    ref = _bounded_small_step(gold, n, lr)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    wrong = sym(gold[f"bounded_z_n{n}"]) - lr * sym(gold[f"bounded_egrad2rgrad_sym_n{n}"])
    assert np.abs(wrong - ref).max() > 1e-8      # the test can tell the two apart


@pytest.mark.gpu
@pytest.mark.parametrize("n", NS)
def test_fused_rsgd_kernel_matches_reference(gold, n):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from sympa_b200 import ops
    z, u = T(gold[f"upper_z_n{n}"]).cuda(), T(gold[f"upper_u_n{n}"]).cuda()
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = ops.rsgd_step("upper", z.clone(), u, 2.5, projected=counter)
    assert int(counter.item()) > 0
    np.testing.assert_allclose(out.cpu().numpy(), gold[f"upper_retr_n{n}"], rtol=1e-9, atol=1e-11)
    out = ops.rsgd_step("upper", z.clone(), u, 1e-3)
    np.testing.assert_allclose(out.cpu().numpy(), gold[f"upper_retr_small_n{n}"], rtol=1e-12, atol=1e-13)
    lr = 1e-4
    out = ops.rsgd_step("bounded", T(gold[f"bounded_z_n{n}"]).cuda(), u, lr).cpu().numpy()
    ref = _bounded_small_step(gold, n, lr)
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    ops.check_status()

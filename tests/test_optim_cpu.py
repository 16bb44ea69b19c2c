"""Riemannian SGD (host-side torch code; CPU).  The sparse-row step must be equivalent to the dense
step, and both must match the oracle's restatement of projx / egrad2rgrad (oracle/siegel_oracle.py,
which follows upper_half.py:25-66, bounded_domain.py:41-84)."""
import pytest
import torch

import siegel_oracle as so
from sympa_b200 import BoundedDomainManifold, UpperHalfManifold
from sympa_b200.embeddings import ManifoldParameter
from sympa_b200.graphs import balanced_tree_triplets, grid_triplets
from sympa_b200.optim import RiemannianSGD


def make(kind, rows, n, seed):
    g = torch.Generator().manual_seed(seed)
    table = so.upper_spread(rows, n, generator=g, scale=0.3)
    if kind == "bounded":
        table = so.to_symmetric(so.cayley_transform(table))
    grad = torch.zeros_like(table)
    touched = torch.randperm(rows, generator=g)[: rows // 3]
    gr = torch.randn(len(touched), 2, n, n, dtype=torch.float64, generator=g)
    grad[touched] = 0.5 * (gr + gr.transpose(-1, -2))
    return table, grad, touched


@pytest.mark.parametrize("kind", ["upper", "bounded"])
@pytest.mark.parametrize("lr", [1e-2, 5.0])      # 5.0 throws points out of the manifold -> projx acts
def test_sparse_step_equals_dense_step_and_oracle(kind, lr):
    n, rows = 3, 30
    man = (UpperHalfManifold if kind == "upper" else BoundedDomainManifold)(dims=n)
    table, grad, touched = make(kind, rows, n, 7)
    out = {}
    for sparse in (False, True):
        p = ManifoldParameter(table.clone(), manifold=man)
        p.grad = grad.clone()
        RiemannianSGD([p], lr=lr, sparse_rows=sparse).step()
        out[sparse] = p.detach().clone()
    untouched = torch.ones(rows, dtype=torch.bool)
    untouched[touched] = False
    assert torch.equal(out[True][untouched], table[untouched])       # sparse: untouched rows bit-identical
    torch.testing.assert_close(out[True], out[False], rtol=1e-12, atol=1e-14)
    ref = so.rsgd_step(kind, table, grad, lr)
    torch.testing.assert_close(out[False], ref, rtol=1e-9, atol=1e-12)
    for i in range(rows):
        assert man.check_point_on_manifold(out[False][i])
    if lr > 1:
        assert man.projected_points > 0


def test_euclidean_parameters_take_plain_sgd():
    w = torch.nn.Parameter(torch.ones(1, 4, dtype=torch.float64))
    w.grad = torch.full_like(w, 2.0)
    RiemannianSGD([w], lr=0.1).step()
    assert torch.allclose(w, torch.full_like(w, 0.8))


def test_graph_triplets_match_known_sizes():
    idx, d, n = grid_triplets(20, 2)            # BASELINE config 1: 400 nodes, 79 800 pairs
    assert n == 400 and idx.shape == (79800, 2) and d.min() == 1 and d.max() == 38
    idx, d, n = balanced_tree_triplets(3, 5)    # BASELINE config 2: 364 nodes, 66 066 pairs
    assert n == 364 and idx.shape == (66066, 2) and d.min() == 1 and d.max() == 10
    # siblings 1 and 2 are at distance 2, root to a leaf at distance 5
    lookup = {(int(a), int(b)): float(x) for (a, b), x in zip(idx.tolist(), d.tolist())}
    assert lookup[(1, 2)] == 2.0 and lookup[(0, 363)] == 5.0 and lookup[(1, 4)] == 1.0


@pytest.mark.parametrize("n", [2, 3, 4, 6, 10])
@pytest.mark.parametrize("lr", [1e-2, 5.0])
def test_row_update_templates_match_oracle_upper(hostcheck, n, lr):
    """the per-row C++ templates behind sympa_rsgd_step (egrad2rgrad + retr + projx fused) against the
    oracle's restatement of upper_half.py:25-66"""
    table, grad, touched = make("upper", 40, n, 11 + n)
    ref = so.rsgd_step("upper", table, grad, lr)
    moved_ref = int((~torch.all(torch.linalg.eigvalsh((table - lr * so.upper_egrad2rgrad(table, grad))[:, 1]) > 1e-5, dim=-1)).sum())
    for variant in ((0, 1) if n <= 6 else (1,)):
        out, projected = hostcheck.rsgd(variant, "upper", n, table.numpy(), grad.numpy(), lr)
        torch.testing.assert_close(torch.from_numpy(out), ref, rtol=1e-9, atol=1e-11)
        assert projected == moved_ref
        if lr > 1:
            assert projected > 0


@pytest.mark.parametrize("n", [2, 4, 7])
def test_row_update_templates_match_oracle_spd(hostcheck, n):
    g = torch.Generator().manual_seed(3)
    x = so.spd_spread(25, n, generator=g)
    gr = torch.randn(25, n, n, dtype=torch.float64, generator=g)
    gr = 0.5 * (gr + gr.transpose(-1, -2))
    lr = 0.05
    u = -lr * (x @ gr @ x)
    ref = so.sym(x + u + 0.5 * u @ torch.linalg.solve(x, u))
    for variant in ((0, 1) if n <= 6 else (1,)):
        out, _ = hostcheck.rsgd(variant, "spd", n, x.numpy(), gr.numpy(), lr)
        torch.testing.assert_close(torch.from_numpy(out), ref, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("n", [2, 3, 4, 6, 10])
@pytest.mark.parametrize("lr", [1e-2, 5.0])
def test_row_update_templates_match_oracle_bounded(hostcheck, n, lr):
    """bounded-domain row update behind sympa_rsgd_step (A G A with A = I - conj(Z) Z, then the Takagi values
    above 1 - eps clamped) against the oracle's restatement of bounded_domain.py:41-84 (projx with a
    factoriser that returns eigenvectors)."""
    table, grad, touched = make("bounded", 40, n, 21 + n)
    ref = so.rsgd_step("bounded", table, grad, lr)
    moved_ref = int((~torch.isclose(ref, so.to_symmetric(table - lr * so.bounded_egrad2rgrad(table, grad)), rtol=0, atol=1e-13)
                     .reshape(len(table), -1).all(dim=1)).sum())
    for variant in ((0, 1) if n <= 6 else (1,)):
        out, projected = hostcheck.rsgd(variant, "bounded", n, table.numpy(), grad.numpy(), lr)
        torch.testing.assert_close(torch.from_numpy(out), ref, rtol=1e-9, atol=1e-11)
        assert projected == moved_ref
        if lr > 1:
            assert projected > 0
        man = BoundedDomainManifold(dims=n)
        for i in range(len(table)):
            assert man.check_point_on_manifold(torch.from_numpy(out[i]))

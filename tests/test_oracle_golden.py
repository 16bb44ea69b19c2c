"""The oracle (oracle/siegel_oracle.py) against (a) the golden vectors produced by the unmodified
reference (oracle/gen_golden.py) and (b) the reference's own known-answer tests."""
import numpy as np
import pytest
import torch

import siegel_oracle as so
from conftest import METRICS, golden_files, load_golden

T = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_matches_reference_golden(path):
    kind, n, regime, r = load_golden(path)
    z1, z2, go = T(r["z1"]), T(r["z2"]), T(r["go"])
    np.testing.assert_allclose(so.vvd(kind, z1, z2).numpy(), r["vvd"], rtol=1e-12, atol=1e-15)
    for m in METRICS:
        w = T(r["wsum_w"]) if m == "wsum" else None
        d, g1, g2, gw = so.dist_and_grads(kind, z1, z2, m, w, go)
        np.testing.assert_allclose(d.numpy(), r["dist_" + m], rtol=1e-12, atol=1e-15)
        scale = np.abs(r["g1_" + m]).max()
        np.testing.assert_allclose(g1.numpy(), r["g1_" + m], rtol=0, atol=1e-9 * scale)
        np.testing.assert_allclose(g2.numpy(), r["g2_" + m], rtol=0, atol=1e-9 * scale)
        if m == "wsum":
            np.testing.assert_allclose(gw.numpy(), r["gw_wsum"], rtol=1e-10)


def test_building_blocks_match_reference():
    r = np.load(golden_files("blocks.npz")[0])
    for n in (2, 3, 4):
        z = T(r[f"z_n{n}"])
        w = so.cayley_transform(z)
        np.testing.assert_allclose(w.numpy(), r[f"cayley_n{n}"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(so.inverse_cayley_transform(so.to_symmetric(w)).numpy(), r[f"icayley_n{n}"],
                                   rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(so.cinverse(z).numpy(), r[f"inverse_n{n}"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(so.takagi_values(so.to_symmetric(w)).numpy(), r[f"takagi_n{n}"], rtol=1e-12,
                                   atol=1e-14)
        np.testing.assert_allclose(so.matrix_sqrt(so.im(z)).numpy(), r[f"msqrt_n{n}"], rtol=1e-12, atol=1e-14)
        d, s = so.takagi_factorize(so.to_symmetric(w))
        # Z = conj(S) D S^H  (tests/test_takagi_factorization.py:15-126 property)
        sc = torch.complex(s[:, 0], s[:, 1])
        rec = sc.conj() @ torch.diag_embed(d).to(sc.dtype) @ sc.transpose(-1, -2).conj()
        wc = torch.complex(w[:, 0], w[:, 1])
        assert (rec - wc).abs().max() < 1e-12


def test_reference_known_answers_bmm():
    """values of /root/reference/tests/test_math.py:175-225"""
    x = so.cplx(T([[[1, -3], [5, -7]]]), T([[[9, -11], [-14, 15]]]))
    y = so.cplx(T([[[9, -11], [-14, 15]]]), T([[[1, -3], [5, -7]]]))
    z = so.cplx(T([[[-3, -1], [-2, 5]]]), T([[[-1, 3], [0, -2]]]))
    out = so.cmatmul(x, y)
    assert torch.equal(out[:, 0], T([[[97, -106], [82, -97]]]))
    assert torch.equal(out[:, 1], T([[[221, -246], [-366, 413]]]))
    out3 = so.cmatmul3(x, y, z)
    assert torch.equal(out3[:, 0], T([[[142, -1782], [-418, 1357]]]))
    assert torch.equal(out3[:, 1], T([[[-268, -948], [190, 2871]]]))


def test_reference_known_answers_inverse():
    """values of /root/reference/tests/test_math.py:227-287"""
    x = so.cplx(T([[[1, -3], [-3, 7]]]), T([[[-9, 11], [11, 15]]]))
    er = T([[[256 / 8105, 141 / 16210], [141 / 16210, 23 / 16210]]])
    ei = T([[[921 / 16210, -356 / 8105], [-356 / 8105, -288 / 8105]]])
    out = so.cinverse(x)
    np.testing.assert_allclose(out[:, 0].numpy(), er.numpy(), rtol=1e-12)
    np.testing.assert_allclose(out[:, 1].numpy(), ei.numpy(), rtol=1e-12)
    x = so.cplx(T([[[-1, -3, 9], [3, 5, 7], [2, 9, 11]]]), T([[[9, 4, -6], [-4, 7, 9], [-2, 7, -3]]]))
    er = T([[[951 / 16589, 3223 / 16589, -4496 / 16589], [5029 / 66356, 2486 / 16589, 1387 / 33178],
             [2137 / 66356, 1030 / 16589, -1828 / 16589]]])
    ei = T([[[-3143 / 16589, -3622 / 16589, 1532 / 16589], [5533 / 66356, 6925 / 33178, -17977 / 66356],
             [-4167 / 66356, -5779 / 33178, 6565 / 66356]]])
    out = so.cinverse(x)
    np.testing.assert_allclose(out[:, 0].numpy(), er.numpy(), rtol=1e-11)
    np.testing.assert_allclose(out[:, 1].numpy(), ei.numpy(), rtol=1e-11)


@pytest.mark.parametrize("kind", ["upper", "bounded"])
def test_reference_metamorphic_properties(kind):
    """dist symmetry and dist(x, x) = 0 (tests/test_upper_half.py:109-186, tests/test_bounded_domain.py:95-127)"""
    g = torch.Generator().manual_seed(42)
    gen = so.upper_random if kind == "upper" else so.bounded_random
    x, y = gen(10, 3, generator=g), gen(10, 3, generator=g)
    x, y = so.to_symmetric(x), so.to_symmetric(y)
    assert torch.allclose(so.dist(kind, x, y), so.dist(kind, y, x), rtol=1e-5, atol=1e-8)
    dxx = so.dist(kind, x, x)
    assert torch.allclose(dxx, torch.zeros_like(dxx), atol=1e-8)


def test_fmin_weights_are_ascending_even_numbers():
    assert so.fmin_weights(4).tolist() == [0.0, 2.0, 4.0, 6.0]

"""oracle/kernel_model.py (numpy statement of the algorithm the CUDA kernels use: Cholesky
congruence, real-arithmetic complex inverse, one-sided Jacobi, analytic backward) against the
reference golden vectors."""
import numpy as np
import pytest

import kernel_model as km
from conftest import METRICS, golden_files, grad_tolerance, load_golden, sym


@pytest.mark.parametrize("path", [p for p in golden_files() if "_n10_" not in p and "_n6_" not in p],
                         ids=lambda p: p.split("/")[-1][:-4])
def test_kernel_model_matches_reference(path):
    kind, n, regime, r = load_golden(path)
    fn = km.upper_pair if kind == "upper" else km.bounded_pair
    for m in METRICS:
        w = r["wsum_w"] if m == "wsum" else None
        gmax = max(np.abs(r["g1_" + m]).max(), np.abs(r["g2_" + m]).max())
        for b in range(0, r["z1"].shape[0], 3):
            z1, z2 = r["z1"][b], r["z2"][b]
            vv, dist, grads, gw = fn(z1[0], z1[1], z2[0], z2[1], m, w)
            np.testing.assert_allclose(vv, r["vvd"][b], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(dist, r["dist_" + m][b], rtol=1e-9)
            g1, g2 = sym(r["g1_" + m][b]), sym(r["g2_" + m][b])
            got = np.stack(grads) * r["go"][b]
            ref = np.stack([g1[0], g1[1], g2[0], g2[1]])
            assert np.abs(got - ref).max() <= grad_tolerance(n) * gmax

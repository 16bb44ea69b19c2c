"""CPU tests added in round 2: the float64 wsum weights (ADVICE), the Cartesian-product graph generator of
BASELINE config 3 against networkx, an independent check of the spd distance restatement (scipy logm + mpmath),
and the rank-invariance of the bounded-domain route choice under a 2-rank gloo group."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import siegel_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wsum_weights_are_float64_without_touching_the_global_default():
    from sympa_b200 import BoundedDomainManifold, MetricType, UpperHalfManifold
    assert torch.get_default_dtype() == torch.float32        # this package never changes it
    for cls in (UpperHalfManifold, BoundedDomainManifold):
        man = cls(dims=5, metric=MetricType.WEIGHTED_SUM)
        assert man.metric.weights.dtype == torch.float64 and man.metric.weights.shape == (1, 5)
        assert any(p is man.metric.weights for p in man.parameters())     # optimised as in the reference


def test_product_cartesian_graph_matches_networkx():
    nx = pytest.importorskip("networkx")
    from sympa_b200.graphs import product_cartesian_triplets
    idx, d, n = product_cartesian_triplets(2, 3, 3, 2)          # tree(2,3) x grid 3x3: 15 * 9 = 135 nodes
    g = nx.cartesian_product(nx.balanced_tree(2, 3), nx.grid_graph(dim=[3, 3]))     # preprocess.py:53-60
    g = nx.convert_node_labels_to_integers(g, ordering="sorted")                     # preprocess.py:151-152
    sp = dict(nx.all_pairs_shortest_path_length(g))
    assert n == len(g) == 135 and idx.shape == (n * (n - 1) // 2, 2)
    ref = np.array([sp[int(a)][int(b)] for a, b in idx.tolist()], dtype=np.float64)
    assert np.array_equal(d.numpy(), ref) and d.min() >= 1
    # the reference's CLI defaults (tree 3/3, nodes 125, grid_dims 3): 5000 nodes
    _, _, n_default = product_cartesian_triplets(3, 3, 5, 3) if os.environ.get("SYMPA_SLOW") else (None, None, 40 * 125)
    assert n_default == 5000


@pytest.mark.parametrize("n", [2, 3, 5, 10])
def test_spd_distance_restatement_against_independent_routes(n):
    """geoopt is absent (parity with it stays UNPINNED); this pins the restated formula itself: the affine-invariant
    distance ||logm(X^-1/2 Y X^-1/2)||_F computed (a) by the oracle's eigendecompositions, (b) by scipy's Schur
    logm of the non-symmetric X^-1 Y (sqrt of sum log^2 of its eigenvalues), (c) in 40-digit mpmath."""
    scipy_linalg = pytest.importorskip("scipy.linalg")
    mp_ = pytest.importorskip("mpmath")
    g = torch.Generator().manual_seed(5 + n)
    x, y = so.spd_spread(4, n, generator=g), so.spd_spread(4, n, generator=g)
    d = so.spd_dist(x, y).numpy()
    for k in range(4):
        lg = scipy_linalg.logm(np.linalg.solve(x[k].numpy(), y[k].numpy()))
        ev = np.linalg.eigvals(lg)
        assert abs(np.sqrt((ev.real ** 2).sum()) - d[k]) <= 1e-9 * d[k] and np.abs(ev.imag).max() < 1e-9
    mp_.mp.dps = 40
    xm, ym = mp_.matrix(x[0].tolist()), mp_.matrix(y[0].tolist())
    lam = mp_.eig(mp_.inverse(xm) * ym, left=False, right=False)
    dm = mp_.sqrt(sum(mp_.log(mp_.re(v)) ** 2 for v in lam))
    assert abs(float(dm) - d[0]) <= 1e-11 * d[0]


def _route_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from sympa_b200 import ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group(backend="gloo", init_method="env://")
    rows = 1000
    local_pairs = [900, 0][rank]                   # unequal shards: one rank dense, the other EMPTY
    res = []
    for kind, n in (("bounded", 3), ("bounded", 7), ("bounded", 8), ("upper", 4)):
        mine = ops.bounded_by_rows(kind, n, local_pairs, rows, sync_grad=True)
        both = [None, None]
        dist.all_gather_object(both, mine)
        res.append((both[0] == both[1], mine))
    # without the collective the local batch may decide
    alone = ops.bounded_by_rows("bounded", 3, local_pairs, rows, sync_grad=False)
    out[rank] = (res, alone)
    dist.barrier()
    dist.destroy_process_group()


def test_bounded_route_choice_is_rank_invariant_under_sync_grad():
    world, port = 2, 29400 + os.getpid() % 200
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_route_worker, args=(world, port, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert all(same for same, _ in r0[0]) and all(same for same, _ in r1[0])
    assert [v for _, v in r0[0]] == [True, True, True, False]      # bounded by rows up to n = 10 (ops.BOUNDED_BY_ROWS_MAX_N)
    assert (r0[1], r1[1]) == (True, False)       # single-process rule: 2 * pairs >= rows


def test_numa_binding_helper_parses_cpulists_and_is_inert_without_a_gpu():
    """distributed.bind_to_gpu_numa_node: sysfs cpulist syntax, and no change of the affinity when there is no GPU
    (or no readable topology) - the bench calls it unconditionally for N > 1."""
    sys.path.insert(0, ROOT)
    from sympa_b200 import distributed as sd
    assert sd._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sd._parse_cpulist("5") == {5} and sd._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert sd.bind_to_gpu_numa_node(0) is None
        assert os.sched_getaffinity(0) == before

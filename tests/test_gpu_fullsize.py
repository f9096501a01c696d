"""Parity at BASELINE.json's full sizes.  Where the oracle finishes in seconds it is the checker;
for the 10^8-row mixture the checks are size-independent properties (additivity over item ranges,
agreement with the oracle on random item windows, shard additivity) plus the oracle on a window."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-10


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))


@pytest.fixture(scope="module")
def fb():
    import torch
    assert torch.cuda.is_available()
    import fetching_b200
    return fetching_b200


def test_c4_lasso_full_size(fb, orc):
    """BASELINE config 4: n = 10^5 rows, p = 1000, FP64, items of 1000 rows, frontier of 8."""
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(100000, 1000, seed=0)
    th = models.blasso_thetas(beta, 8, seed=1)
    want = orc.blasso_frontier(X, y, th, 1000)
    with fb.FrontierEngine.blasso(X, y, item_rows=1000) as e:
        assert e.data_size == 100 and e.theta_len == 1001
        s, q = e.eval_frontier(th)
        s18, q18 = None, None
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
    # the reference's own item size (18000 rows, pyblasso.py:136-137): 5 items, 10000 trailing rows unused
    want = orc.blasso_frontier(X, y, th[:3], 18000)
    with fb.FrontierEngine.blasso(X, y) as e:
        assert e.data_size == 5
        s, q = e.eval_frontier(th[:3])
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


def test_c5_mixture_1e8_properties(fb, orc):
    """BASELINE config 5 on one GPU: K = 8, d = 2, N = 10^8 rows (1.6 GB), items of 1000 rows."""
    import bench
    prm = dict(rows=100_000_000, K=8, d=2, item_rows=1000)
    shard = bench.make_shard("mixture", prm, 0, 1)
    data = shard["data"]
    th = bench.make_thetas("mixture", prm, shard, 4)
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        assert e.data_size == 100000
        s, q = e.eval_frontier(th)
        # additivity over item ranges (what a row-sharded run relies on), uneven cut
        parts = [e.eval_frontier_range(th, a, b - a) for a, b in ((0, 37501), (37501, 99000), (99000, 100000))]
        assert rel(sum(p[0] for p in parts), s) < 1e-12 and rel(sum(p[1] for p in parts), q) < 1e-12
        # oracle on three windows of 1000 items each, incl. the very last items
        for first in (0, 54321, 99000):
            w = orc.mixture_frontier(data[first * 1000:(first + 1000) * 1000], th, 1000)
            g = e.eval_frontier_range(th, first, 1000)
            assert rel(g[0], w[0]) < TOL64 and rel(g[1], w[1]) < TOL64
        # ... and the oracle over ALL 10^8 rows (4 nodes x 10^8 rows: a few seconds on the host's threads)
        orc.set_threads(len(os.sched_getaffinity(0)))
        w = orc.mixture_frontier(data, th, 1000)
        print("C5 full data, GPU vs oracle: rel err sum %.2e sum_sq %.2e" % (rel(s, w[0]), rel(q, w[1])))
        assert rel(s, w[0]) < TOL64 and rel(q, w[1]) < TOL64
        # deterministic
        s2, q2 = e.eval_frontier(th)
        assert np.array_equal(s, s2) and np.array_equal(q, q2)
    # Cauchy-Schwarz style sanity between the two outputs: sum_sq >= sum^2 / n_items
    assert np.all(q >= s * s / 100000 * (1 - 1e-12))

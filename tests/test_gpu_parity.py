"""Parity of the CUDA path with the oracle, THROUGH THE C ABI (libpfetch.so), on a B200.

Tolerances (BASELINE.json north_star): per-node log-likelihood sums within 1e-10 relative in
FP64 mode and 1e-5 relative in FP32 mode.  The oracle (oracle/) is only the checker here.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-5


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.fixture(scope="module")
def fb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import fetching_b200
    from fetching_b200 import _lib
    _lib.load()  # must be the in-tree CUDA library; raises if it is missing
    return fetching_b200


# ------------------------------------------------------------------------------------ normal target
def test_normal_vs_reference_cpp_golden(fb, orc, ref_normal):
    """Golden values are outputs of the reference's own NormalSampler<2> (oracle/_ref/pf_ref)."""
    data = orc.normal_dataset(int(ref_normal["n"]))
    with fb.FrontierEngine.normal(data) as e:
        assert e.data_size == 10000 and e.theta_len == 5
        s, q = e.eval_frontier(ref_normal["thetas"])  # 12 identity + 8 general covariances
    assert rel(s, ref_normal["sum"]) < TOL64
    assert rel(q, ref_normal["sum_sq"]) < TOL64
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(ref_normal["thetas"][:12])  # identity fast path on its own
    assert rel(s, ref_normal["sum"][:12]) < TOL64 and rel(q, ref_normal["sum_sq"][:12]) < TOL64


@pytest.mark.parametrize("B", [1, 2, 3, 5, 8, 17, 32, 33, 64, 65, 130])
def test_normal_frontier_sizes(fb, orc, B):
    from fetching_b200 import models
    data = orc.normal_dataset(50021)  # odd size: ragged last tile
    th = models.normal_thetas(B, seed=B, center=(200.0, 100.0), scale=2.0)
    want = orc.normal_frontier(data, th)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(th)
        s2, q2 = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
    assert np.array_equal(s, s2) and np.array_equal(q, q2)  # fixed-order reductions: bit-stable


def test_normal_full_size_c1(fb, orc):
    """BASELINE config 1 shape: N = 10^6, D = 2, the normal.cc dataset, initial theta and proposals."""
    from fetching_b200 import models
    data = models.normal_dataset(1000000)
    th = models.normal_thetas(8, seed=3)
    th[0] = [0, 0, 1, 0, 1]
    want = orc.normal_frontier(data, th)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
    # the golden chain's first row is the N = 10^4 prefix of the same recipe
    with fb.FrontierEngine.normal(data[:10000]) as e:
        s, _ = e.eval_frontier(th[:1])
    assert abs(s[0] - (-250005158.44676721)) < 1e-10 * 250005158.0


@pytest.mark.parametrize("D", [1, 3, 4, 8])
def test_normal_other_dimensions(fb, orc, D):
    rng = np.random.RandomState(D)
    data = rng.normal(0, 1, (20011, D)) + np.arange(D)
    tl = D + D * (D + 1) // 2
    th = np.zeros((6, tl))
    for b in range(6):
        th[b, :D] = rng.normal(0, 1, D)
        A = rng.normal(0, 1, (D, D))
        cov = A @ A.T + D * np.eye(D) if b >= 3 else np.eye(D)
        th[b, D:] = [cov[i, j] for i in range(D) for j in range(i + 1)]
    want = orc.normal_frontier(data, th)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


def test_normal_f32_mode(fb, orc):
    from fetching_b200 import models
    data = orc.normal_dataset(100000)
    th = models.normal_thetas(9, seed=5)
    want = orc.normal_frontier(data, th)
    with fb.FrontierEngine.normal(data, dtype=fb.F32) as e:
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL32 and rel(q, want[1]) < TOL32


def test_normal_item_ranges_add_up(fb, orc):
    from fetching_b200 import models
    data = orc.normal_dataset(30000)
    th = models.normal_thetas(4, seed=2, center=(200, 100), scale=1.0)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(th)
        s1, q1 = e.eval_frontier_range(th, 0, 12345)
        s2, q2 = e.eval_frontier_range(th, 12345, 30000 - 12345)
        z, _ = e.eval_frontier_range(th, 500, 0)
    assert rel(s1 + s2, s) < 1e-12 and rel(q1 + q2, q) < 1e-12 and np.all(z == 0)
    for b in range(4):
        w = orc.normal_eval(data, th[b], 1, 0, 0, 12345)
        assert rel(s1[b], w[0]) < TOL64 and rel(q1[b], w[1]) < TOL64


def test_normal_ordered_partial_ranges(fb, orc):
    """a3/a9: partial sums in the reference's per-node StrideIndex order (strideindex.hh:46-48), the
    updates a worker sends every batch_size items (worker.hh:141-162)."""
    from fetching_b200 import models
    n = 10000
    data = orc.normal_dataset(n)
    th = models.normal_thetas(5, seed=4, center=(200, 100), scale=1.0)
    th[4, 2:] = [2.0, 0.3, 1.5]  # one general covariance
    order = np.array([orc.normal_prepare_draws(n, 2 * b)[:2] for b in range(5)], dtype=np.uint64)
    with fb.FrontierEngine.normal(data) as e:
        run_s, run_q = np.zeros(5), np.zeros(5)
        for first in range(0, n, 1000):  # ten cumulative updates, like -b 1000
            s, q = e.eval_frontier_range(th, first, 1000, order=order)
            for b in range(5):
                w = orc.normal_eval(data, th[b], int(order[b, 0]), int(order[b, 1]), first, 1000)
                assert rel(s[b], w[0]) < TOL64 and rel(q[b], w[1]) < TOL64
            run_s += s
            run_q += q
        full_s, full_q = e.eval_frontier(th)
        assert rel(run_s, full_s) < 1e-12 and rel(run_q, full_q) < 1e-12  # a permutation: same totals
        with pytest.raises(fb.PfetchError, match="PF_ERR_SHAPE"):
            e.eval_frontier_range(th, 0, 10, order=np.array([[0, 0]] * 5, dtype=np.uint64))
    # the golden values of the reference binary are sums in exactly this order
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_normal_eval.npz"))
    st, off, _ = orc.normal_prepare_draws(n, 0)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier_range(g["thetas"], 0, n, order=np.array([[st, off]] * len(g["thetas"]), dtype=np.uint64))
    assert rel(s, g["sum"]) < TOL64 and rel(q, g["sum_sq"]) < TOL64


# ------------------------------------------------------------------------------------ mixture
def test_mixture_vs_reference_python_golden(fb, py_models):
    """Golden values are outputs of the reference's own Mixture.evaluate (K = 8, d = 8)."""
    g = py_models
    data, b = g["mix_data"], int(g["mix_item_rows"])
    th = g["mix_thetas"].reshape(len(g["mix_thetas"]), -1)
    with fb.FrontierEngine.mixture(data, 8, b) as e:
        assert e.data_size == int(g["mix_n_items"]) and e.theta_len == 64
        s, q = e.eval_frontier(th)
    assert rel(s, g["mix_item_lp"].sum(axis=1)) < TOL64
    assert rel(q, (g["mix_item_lp"] ** 2).sum(axis=1)) < TOL64
    assert rel(s, g["mix_log_posterior"]) < TOL64


@pytest.mark.parametrize("item_rows", [1000, 250, 1017, 7, 100000])
@pytest.mark.parametrize("B", [1, 8, 33])
def test_mixture_k8_d2_items(fb, orc, item_rows, B):
    """BASELINE shape K = 8, d = 2; item sizes that do / do not line up with tiles and CTAs."""
    from fetching_b200 import models
    n = 100000 if item_rows != 7 else 20000
    data, pos = models.mixture_dataset(n, 8, 2, seed=4)
    th = models.mixture_thetas(pos, B, seed=B)
    want = orc.mixture_frontier(data, th, item_rows)
    with fb.FrontierEngine.mixture(data, 8, item_rows) as e:
        assert e.data_size == n // item_rows
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


@pytest.mark.parametrize("K,d", [(5, 3), (3, 1), (16, 4), (2, 8), (8, 8), (4, 2), (16, 2), (8, 4), (2, 2), (3, 2), (5, 2), (6, 2)])
def test_mixture_other_shapes(fb, orc, K, d):
    from fetching_b200 import models
    data, pos = models.mixture_dataset(30000, K, d, seed=K + d)
    th = models.mixture_thetas(pos, 5, seed=1, scale=0.3)
    want = orc.mixture_frontier(data, th, 500)
    with fb.FrontierEngine.mixture(data, K, 500) as e:
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


def test_mixture_full_size_c3(fb, orc):
    """BASELINE config 3: K = 8, d = 2, N = 10^6, items of 1000 rows, a depth-3 frontier."""
    from fetching_b200 import models
    data, pos = models.mixture_dataset(1000000, 8, 2, seed=0)
    th = models.mixture_thetas(pos, 8, seed=1)
    want = orc.mixture_frontier(data, th, 1000)
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        assert e.data_size == 1000
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


def test_mixture_f32_mode(fb, orc):
    from fetching_b200 import models
    data, pos = models.mixture_dataset(200000, 8, 2, seed=2)
    th = models.mixture_thetas(pos, 8, seed=3)
    want = orc.mixture_frontier(data, th, 1000)
    with fb.FrontierEngine.mixture(data, 8, 1000, dtype=fb.F32) as e:
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL32 and rel(q, want[1]) < TOL32


def test_mixture_edge_cases(fb, orc):
    from fetching_b200 import models
    data, pos = models.mixture_dataset(10000, 8, 2, seed=9)
    # ragged: 10000 rows, items of 3000 -> 3 whole items, trailing 1000 rows never scored
    th = models.mixture_thetas(pos, 3, seed=1)
    want = orc.mixture_frontier(data[:9000], th, 3000)
    with fb.FrontierEngine.mixture(data, 8, 3000) as e:
        assert e.data_size == 3
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
    # underflow: every component infinitely far -> log(0) = -inf, as numpy gives the reference
    far = np.full((1, 16), 60.0)
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        s, q = e.eval_frontier(far)
    assert s[0] == -np.inf and orc.mixture_item(data, far[0], 1000, 0) == -np.inf


# ------------------------------------------------------------------------------------ lasso
def test_mixture_factored_and_clamped_forms_agree_with_the_oracle(fb, orc):
    """The FP64 mixture kernel scores a frontier with the factored exponent z_k = A_k + B_k.x when every node's theta
    passes the range test against the data's bounding box (|A_k| + sum_j |B_kj| max|x_j| <= 512) and with the clamped
    squared-distance form otherwise; the switch is per call, for the whole frontier.  Both sides of it, the frontier
    that straddles it, and thetas at the edge of the range, against the oracle."""
    from fetching_b200 import models
    data, pos = models.mixture_dataset(40000, 8, 2, seed=11)
    near = models.mixture_thetas(pos, 6, seed=1)                      # |theta| <= ~2.2: factored
    far = near.copy()
    far[:, 0:2] += 40.0                                               # one component 40 away: |A_k| ~ 2300, clamped form
    edge = near.copy()
    edge[:, 2:4] = [12.0, -12.0]                                      # |A_k| = 208, bound ~ 430: still factored
    over = near.copy()
    over[:, 2:4] = [15.0, -15.0]                                      # |A_k| = 325, bound ~ 560: clamped
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        for name, th in (("near", near), ("far", far), ("edge", edge), ("over", over),
                         ("mixed", np.concatenate([near[:3], far[:2], edge[:1]]))):
            s, q = e.eval_frontier(th)
            want = orc.mixture_frontier(data, th, 1000)
            assert rel(s, want[0]) < 1e-12 and rel(q, want[1]) < 1e-12, name
        # the same node scored inside a frontier that passes the test and inside one that does not: two code paths,
        # one number up to rounding
        s_f, _ = e.eval_frontier(near)
        s_c, _ = e.eval_frontier(np.concatenate([near, far[:1]]))
        assert rel(s_c[:6], s_f) < 1e-13
    # d = 8, K = 8 (the reference model's own shape) through both forms
    data8, pos8 = models.mixture_dataset(20000, 8, 8, seed=12)
    th8 = models.mixture_thetas(pos8, 5, seed=2)
    far8 = th8.copy()
    far8[:, :8] += 12.0                                               # |theta_0|^2 ~ 1150: clamped form
    with fb.FrontierEngine.mixture(data8, 8, 500) as e:
        for th in (th8, far8):
            s, q = e.eval_frontier(th)
            want = orc.mixture_frontier(data8, th, 500)
            assert rel(s, want[0]) < 1e-12 and rel(q, want[1]) < 1e-12


def test_blasso_vs_reference_python_golden(fb, py_models):
    """Golden values are outputs of the reference's own BLasso.evaluate (18000-row items)."""
    g = py_models
    with fb.FrontierEngine.blasso(g["bl_X"], g["bl_y"]) as e:
        assert e.data_size == int(g["bl_n_items"]) and e.theta_len == g["bl_X"].shape[1] + 1
        s, q = e.eval_frontier(g["bl_thetas"])
    assert rel(s, g["bl_item_lp"].sum(axis=1)) < TOL64
    assert rel(q, (g["bl_item_lp"] ** 2).sum(axis=1)) < TOL64


@pytest.mark.parametrize("B", [1, 8, 9, 16, 17, 33, 64, 70])
def test_blasso_frontier_sizes(fb, orc, B):
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(6000, 100, seed=1)
    th = models.blasso_thetas(beta, B, seed=B, scale=0.05)
    want = orc.blasso_frontier(X, y, th, 500)
    with fb.FrontierEngine.blasso(X, y, item_rows=500) as e:
        s, q = e.eval_frontier(th)
        s2, q2 = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
    assert np.array_equal(s, s2) and np.array_equal(q, q2)


@pytest.mark.parametrize("n,p,item_rows", [(4096, 1000, 1000), (5000, 37, 333), (20000, 16, 18000), (777, 3, 100)])
def test_blasso_shapes(fb, orc, n, p, item_rows):
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(n, p, seed=p)
    th = models.blasso_thetas(beta, 6, seed=2, scale=0.02)
    want = orc.blasso_frontier(X, y, th, item_rows)
    with fb.FrontierEngine.blasso(X, y, item_rows=item_rows) as e:
        assert e.data_size == n // item_rows
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


@pytest.mark.parametrize("n,p,item_rows,B", [(6000, 100, 500, 16), (5000, 37, 333, 9), (20000, 16, 18000, 64),
                                              (777, 3, 100, 24), (4096, 1000, 1000, 40), (6000, 100, 500, 3),
                                              (10000, 250, 1000, 8), (9000, 130, 1000, 20), (7000, 64, 700, 47),
                                              (36000, 96, 18000, 50)])  # (node blocks of 8: 3, 6 and 7 of them)
def test_blasso_tensor_variant_matches_vector_variant(fb, orc, n, p, item_rows, B):
    """FP64 runs the contraction on the FP64 tensor cores (DMMA m8n8k4); PF_OPT_MATVEC_TENSOR = 0 selects the
    CUDA-core FMA variant.  Both are exact IEEE FP64 FMAs, only the summation order over columns differs."""
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(n, p, seed=p)
    th = models.blasso_thetas(beta, B, seed=B, scale=0.03)
    want = orc.blasso_frontier(X, y, th, item_rows)
    with fb.FrontierEngine.blasso(X, y, item_rows=item_rows) as e:
        st, qt = e.eval_frontier(th)
        first = 1 if e.data_size > 1 else 0
        pt = e.eval_frontier_range(th, first, max(1, e.data_size // 2))  # a partial item range through the same path
        e.set_option(fb.OPT_MATVEC_TENSOR, 0)
        sv, qv = e.eval_frontier(th)
        pv = e.eval_frontier_range(th, first, max(1, e.data_size // 2))
    assert rel(st, want[0]) < TOL64 and rel(qt, want[1]) < TOL64
    assert rel(st, sv) < 1e-12 and rel(qt, qv) < 1e-12
    assert rel(pt[0], pv[0]) < 1e-12 and rel(pt[1], pv[1]) < 1e-12


@pytest.mark.parametrize("n,p,item_rows,B", [(6000, 100, 500, 16), (5000, 37, 333, 9), (20000, 16, 18000, 64),
                                              (777, 3, 100, 24), (4096, 1000, 1000, 40), (6000, 100, 500, 3),
                                              (1000, 64, 50, 1), (30000, 257, 1000, 8),
                                              # more row tiles per CTA than TMEM has accumulators for: several groups
                                              (200000, 33, 18000, 64), (150000, 8, 100, 64), (260000, 40, 7000, 40)])
def test_blasso_f32_tcgen05_contraction_is_fp32_faithful(fb, orc, n, p, item_rows, B):
    """PF_F32 lasso: the contraction runs on the tcgen05 tensor cores as 3xTF32 (hi.hi + lo.hi + hi.lo, FP32 accumulate
    in TMEM).  Against the oracle on the float-rounded data the error must look like FP32 arithmetic, not like tf32
    (2^-10): well inside the 1e-5 contract; and it must agree with the CUDA-core FP32 variant to the same level.
    Shapes cover items shorter than a 128-row tile, items straddling tiles, p not a multiple of 32, every N padding, and
    CTAs with more tiles than accumulator slots (the chunk loop then runs once per group of tiles)."""
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(n, p, seed=p)
    th = models.blasso_thetas(beta, B, seed=2, scale=0.02)
    X32 = X.astype(np.float32).astype(np.float64)
    want = orc.blasso_frontier(X32, y, th, item_rows)
    with fb.FrontierEngine.blasso(X, y, item_rows=item_rows, dtype=fb.F32) as e:
        e.set_option(fb.OPT_MATVEC_TENSOR, 2)   # the tcgen05 path for every B (by default only above 32 nodes)
        s, q = e.eval_frontier(th)
        s2, q2 = e.eval_frontier(th)
        n_items = e.data_size
        parts = [e.eval_frontier_range(th, a, b - a) for a, b in ((0, n_items // 3), (n_items // 3, n_items))]
        e.set_option(fb.OPT_MATVEC_TENSOR, 0)
        sv, qv = e.eval_frontier(th)
    assert np.array_equal(s, s2) and np.array_equal(q, q2)                       # bit-stable
    print("tcgen05 3xTF32 vs oracle: %.2e / %.2e; CUDA-core FP32 vs oracle: %.2e" % (rel(s, want[0]), rel(q, want[1]), rel(sv, want[0])))
    assert rel(s, want[0]) < 8e-6 and rel(q, want[1]) < 1e-5                     # (tf32 alone would be ~1e-3)
    assert rel(s, sv) < 8e-6 and rel(q, qv) < 1e-5
    assert rel(parts[0][0] + parts[1][0], s) < 1e-12 and rel(parts[0][1] + parts[1][1], q) < 1e-12


def test_blasso_f32_mode(fb, orc):
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(8000, 200, seed=3)
    th = models.blasso_thetas(beta, 8, seed=1, scale=0.02)
    want = orc.blasso_frontier(X, y, th, 1000)
    with fb.FrontierEngine.blasso(X, y, item_rows=1000, dtype=fb.F32) as e:
        s, q = e.eval_frontier(th)
    assert rel(s, want[0]) < TOL32 and rel(q, want[1]) < TOL32


# ------------------------------------------------------------------------------------ Boltzmann
def test_boltzmann_scalar_vs_reference_cpp_golden(fb, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_boltz_eval.npz"))
    with fb.FrontierEngine.boltzmann_scalar(int(g["count"])) as e:
        assert e.data_size == 10000
        s, q = e.eval_frontier(g["thetas"])
    assert rel(s, g["sum"]) < 1e-14 and rel(q, g["sum_sq"]) < 1e-14


@pytest.mark.parametrize("D,B", [(1024, 64), (64, 5), (200, 17), (1000, 33)])
def test_boltzmann_dense(fb, orc, D, B):
    """BASELINE config 2 (D = 1024, B = 64) and odd shapes.  New workload: parity is against the
    oracle's definition -x'Wx/T (no reference implementation exists)."""
    from fetching_b200 import models
    W, X = models.boltzmann_dataset(D, B, seed=D)
    want = orc.boltzmann_dense_frontier(W, X, 2.0)
    with fb.FrontierEngine.boltzmann_dense(W, temperature=2.0) as e:
        assert e.data_size == 1 and e.theta_len == D
        s, q = e.eval_frontier(X)
    # x'Wx cancels heavily (|E| << sum |terms|): compare on the scale of the terms
    scale = np.array([np.abs(np.outer(x, x) * W).sum() / 2.0 for x in X])
    assert np.max(np.abs(s - want[0]) / scale) < TOL64
    assert np.max(np.abs(q - s * s) / np.maximum(s * s, 1e-300)) < 1e-14
    # ... and against the independent statement (symmetric half, pairwise, 80-bit): two unrelated checkers
    ind = np.array([float(orc.boltzmann_dense_logp_independent(W, x, 2.0)) for x in X])
    assert np.max(np.abs(s - ind) / scale) < TOL64
    if True:  # tensor-core (DMMA) variant above, CUDA-core variant here
        with fb.FrontierEngine.boltzmann_dense(W, temperature=2.0) as e:
            e.set_option(fb.OPT_MATVEC_TENSOR, 0)
            sv, _ = e.eval_frontier(X)
        assert np.max(np.abs(s - sv) / scale) < 1e-13


def test_non_finite_inputs_propagate_like_the_reference(fb, orc):
    """A NaN row poisons every node's sums (as numpy / the C++ loop would); an infinitely distant point gives
    log(0) = -inf for the mixture and -inf for the normal target; finite nodes elsewhere are untouched."""
    from fetching_b200 import models
    data, pos = models.mixture_dataset(5000, 8, 2, seed=1)
    th = models.mixture_thetas(pos, 3, seed=1)
    bad = data.copy()
    bad[1234, 1] = np.nan
    with fb.FrontierEngine.mixture(bad, 8, 1000) as e:
        s, q = e.eval_frontier(th)
        s_ok, _ = e.eval_frontier_range(th, 2, 3)          # items 2..4 do not contain row 1234
    assert np.all(np.isnan(s)) and np.all(np.isnan(q))
    want = orc.mixture_frontier(data[2000:5000], th, 1000)
    assert rel(s_ok, want[0]) < TOL64
    far = data.copy()
    far[77] = [np.inf, 0.0]
    with fb.FrontierEngine.mixture(far, 8, 1000) as e:
        s, _ = e.eval_frontier(th)
    assert np.all(s == -np.inf) and orc.mixture_item(far, th[0], 1000, 0) == -np.inf
    nd = orc.normal_dataset(4000)
    nd[5, 0] = np.nan
    with fb.FrontierEngine.normal(nd) as e:
        s, q = e.eval_frontier(models.normal_thetas(2, seed=1))
    assert np.all(np.isnan(s)) and np.all(np.isnan(q))
    X, y, beta = models.blasso_dataset(3000, 10, seed=1)
    X[100, 3] = np.nan
    with fb.FrontierEngine.blasso(X, y, item_rows=1000) as e:
        s, _ = e.eval_frontier(models.blasso_thetas(beta, 2, seed=1))
    assert np.all(np.isnan(s))


# ------------------------------------------------------------------------------------ boundary behaviour
def test_device_pointer_path_matches_host_path(fb, orc):
    import torch
    from fetching_b200 import models
    data, pos = models.mixture_dataset(50000, 8, 2, seed=1)
    th = models.mixture_thetas(pos, 8, seed=2)
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        s, q = e.eval_frontier(th)
        d_th = torch.from_numpy(th).cuda()
        d_out = torch.empty(2, 8, dtype=torch.float64, device="cuda")
        launches = e.launch_count
        e.eval_frontier_device(d_th.data_ptr(), d_out.data_ptr(), 8, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert e.launch_count == launches + 1
    out = d_out.cpu().numpy()
    assert np.array_equal(out[0], s) and np.array_equal(out[1], q)


def test_error_codes(fb):
    data = np.zeros((100, 2))
    with fb.FrontierEngine.normal(data) as e:
        with pytest.raises(fb.PfetchError, match="PF_ERR_INVALID"):
            e.eval_frontier(np.zeros((0, 5)))
        with pytest.raises(fb.PfetchError, match="PF_ERR_SHAPE"):
            e.eval_frontier_range(np.zeros((1, 5)), 50, 51)
    with pytest.raises(fb.PfetchError, match="PF_ERR_UNSUPPORTED"):
        fb.FrontierEngine.normal(np.zeros((100, 5)))
    with pytest.raises(fb.PfetchError, match="PF_ERR_SHAPE"):
        fb.FrontierEngine.blasso(np.zeros((10, 3)), np.zeros(10), item_rows=18000)

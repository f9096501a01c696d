"""Accuracy of the engine's own exp/log device kernels (fetching_b200/csrc/pf_math.cuh)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(which, x):
    from fetching_b200 import _lib
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    _lib.check(lib.pf_debug_math(which, x.ctypes.data_as(dp), out.ctypes.data_as(dp), x.size, 0), "pf_debug_math")
    return out


def test_exp_neg_half_fp64():
    rng = np.random.RandomState(0)
    e = np.concatenate([rng.uniform(0, 60, 200000), rng.uniform(0, 1500, 100000), 10.0 ** rng.uniform(-12, 1, 50000),
                        [0.0, 1e-300, 1415.0, 1416.0, 1417.0, 1480.0, 1490.0, 1500.0, 3000.0, np.inf]])
    got = run(0, e)
    want = np.exp(-0.5 * e)
    normal = want > 1e-300
    rel = np.abs(got[normal] - want[normal]) / want[normal]
    # |z| * 2^-53 argument rounding on top of ~1 ulp: < 5e-14 even at the underflow edge, < 4e-16 for e < 60
    assert rel.max() < 5e-14
    assert rel[e[normal] < 60].max() < 8e-15  # |z| <= 43 -> 43 * 2^-52 * ln2
    assert np.all(np.abs(got[~normal] - want[~normal]) <= 1e-300)
    assert got[-1] == 0.0 and got[-2] == 0.0


def test_log_pos_fp64():
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.uniform(1e-3, 8.0, 200000), 10.0 ** rng.uniform(-300, 300, 100000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 50000), [1.0, 2.0, 0.5, np.sqrt(2), 1 / np.sqrt(2), 1e-310, 0.0]])
    got = run(1, x)
    with np.errstate(divide="ignore"):
        want = np.log(x)
    fin = np.isfinite(want) & (want != 0)
    err = np.abs(got[fin] - want[fin])
    assert (err / np.maximum(np.abs(want[fin]), 1e-300)).max() < 1e-15 or (err / np.spacing(np.abs(want[fin]))).max() <= 4
    assert got[-1] == -np.inf and got[-7] == 0.0


def test_log_table_fp64():
    """The kernel's table-driven log: absolute error <= 5e-16 on the range a mixture row sum lives in
    (what a SUM of logs needs), relative error <= 1e-13 next to 1, exact at 1; full range via the slow path."""
    rng = np.random.RandomState(3)
    x = np.concatenate([rng.uniform(1e-3, 8.0, 300000), 1.0 + rng.uniform(-4e-3, 4e-3, 100000),
                        1.0 + rng.uniform(-1e-9, 1e-9, 1000), 10.0 ** rng.uniform(-300, 300, 100000),
                        [1.0, 2.0, 0.5, 1e-310, 0.0]])
    got = run(4, x)
    with np.errstate(divide="ignore"):
        want = np.log(x)
    mid = (x > 1e-3) & (x < 8.0)
    assert np.abs(got[mid] - want[mid]).max() < 1e-15
    nz = np.isfinite(want) & (want != 0)
    assert (np.abs(got[nz] - want[nz]) / np.abs(want[nz])).max() < 1e-13
    assert got[-5] == 0.0 and got[-1] == -np.inf


def test_fp32_variants():
    rng = np.random.RandomState(2)
    e = rng.uniform(0, 80, 100000)
    got = run(2, e)
    want = np.exp(-0.5 * e)
    assert (np.abs(got - want) / want).max() < 2e-5
    x = rng.uniform(1e-3, 8.0, 100000)
    assert np.abs(run(3, x) - np.log(x)).max() < 2e-6


def test_mixture_loop_exp_fp64():
    """mix_exp_accumulate: 2048-entry table, monic degree-3 polynomial, clamp at e = 1408 (a clamped term is
    ~2^-1021 instead of the true, even smaller value: both are far below the slow-path threshold 2^-959)."""
    rng = np.random.RandomState(5)
    e = np.concatenate([rng.uniform(0, 60, 300000), rng.uniform(0, 1400, 100000), 10.0 ** rng.uniform(-12, 1, 50000),
                        [0.0, 1e-300, 1407.9, 1408.0, 1500.0, 1e6, 1e12, 1e300, np.inf]])
    got = run(5, e)
    want = np.exp(-0.5 * e)
    ok = e < 1400
    rel = np.abs(got[ok] - want[ok]) / want[ok]
    assert rel.max() < 5e-14                     # |z| 2^-53 argument rounding dominates near the edge
    assert rel[e[ok] < 60].max() < 8e-15
    assert np.all(got[~ok] < 2.0 ** -1000) and np.all(got[~ok] >= 0.0)  # clamped: tiny, finite, never wrapped
    assert np.all(np.isnan(run(5, np.array([np.nan]))) | (run(5, np.array([np.nan])) < 2.0 ** -1000))


def test_mixture_loop_log_fp64():
    rng = np.random.RandomState(6)
    x = np.concatenate([rng.uniform(1e-3, 8.0, 300000), 10.0 ** rng.uniform(-280, 300, 100000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 50000), [1.0, 2.0, 0.5, np.sqrt(2), 1 / np.sqrt(2),
                                                                 1.4140625, 1.41406249, 0.70703125]])
    got = run(6, x)
    want = np.log(x.astype(np.longdouble)).astype(np.float64)
    err = np.abs(got - want)
    assert np.max(err / np.maximum(np.abs(want), 1.0)) < 6e-16  # absolute next to 1, relative elsewhere
    assert got[300000 + 100000 + 50000] == 0.0                  # log(1) exact


def test_mixture_loop_exp2_fp64():
    """mix_exp_accumulate_z (the factored loop): 2^z by the 256-entry table (16 conflict-free copies in shared memory)
    and the degree-3 minimax polynomial on |g| <= 2^-9: truncation <= 1.8e-14 relative."""
    rng = np.random.RandomState(7)
    z = np.concatenate([rng.uniform(-60, 60, 400000), rng.uniform(-512, 512, 100000), [0.0, 1.0, -1.0, 511.9, -511.9]])
    got = run(7, z)
    want = np.exp2(z.astype(np.longdouble)).astype(np.float64)
    rel = (got - want) / want
    assert np.abs(rel).max() < 2e-14
    assert abs(rel[:400000].mean()) < 2e-15          # no bias to speak of: what a SUM over rows sees
    assert got[400000 + 100000] == pytest.approx(1.0, rel=2e-14)


def test_mixture_loop_exp_clamped_fp64():
    """mix_exp_accumulate_e (the clamped squared-distance loop): same polynomial and table, clamp at e = 1408."""
    rng = np.random.RandomState(8)
    e = np.concatenate([rng.uniform(0, 60, 300000), rng.uniform(0, 1400, 100000), 10.0 ** rng.uniform(-12, 1, 50000),
                        [0.0, 1e-300, 1407.9, 1408.0, 1500.0, 1e6, 1e12, 1e300, np.inf]])
    got = run(9, e)
    want = np.exp(-0.5 * e.astype(np.longdouble)).astype(np.float64)
    ok = e < 1400
    rel = np.abs(got[ok] - want[ok]) / want[ok]
    assert rel.max() < 7e-14                     # 1.8e-14 truncation + |z| 2^-53 argument rounding near the edge
    assert rel[e[ok] < 60].max() < 2.5e-14
    assert np.all(got[~ok] < 2.0 ** -1000) and np.all(got[~ok] >= 0.0)  # clamped: tiny, finite, never wrapped
    assert np.all(np.isnan(run(9, np.array([np.nan]))) | (run(9, np.array([np.nan])) < 2.0 ** -1000))


def test_mixture_loop_log_degree4_fp64():
    """log_pos_acc + log_n_times_ln2: degree-4 series on |r| <= 2^-9 (r^5/5 <= 5.7e-15 absolute) with n ln 2 added
    apart -- the form the kernels use, where n is summed as an integer over a lane's rows of a tile."""
    rng = np.random.RandomState(9)
    x = np.concatenate([rng.uniform(1e-3, 8.0, 300000), 10.0 ** rng.uniform(-280, 300, 100000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 50000), [1.0, 2.0, 0.5, np.sqrt(2), 1 / np.sqrt(2),
                                                                 1.4140625, 1.41406249, 0.70703125]])
    got = run(8, x)
    want = np.log(x.astype(np.longdouble)).astype(np.float64)
    err = np.abs(got - want)
    assert np.max(err / np.maximum(np.abs(want), 1.0)) < 7e-15
    assert abs(got[300000 + 100000 + 50000]) < 7e-15            # log(1): not exact with midpoint tables, and need not be

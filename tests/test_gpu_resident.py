"""The resident frontier kernel (fetching_b200/csrc/pf_resident.cu, PF_OPT_RESIDENT): host-buffer calls over the whole
data set of a pointwise model are served by a kernel that stays on the SMs between calls.  Parity with the oracle,
bit-identity with the launched kernels at equal geometry, and the life cycle (idle exit, stop on other work, two
engines on one device, interleaved range / device-pointer calls)."""
import time

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
    _lib.load()
    return fetching_b200


def _on_off(fb, make, thetas):
    """(resident result, launched result, resident stats) of the same call on two engines over the same data"""
    from fetching_b200._lib import OPT_RESIDENT
    with make() as e:
        r = e.eval_frontier(thetas)
        stats = e.resident_stats
    with make() as e:
        e.set_option(OPT_RESIDENT, 0)
        l = e.eval_frontier(thetas)
        assert e.resident_stats[1] == 0
    return r, l, stats


@pytest.mark.parametrize("B", [1, 3, 8, 27, 32, 33, 64, 70])
def test_normal_resident_vs_oracle_and_launched(fb, orc, B):
    from fetching_b200 import models
    data = orc.normal_dataset(50021)
    th = models.normal_thetas(B, seed=B, center=(200.0, 100.0), scale=2.0)
    want = orc.normal_frontier(data, th)
    (s, q), (s0, q0), stats = _on_off(fb, lambda: fb.FrontierEngine.normal(data), th)
    assert stats[0] and stats[1] >= 1 and stats[2] == 1, stats
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
    assert np.array_equal(s, s0) and np.array_equal(q, q0)  # same geometry, same order of additions: same bits


def test_normal_general_covariance_switches_per_call(fb, orc, ref_normal):
    """identity and general covariances alternate call by call inside one launch of the resident kernel"""
    data = orc.normal_dataset(int(ref_normal["n"]))
    th, S, Q = ref_normal["thetas"], ref_normal["sum"], ref_normal["sum_sq"]
    with fb.FrontierEngine.normal(data) as e:
        for lo, hi in ((0, 12), (12, 20), (0, 20), (3, 9), (12, 13)):
            s, q = e.eval_frontier(th[lo:hi])
            assert rel(s, S[lo:hi]) < TOL64 and rel(q, Q[lo:hi]) < TOL64
        ok, calls, launches = e.resident_stats
        assert ok and calls == 5 and launches == 1


def test_normal_1e6_bit_identical_to_launched_kernel(fb, orc):
    """BASELINE config 1: same CTA geometry (296 CTAs x 3379 rows) => the very same bits"""
    from fetching_b200 import models
    data = models.normal_dataset(1_000_000)
    th = models.normal_thetas(8, seed=3)
    (s, q), (s0, q0), stats = _on_off(fb, lambda: fb.FrontierEngine.normal(data), th)
    assert stats[0] and stats[1] == 1
    assert np.array_equal(s, s0) and np.array_equal(q, q0)
    want = orc.normal_frontier(data, th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


@pytest.mark.parametrize("B", [1, 8, 24, 32, 40])
def test_mixture_1e6_resident(fb, orc, B):
    """BASELINE config 3: K = 8, d = 2, N = 1e6, items of 1000 rows cut by the CTA boundaries"""
    from fetching_b200 import models
    data, pos = models.mixture_dataset(1_000_000, 8, 2, seed=0)
    th = models.mixture_thetas(pos, B, seed=B)
    (s, q), (s0, q0), stats = _on_off(fb, lambda: fb.FrontierEngine.mixture(data, 8, 1000), th)
    assert stats[0] and stats[1] >= 1 and stats[2] == 1, stats
    assert np.array_equal(s, s0) and np.array_equal(q, q0)
    want = orc.mixture_frontier(data, th, 1000)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64


@pytest.mark.parametrize("shape", [(8, 2, 1000), (8, 2, 7), (8, 8, 500), (5, 3, 100), (3, 1, 64), (4, 2, 250), (16, 2, 1000), (8, 4, 500), (3, 2, 100), (6, 2, 333)])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_mixture_shapes_resident(fb, orc, shape, dtype):
    from fetching_b200 import models
    from fetching_b200._lib import F32, F64
    K, d, batch = shape
    data, pos = models.mixture_dataset(60000, K, d, seed=K + d)
    th = models.mixture_thetas(pos, 6, seed=batch)
    kw = {"dtype": F32 if dtype == "f32" else F64}
    want = orc.mixture_frontier(data.astype(np.float32).astype(np.float64) if dtype == "f32" else data, th, batch)
    (s, q), (s0, q0), stats = _on_off(fb, lambda: fb.FrontierEngine.mixture(data, K, batch, **kw), th)
    tol = TOL32 if dtype == "f32" else TOL64
    assert stats[0] and stats[1] == 1
    assert rel(s, want[0]) < tol and rel(q, want[1]) < tol
    assert np.array_equal(s, s0) and np.array_equal(q, q0)


def test_mixture_far_and_nan_thetas_take_the_clamped_loop(fb, orc):
    """the factored / clamped switch is per call inside the resident kernel; NaN poisons only its node"""
    from fetching_b200 import models
    data, pos = models.mixture_dataset(40000, 8, 2, seed=1)
    near = models.mixture_thetas(pos, 4, seed=2)
    far = near.copy()
    far[1, :2] = 900.0  # fails the range test: the whole call uses the clamped distance form
    nan = near.copy()
    nan[2, 5] = np.nan
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        for th in (near, far, near, nan, near):
            s, q = e.eval_frontier(th)
            want = orc.mixture_frontier(data, th, 1000)
            fin = np.isfinite(want[0])
            assert rel(s[fin], want[0][fin]) < TOL64 and rel(q[fin], want[1][fin]) < TOL64
            assert np.array_equal(np.isnan(s), np.isnan(want[0]))
        assert e.resident_stats[1:] == (5, 1)


def test_many_calls_one_launch_and_bit_stable(fb):
    from fetching_b200 import models
    data = models.normal_dataset(200000)
    th = models.normal_thetas(8, seed=5)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(th)
        for _ in range(300):
            s2, q2 = e.eval_frontier(th)
            assert np.array_equal(s, s2) and np.array_equal(q, q2)
        s3, q3, sec = e.eval_frontier_repeat(th, 2000)
        assert np.array_equal(s, s3) and np.array_equal(q, q3)
        ok, calls, launches = e.resident_stats
        assert ok and calls == 2301 and launches == 1
        assert e.launch_count == 1  # 2301 calls, one kernel launch


def test_idle_exit_and_relaunch(fb, orc):
    """after 2 ms without a call the kernel leaves the SMs by itself; the next call brings it back"""
    from fetching_b200 import models
    data = orc.normal_dataset(30000)
    th = models.normal_thetas(4, seed=1, center=(200.0, 100.0), scale=2.0)
    want = orc.normal_frontier(data, th)
    with fb.FrontierEngine.normal(data) as e:
        for i in range(4):
            s, q = e.eval_frontier(th)
            assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
            time.sleep(0.05)
        ok, calls, launches = e.resident_stats
        assert calls == 4 and launches == 4


def test_interleaved_with_range_device_and_other_engines(fb, orc):
    """anything that launches a kernel asks the resident one to leave first; results never change"""
    import torch
    from fetching_b200 import models
    data = orc.normal_dataset(40000)
    th = models.normal_thetas(6, seed=2, center=(200.0, 100.0), scale=2.0)
    want = orc.normal_frontier(data, th)
    mdata, pos = models.mixture_dataset(50000, 8, 2, seed=3)
    mth = models.mixture_thetas(pos, 5, seed=4)
    mwant = orc.mixture_frontier(mdata, mth, 1000)
    d_th = torch.tensor(th, device="cuda")
    d_out = torch.zeros(2, 6, dtype=torch.float64, device="cuda")
    with fb.FrontierEngine.normal(data) as a, fb.FrontierEngine.mixture(mdata, 8, 1000) as b:
        for _ in range(3):
            s, q = a.eval_frontier(th)  # resident (a)
            assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64
            ms, mq = b.eval_frontier(mth)  # resident (b): a's kernel is stopped first
            assert rel(ms, mwant[0]) < TOL64 and rel(mq, mwant[1]) < TOL64
            s1, q1 = a.eval_frontier_range(th, 0, 10000)  # launched
            s2, q2 = a.eval_frontier_range(th, 10000, 30000)
            assert rel(s1 + s2, want[0]) < TOL64 and rel(q1 + q2, want[1]) < TOL64
            a.eval_frontier_device(d_th.data_ptr(), d_out.data_ptr(), 6, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert rel(d_out[0].cpu().numpy(), want[0]) < TOL64
            x = torch.ones(1 << 20, device="cuda").sum().item()  # foreign work: waits at most for the idle timeout
            assert x == float(1 << 20)
        assert a.resident_stats[1] == 3 and b.resident_stats[1] == 3


def test_data_set_too_large_for_shared_memory_uses_launches(fb, orc):
    from fetching_b200 import models
    data = models.normal_dataset(3_000_000)
    th = models.normal_thetas(4, seed=1)
    with fb.FrontierEngine.normal(data) as e:
        s, q = e.eval_frontier(th)
        ok, calls, launches = e.resident_stats
        assert not ok and calls == 0 and launches == 0
    want = orc.normal_frontier(data, th)
    assert rel(s, want[0]) < TOL64 and rel(q, want[1]) < TOL64

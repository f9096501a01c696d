"""The oracle is pinned before it is trusted (CPU only, no product code involved).

Sources of truth, in order:
  * tests/golden/*.tsv / ref_*.npz : outputs of the reference's own C++ (oracle/_ref/pf_ref);
  * tests/golden/py_models.npz     : outputs of the reference's own Python model classes;
  * oracle/_ref/pf_ref itself when the binary is present (it travels to the GPU box).
"""
import hashlib
import os

import numpy as np
import pytest


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_mt19937_stream_matches_reference_unit_test(orc):
    # testmaster.cc:295-332 pins RandomSequence == raw mt19937(18340); numpy's legacy
    # RandomState is the same MT19937 with the same init_genrand seeding.
    ours = orc.mt_stream(18340, 11000)
    rs = np.random.RandomState(18340)
    theirs = rs.randint(0, 2 ** 32, size=11000, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(ours, theirs)


def test_dataset_recipe_bit_exact(orc, ref_normal):
    data = orc.normal_dataset(int(ref_normal["n"]))
    assert np.array_equal(data[:4].ravel(), ref_normal["data_head"])
    assert hashlib.sha256(data.tobytes()).hexdigest() == str(ref_normal["data_sha256"])
    # SURVEY section 8c.3 probe goldens
    assert data[0, 0] == 200.21343601372288 and data[0, 1] == 99.50441971788122


def test_normal_eval_bit_exact_vs_reference_cpp(orc, ref_normal):
    """a1/a2: same visiting order (StrideIndex from the replay stream at 0), same arithmetic."""
    n = int(ref_normal["n"])
    data = orc.normal_dataset(n)
    stride, start, pos = orc.normal_prepare_draws(n, 0)
    assert pos == 2
    for th, s_ref, q_ref in zip(ref_normal["thetas"], ref_normal["sum"], ref_normal["sum_sq"]):
        s, q = orc.normal_eval(data, th, stride, start)
        assert s == s_ref and q == q_ref, th
        # natural order and extended precision agree to rounding
        s2, q2 = orc.normal_eval(data, th)
        sl, ql = orc.normal_eval_ld(data, th)
        assert rel(s2, s_ref) < 1e-13 and rel(q2, q_ref) < 1e-13
        assert rel(float(sl), s_ref) < 1e-13 and rel(float(ql), q_ref) < 1e-13


def test_normal_eval_accumulates_like_evaluate_many(orc):
    data = orc.normal_dataset(1000)
    th = np.array([1.0, 2.0, 1, 0, 1.0])
    s, q = orc.normal_eval(data, th, 7, 3, 0, 400)
    s, q = orc.normal_eval(data, th, 7, 3, 400, 600, s, q)
    assert (s, q) == orc.normal_eval(data, th, 7, 3)


def test_scalar_boltzmann_vs_reference_cpp(orc):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_boltz_eval.npz"))
    for th, s_ref, q_ref in zip(g["thetas"], g["sum"], g["sum_sq"]):
        s, q = orc.boltzmann_scalar_eval(float(th), int(g["count"]))
        assert s == s_ref and q == q_ref


def test_stride_index(orc):
    # strideindex.hh:50-57: result is coprime to n (or 1)
    for n, s in [(10000, 5000), (10000, 0), (10000, 1), (10000, 9999), (12, 8), (1000000, 123456)]:
        v = orc.valid_stride(n, s)
        assert 1 <= v < n and np.gcd(n, v) == 1
    assert orc.valid_stride(10, 5) == 9  # 5 -> (5+5-1)%10 = 9
    n, stride, start = 12, 5, 7
    seen = sorted(int(orc.lib().orc_stride_index(n, stride, start, i)) for i in range(n))
    assert seen == list(range(n))


def test_mixture_vs_reference_python(orc, py_models):
    """a5: C restatement and numpy statement vs the reference Mixture.evaluate outputs."""
    g = py_models
    data, b = g["mix_data"], int(g["mix_item_rows"])
    assert orc.mixture_n_items(data.shape[0], b) == int(g["mix_n_items"])
    for t, theta in enumerate(g["mix_thetas"]):
        for pos in range(int(g["mix_n_items"])):
            want = g["mix_item_lp"][t, pos]
            assert rel(orc.mixture_item(data, theta, b, pos), want) < 1e-12
            assert rel(orc.np_mixture_evaluate(data, theta, b, pos), want) < 1e-13
        s, q = orc.mixture_eval(data, theta, b)
        assert rel(s, g["mix_item_lp"][t].sum()) < 1e-12
        assert rel(q, (g["mix_item_lp"][t] ** 2).sum()) < 1e-12
        assert rel(s, g["mix_log_posterior"][t]) < 1e-12  # pymixture.py:101-106 cross-check
        sl, ql = orc.mixture_eval_ld(data, theta, b)
        assert rel(float(sl), s) < 1e-12 and rel(float(ql), q) < 1e-12


def test_mixture_short_item_is_zero_like_python_cc(orc, py_models):
    data = py_models["mix_data"][:3999]
    theta = py_models["mix_thetas"][0]
    assert orc.mixture_n_items(3999, 250) == 16
    assert orc.mixture_item(data, theta, 250, 15) == 0.0
    assert orc.mixture_item(data, theta, 250, 14) != 0.0


def test_mixture_underflow_is_minus_inf(orc):
    data = np.zeros((10, 2))
    theta = np.full((8, 2), 60.0)  # exp(-3600) underflows -> log(0) = -inf, as numpy gives
    assert orc.mixture_item(data, theta, 10, 0) == -np.inf


def test_blasso_vs_reference_python(orc, py_models):
    """a7/a8 vs the reference BLasso.evaluate / log_prior outputs."""
    g = py_models
    X, y, b = g["bl_X"], g["bl_y"], int(g["bl_item_rows"])
    assert orc.blasso_n_items(X.shape[0], b) == int(g["bl_n_items"])
    for t, theta in enumerate(g["bl_thetas"]):
        for pos in range(int(g["bl_n_items"])):
            want = g["bl_item_lp"][t, pos]
            assert rel(orc.blasso_item(X, y, theta, b, pos), want) < 1e-12
            assert rel(orc.np_blasso_evaluate(X, y, theta, pos, b), want) < 1e-13
        s, q = orc.blasso_eval(X, y, theta, b)
        assert rel(s, g["bl_item_lp"][t].sum()) < 1e-12
        assert rel(s, g["bl_contrib_sum"][t]) < 1e-12  # pyblasso.py:152-155 cross-check
        assert rel(q, (g["bl_item_lp"][t] ** 2).sum()) < 1e-12
        sl, _ = orc.blasso_eval_ld(X, y, theta, b)
        assert rel(float(sl), s) < 1e-12
        assert rel(orc.blasso_log_prior(theta, float(g["bl_tau"])), g["bl_log_prior"][t]) < 1e-13
        assert rel(orc.np_blasso_log_prior(theta, X.shape[1], float(g["bl_tau"])), g["bl_log_prior"][t]) < 1e-15


def test_dense_boltzmann_definition(orc):
    """New workload (no reference code): sum_i e_i == -x'Wx/T; long-double companion agrees."""
    rng = np.random.RandomState(1)
    D = 64
    W = rng.normal(0, 1 / np.sqrt(D), (D, D))
    W = (W + W.T) / 2
    x = rng.normal(0, 1, D)
    s, q = orc.boltzmann_dense_eval(W, x, 2.0)
    assert rel(s, -(x @ W @ x) / 2.0) < 1e-12
    sl, ql = orc.boltzmann_dense_eval_ld(W, x, 2.0)
    assert rel(float(sl), s) < 1e-12 and rel(float(ql), q) < 1e-12
    assert orc.boltzmann_dense_logp(W, x, 2.0) == s
    S, Q = orc.boltzmann_dense_frontier(W, np.stack([x, 2 * x]), 2.0)
    assert S[0] == s and Q[0] == s * s and rel(S[1], 4 * s) < 1e-14
    s1, q1 = orc.boltzmann_dense_eval(W, x, 2.0, 0, 10)
    s2, q2 = orc.boltzmann_dense_eval(W, x, 2.0, 10, D - 10)
    assert rel(s1 + s2, s) < 1e-12 and rel(q1 + q2, q) < 1e-12


def test_frontier_helpers_match_single_node(orc, py_models):
    g = py_models
    s, q = orc.mixture_frontier(g["mix_data"], g["mix_thetas"].reshape(len(g["mix_thetas"]), -1), 250)
    for t, theta in enumerate(g["mix_thetas"]):
        assert (s[t], q[t]) == orc.mixture_eval(g["mix_data"], theta, 250)
    s, q = orc.blasso_frontier(g["bl_X"], g["bl_y"], g["bl_thetas"], 18000)
    for t, theta in enumerate(g["bl_thetas"]):
        assert (s[t], q[t]) == orc.blasso_eval(g["bl_X"], g["bl_y"], theta, 18000)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "pf_ref")),
                    reason="oracle/_ref/pf_ref not built (needs /root/reference)")
def test_reference_binary_reproduces_goldens(orc, golden_dir, ref_normal):
    assert "OK" in orc.run_pf_ref("unit")  # the reference's own three assert tests
    gold = open(os.path.join(golden_dir, "chain_normal_d10000.tsv")).read()
    rows = orc.parse_chain(gold)
    again = orc.parse_chain(orc.run_pf_ref("chain", "normal", "-n", "5", "-d", "10000", "-b", "2000", "-l", "40"))
    assert again == rows[:41]
    # SURVEY section 8c.3 probe values
    assert rows[0][2] == -250005158.44676721 and rows[1][1] == 1 and rows[4][1] == 0
    data = orc.normal_dataset(int(ref_normal["n"]))
    res = orc.pf_ref_normal_eval(data, ref_normal["thetas"][:3])
    assert np.array_equal(res[:, 0], ref_normal["sum"][:3])


def test_reference_parallel_program_with_threads_for_ranks(orc):
    """`pf_ref mpirun -np N normal`: the reference's run_master on rank 0 and worker<NormalSampler<2>>() on ranks
    1..N-1 (samplers/normal.cc, N = 10^6 points each), unmodified, over the thread-mailbox boost::mpi shim.  The chain
    must not depend on the number of ranks or on thread timing (that is the point of prefetching MH), and its first
    proposals are the ones every golden chain starts with (same replayable stream)."""
    if not orc.have_pf_ref():
        pytest.skip("oracle/_ref/pf_ref not built")

    def chain(np_ranks):
        out = orc.run_pf_ref("mpirun", "-np", np_ranks, "normal", "-l", 25)
        rows = [l.split("\t") for l in out.splitlines() if l[:1].isdigit()]
        return [(int(r[1]), int(r[2]), r[4]) for r in rows], [float(r[3]) for r in rows]

    a, sa = chain(2)
    b, sb = chain(5)
    assert len(a) >= 26 and a[:26] == b[:26]
    assert all(abs(x - y) <= 1e-12 * abs(x) for x, y in zip(sa[:26], sb[:26]))
    assert a[0] == (0, 1, "[0, 0; 1, 0, 1]") and a[1][2] == "[0.157538, -0.10592; 1, 0, 1]"


def test_dense_boltzmann_two_unrelated_statements_agree(orc):
    """No reference code exists for the dense energy, so the C restatement is checked against an independent one:
    symmetric half, pairwise summation, 80-bit (oracle.boltzmann_dense_logp_independent)."""
    from fetching_b200 import models
    for D, seed in ((64, 1), (200, 2), (1024, 3)):
        W, X = models.boltzmann_dataset(D, 3, seed=seed)
        W = W + 0.3 * np.random.RandomState(seed).standard_normal((D, D)) / np.sqrt(D)   # not symmetric on purpose
        for x in X:
            a = orc.boltzmann_dense_logp(W, x, 1.7)
            b = orc.boltzmann_dense_logp_independent(W, x, 1.7)
            scale = np.abs(np.outer(x, x) * W).sum() / 1.7
            assert abs(a - float(b)) <= 2e-16 * D * scale / np.sqrt(D) + 1e-13 * abs(float(b))

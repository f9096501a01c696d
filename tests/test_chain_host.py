"""Host logic on the CPU: the product's speculative tree (SpecTree), executors and frontier
communicator against traces of the REFERENCE's own JobTree / run_master / samplers.

The golden traces (tests/golden/frontier_trace_*.tsv) were produced by oracle/_ref/pf_ref, i.e.
the reference's C++ compiled in place, driven by the same FrontierCommunicator header.  Here the
product's tree + host executors are run with the ORACLE as frontier evaluator (through
pf_chain_run_with_evaluator; no GPU involved) and must reproduce, bit for bit: every node id,
generation, job state, replay path and random-stream position (tree indexing / node ordering),
every proposal (theta bytes) and every accept/reject decision.  metric_sum: 1e-10 relative.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TRACES = {
    "frontier_trace_normal_W8.tsv": dict(model="normal", frontier=8, levels=80),
    "frontier_trace_normal_W16_deferred_naive.tsv": dict(model="normal", frontier=16, levels=60, policy=0, deferred=True),
    "frontier_trace_normal_W1.tsv": dict(model="normal", frontier=1, levels=30),
    "frontier_trace_normal_W64_const.tsv": dict(model="normal", frontier=64, levels=60, policy=1),
    "frontier_trace_boltzmann_W8.tsv": dict(model="boltzmann", frontier=8, levels=80),
    "frontier_trace_boltzmann_W5_deferred.tsv": dict(model="boltzmann", frontier=5, levels=60, deferred=True),
    "frontier_trace_normal_W8_b1000_fastslow.tsv": dict(model="normal", frontier=8, levels=80, batch=1000, style=0),
    "frontier_trace_normal_W16_b2000_incremental.tsv": dict(model="normal", frontier=16, levels=40, batch=2000, style=1),
    "frontier_trace_boltzmann_W8_b1000_fastslow.tsv": dict(model="boltzmann", frontier=8, levels=60, batch=1000, style=0),
}


def parse_trace(path):
    jobs, rows = [], []
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if f[0] == "J":
            jobs.append((int(f[1]), int(f[2]), int(f[3]), "" if f[4] == "-" else f[4], int(f[5]), int(f[6]), int(f[7]),
                         float(f[8]), bytes.fromhex(f[9])))
        elif f[0] == "R":
            rows.append((int(f[1]), int(f[2]), float(f[3]), bytes.fromhex(f[4])))
    return jobs, rows


def check_against_trace(res, path):
    jobs, rows = parse_trace(path)
    got_jobs = [(j[0], j[1], j[2], j[3], j[4], j[5], j[6], j[7], j[8].tobytes()) for j in res.jobs]
    assert len(got_jobs) == len(jobs)  # the very same nodes are handed out, in the same order
    for a, b in zip(got_jobs, jobs):
        assert a == b, (a[:8], b[:8])
    assert len(res.depth) == len(rows)
    for i, (depth, accept, metric_sum, theta) in enumerate(rows):
        assert (res.depth[i], res.accept[i], res.theta[i].tobytes()) == (depth, accept, theta), i
        assert abs(res.metric_sum[i] - metric_sum) <= 1e-10 * abs(metric_sum), i


def oracle_run(orc, cfg, **extra):
    from fetching_b200 import _lib, chain
    if cfg["model"] == "normal":
        data = orc.normal_dataset(10000)
        ev = lambda th: orc.normal_frontier(data, th)

        def ev_range(th, first, count, order):  # each node in its own StrideIndex order, like a reference worker
            out = np.array([orc.normal_eval(data, t, int(o[0]), int(o[1]), first, count) for t, o in zip(th, order)])
            return out[:, 0], out[:, 1]
        model, dim = _lib.MODEL_NORMAL, 2
    else:
        def ev(th):
            out = np.array([orc.boltzmann_scalar_eval(float(t[0]), 10000) for t in th])
            return out[:, 0], out[:, 1]

        def ev_range(th, first, count, order):
            out = np.array([orc.boltzmann_scalar_eval(float(t[0]), count) for t in th])
            return out[:, 0], out[:, 1]
        model, dim = _lib.MODEL_BOLTZMANN_SCALAR, 1
    return chain.run_chain_with_evaluator(model, dim, 0, 10000, ev, chain_length=cfg["levels"], frontier=cfg["frontier"],
                                          policy=cfg.get("policy", 2), deferred_proposals=cfg.get("deferred", False),
                                          batch_items=cfg.get("batch", 0), update_style=cfg.get("style", 1),
                                          range_evaluator=ev_range, trace_jobs=True, **extra)


@pytest.mark.parametrize("name", sorted(TRACES))
def test_tree_indexing_and_chain_match_reference_trace(orc, name):
    from fetching_b200 import build
    build.build()
    res = oracle_run(orc, TRACES[name])
    check_against_trace(res, os.path.join(GOLD, name))
    assert res.stats["max_frontier"] <= TRACES[name]["frontier"] and res.stats["engine_calls"] > 0


@pytest.mark.parametrize("frontier", [1, 2, 3, 7, 33])
def test_chain_does_not_depend_on_frontier_size(orc, frontier):
    """The invariant the reference is built on: same chain whatever was speculated."""
    res = oracle_run(orc, dict(model="normal", frontier=frontier, levels=120))
    gold = orc.parse_chain(open(os.path.join(GOLD, "chain_normal_d10000.tsv")).read())
    for i in range(121):
        assert (res.depth[i], res.accept[i], res.theta[i].tobytes()) == (gold[i][0], gold[i][1], gold[i][3])
        assert abs(res.metric_sum[i] - gold[i][2]) <= 1e-10 * abs(gold[i][2])
    # SURVEY 8c.3 probe goldens
    assert res.accept[:9].tolist() == [1, 1, 1, 1, 0, 1, 1, 0, 1]


def test_boltzmann_chain_matches_mastertester_golden(orc):
    res = oracle_run(orc, dict(model="boltzmann", frontier=9, levels=300))
    gold = orc.parse_chain(open(os.path.join(GOLD, "chain_boltzmann_d10000.tsv")).read())
    assert len(res.depth) == len(gold) == 301
    for i, g in enumerate(gold):
        assert (res.depth[i], res.accept[i], res.theta[i].tobytes()) == (g[0], g[1], g[3])
        assert abs(res.metric_sum[i] - g[2]) <= 1e-12 * abs(g[2])


def test_native_mixture_and_lasso_chains_run_and_are_frontier_invariant(orc):
    """The native samplers behind the same tree: the chain must not depend on the frontier size."""
    from fetching_b200 import _lib, chain, models
    data, pos = models.mixture_dataset(4000, 8, 2, seed=3)
    ev = lambda th: orc.mixture_frontier(data, th, 500)
    runs = [chain.run_chain_with_evaluator(_lib.MODEL_MIXTURE, 2, 8, 8, ev, chain_length=40, frontier=f, dimension=16,
                                           seed=0) for f in (1, 6)]
    assert np.array_equal(runs[0].theta, runs[1].theta) and np.array_equal(runs[0].accept, runs[1].accept)
    # first_proposal is numpy's: np.random.seed(0); 4 * np.random.random((8, 2)) - 2   (pymixture.py:69-73)
    np.random.seed(0)
    assert np.array_equal(runs[0].theta[0], (4.0 * np.random.random((8, 2)) - 2.0).ravel())
    assert 0 < runs[0].accept[1:].sum() < 40

    X, y, beta = models.blasso_dataset(3000, 6, seed=1)
    ev = lambda th: orc.blasso_frontier(X, y, th, 500)
    init = np.concatenate([[1.0], beta])
    runs = [chain.run_chain_with_evaluator(_lib.MODEL_BLASSO, 6, 0, 6, ev, chain_length=30, frontier=f, dimension=7,
                                           initial_theta=init, trace_jobs=True) for f in (1, 5)]
    assert np.array_equal(runs[0].theta, runs[1].theta) and np.array_equal(runs[0].accept, runs[1].accept)
    assert np.all(runs[0].theta[:, 0] > 0)  # sigma stays positive (pyblasso.py:117-121)
    for j in runs[1].jobs:  # log_prior travels with the proposal (python.cc:168-176 -> pyblasso.py:123-133)
        assert abs(j[7] - orc.blasso_log_prior(j[8])) <= 1e-12 * abs(j[7])


def test_predictive_partial_evaluation_saves_work_and_keeps_the_chain(orc):
    """Batched mode (the reference's -b / -y): partial sums drive the predictive scheduler and useless
    nodes are abandoned, so far fewer (node x item) evaluations buy the same chain."""
    whole = oracle_run(orc, dict(model="normal", frontier=16, levels=120))
    part = oracle_run(orc, dict(model="normal", frontier=16, levels=120, batch=1000, style=0))
    assert np.array_equal(whole.theta[:121], part.theta[:121]) and np.array_equal(whole.accept[:121], part.accept[:121])
    assert np.max(np.abs(whole.metric_sum[:121] - part.metric_sum[:121]) / np.abs(whole.metric_sum[:121])) < 1e-12
    assert whole.stats["items_scored"] == whole.stats["nodes_scored"] * 10000
    assert part.stats["items_scored"] < 0.5 * whole.stats["items_scored"]
    assert part.stats["abandoned"] > 0


def test_chain_stats_count_acceptances_without_a_row_callback(orc):
    """Timing runs pass no row callback (collect_rows=False semantics of run_chain): the acceptance count must
    then come from pf_chain_stats.accepted."""
    from fetching_b200 import chain as pf_chain
    data = orc.normal_dataset(2000)

    def ev(th):
        return orc.normal_frontier(data, th)

    r = pf_chain.run_chain_with_evaluator(1, 2, 0, len(data), ev, chain_length=60, frontier=8)
    assert int(r.stats["accepted"]) == int(r.accept[r.depth >= 1].sum())
    assert int(r.stats["levels"]) >= 60


def test_shared_noise_and_stream_do_not_change_proposals(orc):
    """Executors of the virtual workers share one replayable stream and, per stream position, one set of Box-Muller
    products (host/rng.hpp: ReplayStream, NoiseCache).  With a frontier of 1 every position is visited by a single
    node (the cache never hits), with a wide frontier most nodes are served from it: the chains must be identical,
    also through the lasso's redraw path (sigma <= 0 -> draw again from the SAME distribution object, odd length)."""
    from fetching_b200 import _lib, chain, models
    D = 301
    W, _ = models.boltzmann_dataset(D, 4, seed=2)
    ev = lambda th: orc.boltzmann_dense_frontier(W, th, 1.0)
    runs = [chain.run_chain_with_evaluator(_lib.MODEL_BOLTZMANN_DENSE, D, 0, 1, ev, chain_length=40, frontier=f,
                                           dimension=D, deferred_proposals=d) for f, d in ((1, False), (64, False), (8, True))]
    for r in runs[1:]:
        assert np.array_equal(runs[0].theta, r.theta) and np.array_equal(runs[0].accept, r.accept)
        assert np.array_equal(runs[0].metric_sum, r.metric_sum)
    # the scalar distribution gives the same values as the cached products: first proposal of the chain by hand
    step = runs[0].theta[1] - runs[0].theta[0] if runs[0].accept[1] else None
    assert runs[0].theta.shape[1] == D and (step is None or np.all(np.abs(step) < 1.0))

    p = 40  # theta has 41 values (odd); sigma starts tiny, so proposals with sigma <= 0 and redraws are frequent
    X, y, beta = models.blasso_dataset(2000, p, seed=4)
    ev = lambda th: orc.blasso_frontier(X, y, th, 500)
    init = np.concatenate([[1e-3], beta])
    runs = [chain.run_chain_with_evaluator(_lib.MODEL_BLASSO, p, 0, 4, ev, chain_length=40, frontier=f, dimension=p + 1,
                                           initial_theta=init) for f in (1, 16)]
    assert np.array_equal(runs[0].theta, runs[1].theta) and np.array_equal(runs[0].accept, runs[1].accept)
    assert np.all(runs[0].theta[:, 0] > 0)


def test_failing_evaluator_stops_the_chain(orc):
    """An engine failure must end the chain at once: no level may be decided on the zeroed sums of a
    failed call, no further evaluator call is made, and the error comes back as the status."""
    from fetching_b200 import _lib, chain
    data = orc.normal_dataset(10000)
    calls = []

    def ev(th):
        calls.append(len(th))
        if len(calls) == 4:
            return _lib.PF_ERR_CUDA
        return orc.normal_frontier(data, th)

    good = chain.run_chain_with_evaluator(_lib.MODEL_NORMAL, 2, 0, 10000, lambda th: orc.normal_frontier(data, th),
                                          chain_length=200, frontier=8)
    bad = chain.run_chain_with_evaluator(_lib.MODEL_NORMAL, 2, 0, 10000, ev, chain_length=200, frontier=8,
                                         check_status=False)
    assert bad.status == _lib.PF_ERR_CUDA
    assert len(calls) == 4                      # nothing was evaluated after the failure
    n = len(bad.depth)
    assert 0 < n < len(good.depth)              # the chain stopped early ...
    assert np.array_equal(bad.theta, good.theta[:n]) and np.array_equal(bad.accept, good.accept[:n])  # ... and every
    assert np.array_equal(bad.metric_sum, good.metric_sum[:n])  # row it did deliver is a row of the healthy chain
    with pytest.raises(_lib.PfetchError):
        chain.run_chain_with_evaluator(_lib.MODEL_NORMAL, 2, 0, 10000, lambda th: _lib.PF_ERR_CUDA, chain_length=50,
                                       frontier=4)


def test_failing_range_evaluator_stops_the_chain(orc):
    from fetching_b200 import _lib, chain
    data = orc.normal_dataset(10000)
    calls = []

    def ev_range(th, first, count, order):
        calls.append((first, count))
        if len(calls) == 7:
            return _lib.PF_ERR_NCCL
        out = np.array([orc.normal_eval(data, t, int(o[0]), int(o[1]), first, count) for t, o in zip(th, order)])
        return out[:, 0], out[:, 1]

    bad = chain.run_chain_with_evaluator(_lib.MODEL_NORMAL, 2, 0, 10000, lambda th: orc.normal_frontier(data, th),
                                         chain_length=100, frontier=8, batch_items=1000, range_evaluator=ev_range,
                                         check_status=False)
    assert bad.status == _lib.PF_ERR_NCCL and len(calls) == 7


@pytest.mark.parametrize("style", [0, 2])
def test_probe_batch_styles_leave_the_chain_alone(orc, style):
    """-y 0 (FAST_SLOW) and -y 2 (LOGARITHMIC, worker.hh:141-147: the worker's probe batch doubles with every job it
    starts from item 0): same chain as whole-node scoring; with style 2 the probe batches of a worker grow."""
    from fetching_b200 import _lib, chain
    data = orc.normal_dataset(10000)
    ranges = []

    def ev_range(th, first, count, order):
        ranges.append((first, count, len(th)))
        out = np.array([orc.normal_eval(data, t, int(o[0]), int(o[1]), first, count) for t, o in zip(th, order)])
        return out[:, 0], out[:, 1]

    ev = lambda th: orc.normal_frontier(data, th)
    whole = chain.run_chain_with_evaluator(_lib.MODEL_NORMAL, 2, 0, 10000, ev, chain_length=50, frontier=4)
    part = chain.run_chain_with_evaluator(_lib.MODEL_NORMAL, 2, 0, 10000, ev, chain_length=50, frontier=4,
                                          batch_items=100, update_style=style, range_evaluator=ev_range)
    n = min(len(whole.depth), len(part.depth))
    assert n >= 51
    assert np.array_equal(whole.theta[:n], part.theta[:n]) and np.array_equal(whole.accept[:n], part.accept[:n])
    assert np.max(np.abs(whole.metric_sum[:n] - part.metric_sum[:n]) / np.abs(whole.metric_sum[:n])) < 1e-10
    probes = sorted({c for f, c, _ in ranges if f == 0})
    if style == 0:
        assert probes == [100]
    else:
        assert probes[0] == 100 and len(probes) > 2 and all(p % 100 == 0 and (p // 100) & (p // 100 - 1) == 0 or p == 10000
                                                             for p in probes)


def test_cli_resolves_samplers_through_the_registry():
    """tree.cc:121-131: the sampler is looked up by name, an unknown name lists the registered ones, and the plugin
    parses its own key=value arguments (set_arguments < 0 aborts).  None of this needs a GPU."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fetching_b200", "lib", "pfetch")
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True, timeout=60)
    p = run("nosuch")
    assert p.returncode == 1 and p.stderr.startswith('No such sampler "nosuch"\n')
    assert p.stderr.split("\n")[1:7] == ["normal", "boltzmann", "pymixture", "pyblasso", "py", "python"]
    p = run("py")
    assert p.returncode == 1 and "too few arguments, expected 'fetching py MODULE" in p.stderr
    p = run("py", "numpy")
    assert p.returncode == 1 and "No module named numpy" in p.stderr
    p = run("pymixture", "train_size")
    assert p.returncode == 1 and "expected key=value" in p.stderr
    p = run("-S", "pymixture", "train_size=12x")
    assert p.returncode == 1 and "not an integer" in p.stderr
    p = run("-f")
    assert p.returncode == 1
    p = run("-h")
    assert p.returncode == 0 and all(f in p.stdout for f in ("-b BATCH", "-y STYLE", "-u USEC", "-S NAME"))

"""The in-kernel peer exchange (pf_engine_peer_attach*, PeerXchg / peer_allreduce in csrc/pf_common.cuh).

On one GPU the ranks of a row-sharded group are engines of THIS process over shards of one data set,
attached with pf_engine_peer_attach_local: the protocol (peer stores -> fence -> flag, acquire, rank-order
fold, double-buffered slots) is exactly what runs across GPUs, only the pointers are local.  The
cross-process / cross-GPU form (CUDA IPC handles over torch.distributed) is scripts/multigpu_check.py.
Checker: the single-engine result over the whole data set, itself parity-tested against the oracle.
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.fixture(scope="module")
def fb():
    import torch
    assert torch.cuda.is_available()
    import fetching_b200
    from fetching_b200 import _lib
    _lib.load()
    return fetching_b200


def run_ranks(engs, thetas, calls=1):
    """One host thread per rank (the host path blocks on its completion flag)."""
    out = [None] * len(engs)
    err = []

    def work(i):
        try:
            for _ in range(calls):
                out[i] = engs[i].eval_frontier(thetas)
        except Exception as ex:  # noqa: BLE001
            err.append(ex)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(engs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    assert not err, err
    assert all(o is not None for o in out)
    return out


@pytest.mark.parametrize("G,B", [(2, 8), (4, 16), (3, 1), (8, 64), (2, 40)])
def test_mixture_row_shards_exchange(fb, G, B):
    from fetching_b200 import distributed as pfd, models
    data, pos = models.mixture_dataset(240_000, 8, 2, seed=11)
    thetas = models.mixture_thetas(pos, B, seed=B)
    with fb.FrontierEngine.mixture(data, 8, 1000) as full:
        fs, fq = full.eval_frontier(thetas)
    engs = []
    for r in range(G):
        lo, hi = pfd.shard_rows(len(data), 1000, r, G)
        engs.append(fb.FrontierEngine.mixture(data[lo:hi], 8, 1000))
    try:
        fb.peer_attach_local(engs)
        launches0 = [e.launch_count for e in engs]
        res = run_ranks(engs, thetas, calls=3)  # three calls: both slots and a slot reuse
        for e, l0 in zip(engs, launches0):
            # one kernel per call (two for a mixture frontier of more than 32 nodes): no collective kernel behind it
            assert e.launch_count - l0 == 3 * (2 if B > 32 else 1)
            e.peer_status()
        for s, q in res:
            assert rel(s, fs) < 1e-12 and rel(q, fq) < 1e-12
            assert np.array_equal(s, res[0][0]) and np.array_equal(q, res[0][1])  # rank-order fold: same bits
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("f32", [False, True])
def test_lasso_row_shards_exchange_device_path(fb, f32):
    """Device-pointer path: all ranks enqueue on their own streams, nothing blocks on the host.  f32: the tcgen05 kernel
    (both split teams run the finalize and the exchange)."""
    import torch
    from fetching_b200 import distributed as pfd, models
    G, B = 4, 9
    kw = dict(dtype=fb.F32) if f32 else {}
    tol = 1e-5 if f32 else 1e-12
    X, y, beta = models.blasso_dataset(20000, 96, seed=1)
    th = models.blasso_thetas(beta, B, seed=1)
    with fb.FrontierEngine.blasso(X, y, item_rows=500, **kw) as full:
        fs, fq = full.eval_frontier(th)
    engs = []
    for r in range(G):
        lo, hi = pfd.shard_rows(20000, 500, r, G)
        engs.append(fb.FrontierEngine.blasso(X[lo:hi], y[lo:hi], item_rows=500, **kw))
    try:
        fb.peer_attach_local(engs)
        d_th = torch.from_numpy(th).cuda()
        outs = [torch.zeros(2, B, dtype=torch.float64, device="cuda") for _ in range(G)]
        for it in range(4):
            for e, o in zip(engs, outs):
                e.eval_frontier_device(d_th.data_ptr(), o.data_ptr(), B, 0)
            # virtual ranks share one GPU: never leave a rank's NEXT call queued in front of another rank's
            # current one (streams may share a hardware queue); across GPUs every rank has its own device
            torch.cuda.synchronize()
        for e, o in zip(engs, outs):
            e.peer_status()
            got = o.cpu().numpy()
            assert rel(got[0], fs) < tol and rel(got[1], fq) < tol
            assert np.array_equal(got, outs[0].cpu().numpy())
    finally:
        for e in engs:
            e.close()


@pytest.mark.parametrize("G,B", [(2, 64), (4, 13), (8, 8), (3, 64), (4, 33)])
def test_node_split_gather(fb, G, B):
    """NODE SPLIT (small data set replicated on every rank): rank r scores nodes [r * ceil(B/G), ...) and the kernels
    gather each other's sums through the same exchange buffers."""
    from fetching_b200 import models
    data, pos = models.mixture_dataset(60_000, 8, 2, seed=4)
    thetas = models.mixture_thetas(pos, B, seed=B)
    with fb.FrontierEngine.mixture(data, 8, 1000) as full:
        fs, fq = full.eval_frontier(thetas)
    engs = [fb.FrontierEngine.mixture(data, 8, 1000) for _ in range(G)]
    try:
        fb.peer_attach_local(engs)
        for e in engs:
            e.set_option(fb.OPT_NODE_SPLIT, 1)
        l0 = [e.launch_count for e in engs]
        res = run_ranks(engs, thetas, calls=3)
        for e, l in zip(engs, l0):
            assert e.launch_count - l == 3
            e.peer_status()
        for s, q in res:
            # a node's sums come from exactly one kernel; a slice of fewer nodes runs another lane geometry than
            # the full frontier (rows are summed in a different order), so: equal to rounding, and identical bits
            # on every rank
            assert rel(s, fs) < 1e-13 and rel(q, fq) < 1e-13
            assert np.array_equal(s, res[0][0]) and np.array_equal(q, res[0][1])
    finally:
        for e in engs:
            e.close()


def test_node_split_gather_lasso_and_boltzmann(fb):
    from fetching_b200 import models
    X, y, beta = models.blasso_dataset(6000, 100, seed=1)
    th = models.blasso_thetas(beta, 24, seed=2, scale=0.05)
    W, Xs = models.boltzmann_dataset(256, 40, seed=3)
    for make, thetas, tol in ((lambda: fb.FrontierEngine.blasso(X, y, item_rows=500), th, 1e-12),
                              (lambda: fb.FrontierEngine.blasso(X, y, item_rows=500, dtype=fb.F32), th, 1e-5),  # tcgen05 kernel
                              (lambda: fb.FrontierEngine.boltzmann_dense(W), Xs, 1e-12)):
        with make() as full:
            fs, fq = full.eval_frontier(thetas)
        engs = [make() for _ in range(4)]
        try:
            fb.peer_attach_local(engs)
            for e in engs:
                e.set_option(fb.OPT_NODE_SPLIT, 1)
            res = run_ranks(engs, thetas, calls=2)
            for s, q in res:
                # the slice of a rank runs with a different node-block count than the full frontier: same IEEE
                # operations per node, possibly another summation order over columns / CTAs
                assert rel(s, fs) < tol and rel(q, fq) < tol
                assert np.array_equal(s, res[0][0])
        finally:
            for e in engs:
                e.close()


def test_exchange_option_needs_attachment(fb):
    from fetching_b200 import models
    data, pos = models.mixture_dataset(20_000, 8, 2, seed=0)
    with fb.FrontierEngine.mixture(data, 8, 1000) as e:
        with pytest.raises(fb.PfetchError):
            e.set_option(fb.OPT_PEER_EXCHANGE, 1)
    # an engine joins a group once
    engs = [fb.FrontierEngine.mixture(data, 8, 1000) for _ in range(2)]
    try:
        fb.peer_attach_local(engs)
        with pytest.raises(fb.PfetchError):
            fb.peer_attach_local(engs)
    finally:
        for e in engs:
            e.close()

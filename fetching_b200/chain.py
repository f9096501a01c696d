"""Run a whole prefetching Metropolis-Hastings chain through the engine (pf_chain_run).

Python face of the host side: the speculative tree, the sampler's proposal/prior logic and the
frontier hand-off all run in C++ (fetching_b200/csrc/host); this module only marshals the
configuration and collects the chain rows.  Equivalent of
`mpirun -np frontier+1 ./fetching -S <sampler> -l chain_length ...` (tree.cc:43-170).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from . import _lib
from ._lib import EVAL_FN, JOB_FN, RANGE_FN, ROW_FN, ChainConfig, ChainStats, check


@dataclass
class ChainResult:
    depth: np.ndarray          # [rows]
    accept: np.ndarray         # [rows]
    metric_sum: np.ndarray     # [rows]
    theta: np.ndarray          # [rows][theta_len]
    stats: dict
    jobs: List[tuple] = field(default_factory=list)  # (id, generation, state, path, rp, prop_rp, next_rp, log_prior, theta)
    status: int = 0            # pf_status of the run (only ever non-zero with check_status=False)

    @property
    def chain_steps_per_sec(self) -> float:
        return self.stats["levels"] / self.stats["seconds"] if self.stats["seconds"] > 0 else float("nan")

    @property
    def evals_per_sec(self) -> float:
        return self.stats["nodes_scored"] / self.stats["seconds"] if self.stats["seconds"] > 0 else float("nan")


def _config(frontier, chain_length, policy, correlation, reassign_ratio, dimension, threshold, verbose, print_tsv,
            seed, tau, initial_theta, deferred=False, batch_items=0, update_style=1):
    cfg = ChainConfig()
    cfg.frontier, cfg.scheduling_policy, cfg.chain_length = frontier, policy, chain_length
    cfg.correlation, cfg.reassign_ratio, cfg.dimension = correlation, reassign_ratio, dimension
    cfg.threshold, cfg.verbose, cfg.print_tsv, cfg.seed, cfg.tau = int(threshold), int(verbose), int(print_tsv), seed, tau
    cfg.deferred_proposals = int(deferred)
    cfg.batch_items, cfg.update_style = int(batch_items), int(update_style)
    keep = None
    if initial_theta is not None:
        keep = np.ascontiguousarray(initial_theta, dtype=np.float64)
        cfg.initial_theta = keep.ctypes.data_as(C.POINTER(C.c_double))
    return cfg, keep


def _collect(trace_jobs):
    rows, jobs = [], []

    def on_row(_u, depth, accept, metric_sum, theta, n):
        rows.append((depth, accept, metric_sum, np.ctypeslib.as_array(theta, shape=(n,)).copy()))

    def on_job(_u, nid, gen, state, path, rp, prp, nrp, log_prior, theta, n):
        jobs.append((nid, gen, state, path.decode(), rp, prp, nrp, log_prior,
                     np.ctypeslib.as_array(theta, shape=(n,)).copy()))

    return rows, jobs, ROW_FN(on_row), (JOB_FN(on_job) if trace_jobs else JOB_FN())


def _result(rows, jobs, st):
    stats = {k: getattr(st, k) for k, _ in ChainStats._fields_}
    if rows:
        theta = np.stack([r[3] for r in rows])
    else:
        theta = np.zeros((0, 0))
    return ChainResult(depth=np.array([r[0] for r in rows], dtype=np.int64),
                       accept=np.array([r[1] for r in rows], dtype=np.int64),
                       metric_sum=np.array([r[2] for r in rows]), theta=theta, stats=stats, jobs=jobs)


def run_chain(engine, chain_length: int = 100, frontier: int = 8, *, policy: int = 2, correlation: float = 0.99,
              reassign_ratio: float = 1.1, dimension: int = 1, threshold: bool = False, verbose: bool = False,
              print_tsv: bool = False, seed: int = 0, tau: float = 5.0, initial_theta=None,
              trace_jobs: bool = False, deferred_proposals: bool = False, batch_items: int = 0,
              update_style: int = 1, collect_rows: bool = True) -> ChainResult:
    """Defaults follow tree.cc:45-48 (-l 100 -r 1.1 -c 0.99 -d 1 -s 2).  batch_items > 0 turns on partial
    updates (the reference's -b / -y), i.e. predictive scheduling from partial sums.  collect_rows = False passes
    no row callback (a Python callback costs ~25 us per chain level, more than a small model's kernel pass):
    timing runs use it and read the acceptance count from the stats."""
    cfg, _keep = _config(frontier, chain_length, policy, correlation, reassign_ratio, dimension, threshold, verbose,
                         print_tsv, seed, tau, initial_theta, deferred_proposals, batch_items, update_style)
    rows, jobs, row_cb, job_cb = _collect(trace_jobs)
    if not collect_rows:
        row_cb = ROW_FN()
    st = ChainStats()
    check(_lib.load().pf_chain_run(engine._h, C.byref(cfg), row_cb, job_cb, None, C.byref(st)), "pf_chain_run")
    return _result(rows, jobs, st)


def run_chain_with_evaluator(model: int, dim: int, components: int, data_size: int,
                             evaluator: Callable[[np.ndarray], tuple], chain_length: int = 100, frontier: int = 8, *,
                             policy: int = 2, correlation: float = 0.99, reassign_ratio: float = 1.1,
                             dimension: int = 1, threshold: bool = False, verbose: bool = False, seed: int = 0,
                             tau: float = 5.0, initial_theta=None, trace_jobs: bool = False,
                             deferred_proposals: bool = False, batch_items: int = 0, update_style: int = 1,
                             range_evaluator=None, check_status: bool = True) -> ChainResult:
    """Same tree / sampler / hand-off, frontier scored by `evaluator(thetas[B][L]) -> (sum[B], sum_sq[B])`.
    No GPU involved: used to exercise the host logic (the CPU test-suite passes the oracle here).
    An evaluator may return an int instead: a negative pf_status, i.e. a failed engine call -- the chain must stop
    there (check_status=False returns the rows delivered so far with `.status` instead of raising)."""
    cfg, _keep = _config(frontier, chain_length, policy, correlation, reassign_ratio, dimension, threshold, verbose,
                         False, seed, tau, initial_theta, deferred_proposals, batch_items, update_style)
    rows, jobs, row_cb, job_cb = _collect(trace_jobs)
    theta_len = {1: dim + dim * (dim + 1) // 2, 2: 1, 3: dim, 4: dim * components, 5: dim + 1}[model]

    def on_eval(_u, B, thetas, s_out, q_out):
        th = np.ctypeslib.as_array(thetas, shape=(B, theta_len))
        r = evaluator(th)
        if isinstance(r, int):
            return r
        s, q = r
        np.ctypeslib.as_array(s_out, shape=(B,))[:] = s
        np.ctypeslib.as_array(q_out, shape=(B,))[:] = q
        return 0

    def on_range(_u, B, thetas, first, count, order, s_out, q_out):
        th = np.ctypeslib.as_array(thetas, shape=(B, theta_len))
        od = np.ctypeslib.as_array(order, shape=(B, 2))
        r = range_evaluator(th, int(first), int(count), od)
        if isinstance(r, int):
            return r
        s, q = r
        np.ctypeslib.as_array(s_out, shape=(B,))[:] = s
        np.ctypeslib.as_array(q_out, shape=(B,))[:] = q
        return 0

    st = ChainStats()
    range_cb = RANGE_FN(on_range) if range_evaluator is not None else RANGE_FN()
    rc = _lib.load().pf_chain_run_with_evaluator(model, dim, components, data_size, C.byref(cfg), EVAL_FN(on_eval),
                                                 range_cb, None, row_cb, job_cb, None, C.byref(st))
    if check_status:
        check(rc, "pf_chain_run_with_evaluator")
    res = _result(rows, jobs, st)
    res.status = int(rc)
    return res

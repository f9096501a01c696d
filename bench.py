#!/usr/bin/env python
"""bench.py -- frontier log-likelihood throughput on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--frontier B] [--dtype f64|f32]
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

One "step" = one pass of the hot path over one frontier: B speculative nodes scored over the
whole data set (pf_eval_frontier*, include/pfetch.h), plus the 2*B-double all-reduce when the
rows are sharded over several GPUs.  Default workload: BASELINE.json configs[4], the Gaussian
mixture K=8, d=2 with N = 10^8 rows in items of 1000 rows -- it fits one GPU (1.6 GB), so the
same job is run at every N (strong scaling; rows sharded contiguously at item boundaries).

`value`     loglik evals/s = frontier nodes fully scored per second, inputs resident in HBM,
            timed on the device with CUDA events, max over ranks.
`e2e`       the same through the host-buffer C-ABI call (pf_eval_frontier): theta H2D copy,
            kernel, all-reduce, result D2H copy, all inside the timed region.
`roofline`  dominant kernel, algorithmic bytes per launch / its measured duration vs the measured
            HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline`  the oracle's C restatement of the same arithmetic ("port"; the reference's own
            Python-2/numpy model code cannot run in this image) on all host threads, on a bounded
            row sample, extrapolated to the full row count.
Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BLOCKS = 8  # the synthetic data set is defined as 8 seeded blocks so every N in {1,2,4,8} sees the same rows

WORKLOADS = {
    # name: (description, kind, params)
    "mixture1e8": ("BASELINE configs[4]: Gaussian mixture K=8 d=2, N=1e8 rows, items of 1000 rows, rows sharded over GPUs",
                   "mixture", dict(rows=100_000_000, K=8, d=2, item_rows=1000)),
    "mixture1e6": ("BASELINE configs[2]: Gaussian mixture K=8 d=2, N=1e6 rows, items of 1000 rows",
                   "mixture", dict(rows=1_000_000, K=8, d=2, item_rows=1000)),
    "normal1e6": ("BASELINE configs[0]: normal target D=2, N=1e6 points (samplers/normal.cc recipe)",
                  "normal", dict(rows=1_000_000)),
    "blasso": ("BASELINE configs[3]: Bayesian lasso n=1e5 rows, p=1000, items of 1000 rows, FP64",
               "blasso", dict(rows=100_000, p=1000, item_rows=1000)),
    "boltzmann": ("BASELINE configs[1]: dense Boltzmann D=1024, frontier of 64 nodes",
                  "boltzmann", dict(D=1024)),
}


def measured_traffic(workload, dtype, B):
    """DRAM bytes per launch of the dominant kernel from an ncu --set full capture, if one was taken for
    this exact configuration (profiles/traffic.json, written from profiles/*_full.md); else None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("%s/%s/%d" % (workload, dtype, B))
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ data
def make_shard(kind, prm, rank, world):
    """Host arrays for this rank's contiguous row shard (item aligned), same rows for every `world`."""
    from fetching_b200 import models
    if kind == "mixture":
        rows, K, d = prm["rows"], prm["K"], prm["d"]
        rng = np.random.RandomState(12345)
        pos = 4.0 * rng.random_sample((K, d)) - 2.0
        if rows % (BLOCKS * prm["item_rows"]) == 0 and BLOCKS % world == 0:
            per_block = rows // BLOCKS
            mine = range(rank * BLOCKS // world, (rank + 1) * BLOCKS // world)
        else:
            assert world == 1
            per_block, mine = rows, [0]
        parts = []
        for b in mine:
            r = np.random.RandomState(1000 + b)
            comp = (r.random_sample(per_block) * K).astype(np.int64)
            parts.append(r.standard_normal((per_block, d)) + pos[comp])
        return dict(data=np.concatenate(parts) if len(parts) > 1 else parts[0], pos=pos)
    if kind == "normal":
        assert world == 1
        return dict(data=models.normal_dataset(prm["rows"]))
    if kind == "blasso":
        rows, p, ir = prm["rows"], prm["p"], prm["item_rows"]
        assert (rows // ir) % world == 0
        X, y, beta = models.blasso_dataset(rows, p, seed=0)
        lo, hi = rank * rows // world, (rank + 1) * rows // world
        return dict(data=X[lo:hi], y=y[lo:hi], beta=beta)
    if kind == "boltzmann":
        assert world == 1
        W, X = models.boltzmann_dataset(prm["D"], 64, seed=1)
        return dict(data=W, X=X)
    raise ValueError(kind)


def make_thetas(kind, prm, shard, B):
    from fetching_b200 import models
    if kind == "mixture":
        return models.mixture_thetas(shard["pos"], B, seed=1)
    if kind == "normal":
        t = models.normal_thetas(B, seed=1)
        t[0] = [0, 0, 1, 0, 1]
        return t
    if kind == "blasso":
        return models.blasso_thetas(shard["beta"], B, seed=1)
    if kind == "boltzmann":
        return np.ascontiguousarray(np.resize(shard["X"], (B, prm["D"])))
    raise ValueError(kind)


def make_engine(kind, prm, shard, dtype, device):
    import fetching_b200 as fb
    dt = fb.F32 if dtype == "f32" else fb.F64
    if kind == "mixture":
        return fb.FrontierEngine.mixture(shard["data"], prm["K"], prm["item_rows"], dtype=dt, device=device)
    if kind == "normal":
        return fb.FrontierEngine.normal(shard["data"], dtype=dt, device=device)
    if kind == "blasso":
        return fb.FrontierEngine.blasso(shard["data"], shard["y"], item_rows=prm["item_rows"], dtype=dt, device=device)
    if kind == "boltzmann":
        return fb.FrontierEngine.boltzmann_dense(shard["data"], dtype=dt, device=device)
    raise ValueError(kind)


def algorithmic_bytes(kind, prm, rows_local, B, dtype):
    """the data set is read once per call whatever B is."""
    T = 4 if dtype == "f32" else 8
    if kind == "mixture":
        return rows_local * prm["d"] * T + B * (prm["K"] * prm["d"] * 8 + 16)
    if kind == "normal":
        return rows_local * 2 * T + B * (5 * 8 + 16)
    if kind == "blasso":
        return rows_local * prm["p"] * T + rows_local * 8 + B * (prm["p"] + 1) * 8
    if kind == "boltzmann":
        return prm["D"] * prm["D"] * T + B * prm["D"] * 8 + 8 * B
    raise ValueError(kind)


def fp64_view(kind, prm, rows_local, B, dtype, kern_ms, device):
    """How busy the FP64 pipe is (MEASURED_PEAKS.json has no FP64 figure: the denominator is the FP64 FMA rate
    measured on this GPU by pf_debug_fp64_probe; the DMMA rate measured by pf_debug_dmma_probe is the same).
    Mixture (bound by this pipe for every B): FP64 instructions issued per (row, node) --
    100 for K=8, d=2, counted in the SASS of the inner loop (scripts/loopstat.sh) -- x rows x B.
    Lasso / dense Boltzmann: FMAs of the contraction (rows x cols x B), issued as DMMA m8n8k4."""
    if dtype != "f64":
        return None
    if kind == "mixture" and (prm["K"], prm["d"]) == (8, 2):
        instr = 100.0 * rows_local * B
        note = ("100 FP64 instructions per (row,node) from the inner-loop SASS; with the loop's 72 other instructions "
                "the dispatch ceiling of this mix is 200/272 = 0.74 of the pipe (an FP64 instruction holds the "
                "dispatch port 2 cycles, any other 1: scripts/fp64_probe.py)")
    elif kind == "blasso":
        instr = float(rows_local) * prm["p"] * B
        note = "FMAs of the rows x p x B contraction, issued as DMMA m8n8k4 (HBM is the binding unit for B <= 16)"
    elif kind == "boltzmann":
        instr = float(prm["D"]) * prm["D"] * B
        note = "FMAs of x'Wx for B nodes, issued as DMMA m8n8k4 (a ~20 us latency-bound call)"
    else:
        return None
    import ctypes as C
    from fetching_b200 import _lib
    out = np.zeros(16)
    _lib.check(_lib.load().pf_debug_fp64_probe(device, out.ctypes.data_as(C.POINTER(C.c_double))), "fp64 probe")
    peak_tfma = float(out.max())
    achieved = instr / (kern_ms * 1e-3) / 1e12
    return {"bound": "fp64", "achieved": achieved, "peak": peak_tfma, "unit": "T FP64 instr-lanes/s (1 FMA = 1)",
            "frac": achieved / peak_tfma, "peak_tflops": 2 * peak_tfma,
            "note": "peak measured by pf_debug_fp64_probe; " + note}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_leg(kind, prm, B, steps, warmup, budget_s=12.0):
    """The reference's CPU implementation of the path on the host cores, on a bounded row sample.

    mixture / lasso: the reference model code is Python 2 + numpy and cannot run in this image, so
    this is the oracle's C restatement of the same arithmetic ("port"), all host threads.  normal:
    oracle/_ref/pf_ref is the reference's own C++ ("reference") when the binary travelled."""
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    rows_total = prm.get("rows", prm.get("D", 1))
    if kind == "mixture":
        sample = min(rows_total, 2_000_000)
        shard = make_shard(kind, dict(prm, rows=sample), 0, 1)
        th = make_thetas(kind, prm, shard, B)
        fn = lambda: orc.mixture_frontier(shard["data"], th, prm["item_rows"])
        what = "oracle port (loglik_oracle.c, pymixture.py:86-92 arithmetic), OpenMP over node x item"
    elif kind == "blasso":
        sample = min(rows_total, 20_000)
        shard = make_shard(kind, dict(prm, rows=sample), 0, 1)
        th = make_thetas(kind, prm, shard, B)
        fn = lambda: orc.blasso_frontier(shard["data"], shard["y"], th, prm["item_rows"])
        what = "oracle port (loglik_oracle.c, pyblasso.py:135-141 arithmetic), OpenMP over node x item"
    elif kind == "normal":
        sample = rows_total
        if orc.have_pf_ref():
            # the reference's OWN NormalSampler<2>::evaluate_many (oracle/_ref/pf_ref, compiled in place from the
            # reference sources): one thread per node, like one MPI worker per node holding the whole data set
            reps = max(1, min(steps, 5))
            out = json.loads(orc.run_pf_ref("normal-bench", rows_total, B, reps, min(threads, B)).strip().splitlines()[-1])
            dt = out["seconds"] / reps
            return dict(value=B / dt, unit="evals/s", cores=int(out["threads"]), kind="reference",
                        sample="%d points x %d nodes per step, %d steps, %.3f s/step; reference NormalSampler<2>::"
                               "evaluate_many (normalsampler.hh:157-165) rebuilt against oracle/shim, not the stock MPI "
                               "binary; one thread per node" % (rows_total, B, reps, dt)), dt, reps
        shard = make_shard(kind, prm, 0, 1)
        th = make_thetas(kind, prm, shard, B)
        fn = lambda: orc.normal_frontier(shard["data"], th)
        what = "oracle port (loglik_oracle.c, normalsampler.hh:128-165 arithmetic), one thread per node"
    else:
        sample = rows_total
        shard = make_shard(kind, prm, 0, 1)
        th = make_thetas(kind, prm, shard, B)
        fn = lambda: orc.boltzmann_dense_frontier(shard["data"], th, 1.0)
        what = "oracle port (loglik_oracle.c, dense x'Wx definition), one thread per node"
    for _ in range(min(warmup, 1)):
        fn()
    t0 = time.perf_counter()
    done = 0
    while done < steps and (done == 0 or time.perf_counter() - t0 < budget_s):
        fn()
        done += 1
    dt = (time.perf_counter() - t0) / done
    scale = rows_total / sample
    evals = B / (dt * scale)
    return dict(value=evals, unit="evals/s", cores=threads, kind="port",
                sample="%d of %d rows x %d nodes per step, %d steps, %.3f s/step, scaled x%.0f to the full row count; %s"
                       % (sample, rows_total, B, done, dt, scale, what)), dt * scale, done


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mixture1e8", choices=sorted(WORKLOADS))
    ap.add_argument("--frontier", type=int, default=8, help="frontier nodes per step (B)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debugging only)")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1: reduce the per-node partials inside the frontier kernel over NVLink peer memory "
                         "(default) or with ncclAllReduce after it")
    ap.add_argument("--split", default="rows", choices=["rows", "nodes"],
                    help="N > 1: rows = item-aligned row shards (large N); nodes = every GPU holds the whole (small) data "
                         "set and scores ceil(B/N) nodes of the frontier (PF_OPT_NODE_SPLIT, in-kernel gather)")
    ap.add_argument("--chain-length", type=int, default=50, help="levels of the chain-steps/s measurement")
    args = ap.parse_args()

    desc, kind, prm = WORKLOADS[args.workload]
    prm = dict(prm)
    if args.rows:
        prm["rows"] = args.rows
    B = 64 if kind == "boltzmann" and args.frontier == 8 else args.frontier
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    config = {"workload": args.workload + ": " + desc, "frontier_nodes": B, "rows_total": prm.get("rows", prm.get("D")),
              "parallelism": ("row-shard x%d + %s of 2*B doubles" % (
                  world, "in-kernel NVLink peer exchange" if args.reduce == "peer" else "NCCL all-reduce"))
              if world > 1 else "single GPU",
              "l2": "per-GPU input larger than L2 (no flush needed)" if kind in ("mixture", "blasso") and
              prm.get("rows", 0) * (16 if kind == "mixture" else 8 * prm.get("p", 1)) / world > 126e6
              else "input fits L2: steady-state L2-resident re-reads, as in a running chain"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        base, sec_per_step, done = cpu_leg(kind, prm, B, K, W, budget_s=60.0)
        line = {"impl": "reference", "metric": "loglik_evals_per_sec", "value": base["value"], "unit": "evals/s",
                "n_gpus": args.gpus, "steps": done, "warmup": W, "ms_per_step": sec_per_step * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    import fetching_b200 as fb

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    node_split = world > 1 and args.split == "nodes"
    shard = make_shard(kind, prm, 0, 1) if node_split else make_shard(kind, prm, rank, world)
    rows_local = shard["data"].shape[0]
    thetas = make_thetas(kind, prm, shard, B)
    eng = make_engine(kind, prm, shard, args.dtype, local_rank)
    assert B <= eng.max_frontier, "frontier larger than one kernel pass; use <= %d" % eng.max_frontier
    if kind == "normal":
        eng.set_option(fb.OPT_ASSUME_IDENTITY_COV, 1)
    if world > 1:
        ids = [fb.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init(ids[0], rank, world)
        if args.reduce == "peer":
            from fetching_b200 import distributed as pfd
            try:
                pfd.init_engine_peers(eng)
                peer_ok = 1
            except fb.PfetchError as ex:  # e.g. peers hidden by CUDA_VISIBLE_DEVICES: cudaIpcOpenMemHandle fails
                print("peer exchange unavailable on rank %d (%s): NCCL all-reduce instead" % (rank, ex), file=sys.stderr)
                peer_ok = 0
            # all ranks must use the same reduction
            flag = torch.tensor([peer_ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if peer_ok:
                    eng.set_option(fb.OPT_PEER_EXCHANGE, 0)
                args.reduce = "nccl"
                config["parallelism"] = "row-shard x%d + NCCL all-reduce of 2*B doubles" % world
    if node_split:
        assert args.reduce == "peer" and world <= B <= 64, "node split: peers attached and N <= B <= 64"
        eng.set_option(fb.OPT_NODE_SPLIT, 1)
        config["parallelism"] = "node split x%d (replicated data, %d nodes per GPU, in-kernel NVLink gather)" % (
            world, -(-B // world))
    shard.pop("data")  # the engine owns the device copy; free the host rows

    # an explicit stream: handle 0 (the legacy default stream) means "the engine's own stream" in the C ABI
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d_th = torch.from_numpy(thetas).cuda()
    d_out = torch.zeros(2, B, dtype=torch.float64, device="cuda")

    def step_device():
        eng.eval_frontier_device(d_th.data_ptr(), d_out.data_ptr(), B, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step_device()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    clock_info = clocks.stop() if rank == 0 else None
    result_dev = d_out.cpu().numpy().copy()

    # ---- end to end through the host-buffer C ABI (what the master calls)
    # (K complete pf_eval_frontier calls issued from a C loop, pf_eval_frontier_repeat: what the reference's C++
    # master would pay per burst; a Python-level loop adds ~17 us of interpreter / ctypes time per call on top)
    for _ in range(W):
        eng.eval_frontier(thetas)
    barrier()
    s_host, q_host, t_e2e = eng.eval_frontier_repeat(thetas, K)
    torch.cuda.synchronize()
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    assert np.allclose(s_host, result_dev[0], rtol=1e-12, atol=0), "host and device paths disagree"

    # ---- dominant kernel: the step is that one launch (+ a <= 1 KB all-reduce when sharded), so the
    # CUDA-event time per step is its launch duration (an upper bound for N > 1)
    kern_ms = ms_total / K

    # ---- chain level: the speculative tree + sampler on the host, every scheduling burst scored by one
    # engine call (pf_chain_run).  All ranks run the same deterministic chain in lock-step.
    from fetching_b200 import chain as pf_chain
    chain_kw = dict(chain_length=args.chain_length, frontier=B)
    if kind == "mixture":
        chain_kw.update(dimension=prm["K"] * prm["d"])
    elif kind == "blasso":
        chain_kw.update(dimension=prm["p"] + 1, initial_theta=np.concatenate([[1.0], shard["beta"]]))
    elif kind == "boltzmann":
        chain_kw.update(dimension=prm["D"])
    def chain_run(**kw):
        barrier()
        r = pf_chain.run_chain(eng, **chain_kw, collect_rows=False, **kw)
        torch.cuda.synchronize()
        return {"steps_per_sec": r.chain_steps_per_sec, "levels": int(r.stats["levels"]),
                "engine_calls": int(r.stats["engine_calls"]), "node_passes": int(r.stats["nodes_scored"]),
                "node_item_evals": int(r.stats["items_scored"]), "dropped_early": int(r.stats["abandoned"]),
                "frontier": B, "accept_rate": float(r.stats["accepted"]) / max(1, int(r.stats["levels"]) - 1),
                "seconds": r.stats["seconds"], "eval_seconds": r.stats["eval_seconds"]}

    chain_info = chain_run()  # nodes scored whole: one kernel pass per scheduling burst
    if world == 1 and chain_info["seconds"] < 0.05:
        # a 50-level chain of a small model lasts a few ms, most of it start-up: lengthen it to ~0.2 s
        # (single process only: the length must be identical on every rank of a sharded run)
        chain_kw["chain_length"] = int(min(5000, args.chain_length * 0.2 / max(chain_info["seconds"], 1e-4)))
        chain_info = chain_run()
    n_items = eng.data_size * world if kind in ("mixture", "blasso") and world > 1 and not node_split else eng.data_size
    if eng.data_size >= 16 and world == 1:
        # predictive mode (the reference's -b / -y 0): a probe batch of 1/16 of the items, then the rest,
        # only for the nodes the tree still believes in
        chain_info["predictive"] = chain_run(batch_items=max(1, eng.data_size // 16), update_style=0)

    if rank == 0:
        pk, pk_src = peaks()
        B_kernel = -(-B // world) if node_split else B  # nodes one GPU's kernel scores
        abytes = algorithmic_bytes(kind, prm, rows_local, B_kernel, args.dtype)
        achieved = abytes / (kern_ms * 1e-3) / 1e9
        evals = B * K / (ms_total * 1e-3)
        line = {
            "metric": "loglik_evals_per_sec", "value": evals, "unit": "evals/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config,
            "item_evals_per_sec": evals * prm.get("rows", prm.get("D", 1)),
            "clocks": clock_info,
            "e2e": {"value": B * K / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": int(thetas.nbytes),
                    "d2h_bytes_per_step": int(2 * B * 8), "ms_per_step": t_e2e / K * 1e3},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"],
                         "traffic": measured_traffic(args.workload, args.dtype, B) if world == 1 else None,
                         "peak_source": pk_src,
                         "kernel": "pw_frontier_kernel / mv_frontier_kernel (per-GPU shard, %d rows)" % rows_local,
                         "algorithmic_bytes_per_launch": int(abytes), "kernel_ms": kern_ms},
        }
        line["chain"] = chain_info
        fp64 = fp64_view(kind, prm, rows_local, B_kernel, args.dtype, kern_ms, local_rank)
        if fp64:
            line["roofline"]["fp64_pipe"] = fp64
        # which unit actually bounds this configuration: the contract's "bound" stays "hbm"
        # (achieved / peak / frac above are the HBM view it asks for), this names the unit the time goes to
        per_gpu_bytes = abytes
        if kind == "mixture":
            binding = "fp64_pipe" if per_gpu_bytes > 64e6 else "fp64_pipe + call latency (small model)"
        elif kind == "blasso":
            binding = "hbm" if B_kernel <= 16 else "fp64_tensor (DMMA)"
        else:
            binding = "call latency (L2-resident data, tens of microseconds per call)"
        line["roofline"]["binding_unit"] = binding
        if world == 1:
            base, _, _ = cpu_leg(kind, prm, B, 3, 1)
            line["cpu_baseline"] = base
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

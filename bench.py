#!/usr/bin/env python
"""bench.py -- frontier log-likelihood throughput on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--frontier B] [--dtype f64|f32]
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

One "step" = one pass of the hot path over one frontier: B speculative nodes scored over the
whole data set (pf_eval_frontier*, include/pfetch.h), including the exchange of the 2*B partial sums
when the rows are sharded over several GPUs.  Default workload: BASELINE.json configs[4], the Gaussian
mixture K=8, d=2 with N = 10^8 rows in items of 1000 rows -- it fits one GPU (1.6 GB), so the
same job is run at every N (strong scaling; rows sharded contiguously at item boundaries).

`value`     loglik evals/s = frontier nodes fully scored per second, inputs resident in HBM,
            timed on the device with CUDA events, max over ranks.
`e2e`       the same through the host-buffer C-ABI call (pf_eval_frontier): theta from host memory,
            kernel (+ exchange), results back in host memory, all inside the timed region.
`roofline`  dominant kernel against the unit that bounds it (`bound`): the FP64 pipe for the mixture
            (peak measured on this GPU by pf_debug_fp64_probe), HBM for the lasso at B <= 16
            (peak = MEASURED_PEAKS.json); `roofline.hbm` is always the HBM view of the same launch.
`checksum`  the per-node sums themselves: the data set is the same 8 seeded blocks at every N, so the lines of
            N = 1, 2, 4, 8 can be compared with each other.
`parity`    checks made inside this run: a window of every rank's shard against the oracle, bit-identity of the
            reduced sums across ranks, host path == device path.
`configs`   (N = 1) one short measurement per BASELINE config C1..C4 (+ the PF_F32 lasso on tcgen05), each checked against the oracle.
`scale_extra`, `node_split`  the multi-GPU cases where the exchange overhead shows: B = 1, lasso row shards,
            node split of the 10^6-row mixture.
`cpu_baseline`  the oracle's C restatement of the same arithmetic ("port"; the reference's own
            Python-2/numpy model code cannot run in this image) on ALL host threads (set explicitly: an
            inherited OMP_NUM_THREADS is ignored), plus the numpy statement of the reference model on one core.
Only the cpu_baseline / --impl reference legs and the parity checks touch oracle/ (as the checker, never as the
thing measured).
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

# FP64 instructions the mixture kernel issues per (row, node) for K = 8, d = 2, counted in the SASS of its inner
# loop (scripts/loopstat.sh -> profiles/r2_loopstat.txt); the denominator of the FP64-pipe roofline.
MIX_FP64_INSTR = 75.0  # (81 before the product form: 72 for the eight exponentials, 2 for |x|^2, 1 multiply)

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


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


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
        from fetching_b200 import distributed as pfd
        rows, p, ir = prm["rows"], prm["p"], prm["item_rows"]
        X, y, beta = models.blasso_dataset(rows, p, seed=0)
        lo, hi = pfd.shard_rows(rows, ir, rank, world)  # item aligned, as even as possible
        return dict(data=np.ascontiguousarray(X[lo:hi]), y=np.ascontiguousarray(y[lo:hi]), beta=beta)
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
        eng = fb.FrontierEngine.normal(shard["data"], dtype=dt, device=device)
        eng.set_option(fb.OPT_ASSUME_IDENTITY_COV, 1)
        return eng
    if kind == "blasso":
        return fb.FrontierEngine.blasso(shard["data"], shard["y"], item_rows=prm["item_rows"], dtype=dt, device=device)
    if kind == "boltzmann":
        return fb.FrontierEngine.boltzmann_dense(shard["data"], dtype=dt, device=device)
    raise ValueError(kind)


def oracle_frontier(kind, prm, shard, thetas, rows=None):
    """The checker: oracle sums over the first `rows` rows of `shard` (all of it by default)."""
    from oracle import oracle as orc
    if kind == "mixture":
        return orc.mixture_frontier(shard["data"][:rows], thetas, prm["item_rows"])
    if kind == "blasso":
        return orc.blasso_frontier(shard["data"][:rows], shard["y"][:rows], thetas, prm["item_rows"])
    if kind == "normal":
        return orc.normal_frontier(shard["data"][:rows], thetas)
    return orc.boltzmann_dense_frontier(shard["data"], thetas, 1.0)


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


_FP64_PEAK = {}


def fp64_peak(device):
    """FP64 FMA issue rate (1e12 warp-lane instructions / s) measured on this GPU by pf_debug_fp64_probe
    (MEASURED_PEAKS.json has no FP64 figure; the DMMA rate measured by pf_debug_dmma_probe is the same)."""
    if device not in _FP64_PEAK:
        import ctypes as C
        from fetching_b200 import _lib
        out = np.zeros(16)
        _lib.check(_lib.load().pf_debug_fp64_probe(device, out.ctypes.data_as(C.POINTER(C.c_double))), "fp64 probe")
        _FP64_PEAK[device] = float(out.max())
    return _FP64_PEAK[device]


def compute_view(kind, prm, rows_local, B, dtype, kern_ms, device, sm_mhz=None):
    """The arithmetic unit's view of one launch, or None when the model has none worth stating.
    Mixture f64: FP64 instructions issued per (row, node) x rows x B against the measured FP64 issue rate.
    Lasso / dense Boltzmann f64: FMAs of the contraction (rows x cols x B), issued as DMMA m8n8k4, same peak.
    Mixture f32: K exp2 + 1 log2 on the special-function unit (MUFU, 16 per clock per SM, nominal)."""
    if dtype == "f32":
        if kind != "mixture":
            return None
        mufu = (prm["K"] + 1.0) * rows_local * B
        peak = 148 * 16 * (sm_mhz or 1965.0) * 1e6 / 1e12
        ach = mufu / (kern_ms * 1e-3) / 1e12
        return {"bound": "mufu", "achieved": ach, "peak": peak, "unit": "T MUFU op/s", "frac": ach / peak,
                "peak_source": "nominal: 148 SMs x 16 MUFU/clk x SM clock",
                "note": "%d MUFU.EX2 + 1 MUFU.LG2 per (row, node)" % prm["K"]}
    if kind == "mixture" and (prm["K"], prm["d"]) == (8, 2):
        instr = MIX_FP64_INSTR * rows_local * B
        note = ("%.0f FP64 instructions per (row,node) in the inner-loop SASS (profiles/r2_loopstat.txt); an FP64 "
                "instruction holds the dispatch port 2 cycles, any other 1 (scripts/fp64_probe.py)" % MIX_FP64_INSTR)
    elif kind == "blasso":
        instr = float(rows_local) * prm["p"] * B
        note = "FMAs of the rows x p x B contraction, issued as DMMA m8n8k4"
    elif kind == "boltzmann":
        instr = float(prm["D"]) * prm["D"] * B
        note = "FMAs of x'Wx for B nodes, issued as DMMA m8n8k4"
    else:
        return None
    peak = fp64_peak(device)
    ach = instr / (kern_ms * 1e-3) / 1e12
    return {"bound": "fp64" if kind == "mixture" else "tensor", "achieved": ach, "peak": peak,
            "unit": "T FP64 instr-lanes/s (1 FMA = 1)", "frac": ach / peak, "peak_tflops": 2 * peak,
            "peak_source": "pf_debug_fp64_probe (measured on this GPU)", "note": note}


def roofline_record(workload, kind, prm, rows_local, B, dtype, kern_ms, device, world, sm_mhz=None):
    """`bound / achieved / peak / frac` describe the unit that binds this configuration; `hbm` is always there."""
    pk, pk_src = peaks()
    abytes = algorithmic_bytes(kind, prm, rows_local, B, dtype)
    ach = abytes / (kern_ms * 1e-3) / 1e9
    hbm = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
           "peak_source": pk_src}
    cv = compute_view(kind, prm, rows_local, B, dtype, kern_ms, device, sm_mhz)
    # which unit the time goes to (DESIGN.md sections 4 and 6)
    small = abytes < 64e6
    if kind == "mixture":
        binding, note = cv, ("FP64 pipe" if dtype == "f64" else "special-function unit") + \
            (" + call latency (L2-resident data)" if small else "")
    elif kind == "blasso" and B > 16 and cv:
        binding, note = cv, "FP64 tensor instruction (DMMA)"
    elif kind == "blasso":
        binding, note = hbm, "HBM"
    else:
        binding, note = hbm, "call latency (L2-resident data, tens of microseconds per call); HBM view reported"
    rec = dict(binding or hbm)
    rec.update({"binding_unit": note, "hbm": hbm,
                "traffic": measured_traffic(workload, dtype, B) if world == 1 else None,
                "kernel": "pw_frontier_kernel / mv_frontier_kernel (per-GPU shard, %d rows)" % rows_local,
                "algorithmic_bytes_per_launch": int(abytes), "kernel_ms": kern_ms})
    if cv and binding is not cv:
        rec["compute"] = cv
    return rec


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
def host_cores():
    """All hardware threads this process may use -- NOT what OMP_NUM_THREADS says (torchrun exports 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_leg(kind, prm, B, steps, warmup, budget_s):
    """The reference's CPU implementation of the path on the host cores, timed for `steps` steps after `warmup`.

    mixture / lasso: the reference model code is Python 2 + numpy and cannot run in this image, so this is the
    oracle's C restatement of the same arithmetic ("port"), OpenMP over node x item on ALL host threads.  normal:
    oracle/_ref/pf_ref is the reference's own C++ ("reference") when the binary travelled.
    The whole data set is scored when (warmup + steps) passes fit `budget_s`; otherwise every step scores a
    contiguous sample of whole items and the throughput is scaled to the full row count -- `ms_per_step` is always
    the measured time of what ran.  Returns (cpu_baseline record, measured seconds per step, steps done)."""
    from oracle import oracle as orc
    orc.build()
    threads = orc.set_threads(host_cores())
    rows_total = prm.get("rows", prm.get("D", 1))
    if kind == "normal" and orc.have_pf_ref():
        # the reference's OWN NormalSampler<2>::evaluate_many (oracle/_ref/pf_ref, compiled in place from the
        # reference sources): one thread per node, like one MPI worker per node holding the whole data set
        reps = max(1, steps)
        use = min(threads, B)
        out = json.loads(orc.run_pf_ref("normal-bench", rows_total, B, reps, use).strip().splitlines()[-1])
        dt = out["seconds"] / reps
        return dict(value=B / dt, unit="evals/s", cores=int(out["threads"]), kind="reference",
                    sample="all %d points x %d nodes per step, %d steps, %.4f s/step; reference NormalSampler<2>::"
                           "evaluate_many (normalsampler.hh:157-165) rebuilt against oracle/shim, not the stock MPI "
                           "binary; one thread per node" % (rows_total, B, reps, dt)), dt, reps

    def build_fn(rows):
        sh = make_shard(kind, dict(prm, rows=rows), 0, 1) if "rows" in prm else make_shard(kind, prm, 0, 1)
        th = make_thetas(kind, prm, sh, B)
        if kind == "mixture":
            return lambda: orc.mixture_frontier(sh["data"], th, prm["item_rows"])
        if kind == "blasso":
            return lambda: orc.blasso_frontier(sh["data"], sh["y"], th, prm["item_rows"])
        if kind == "normal":
            return lambda: orc.normal_frontier(sh["data"], th)
        return lambda: orc.boltzmann_dense_frontier(sh["data"], th, 1.0)

    what = {"mixture": "oracle port (loglik_oracle.c, pymixture.py:86-92 arithmetic), OpenMP over node x item",
            "blasso": "oracle port (loglik_oracle.c, pyblasso.py:135-141 arithmetic), OpenMP over node x item",
            "normal": "oracle port (loglik_oracle.c, normalsampler.hh:128-165 arithmetic), one thread per node",
            "boltzmann": "oracle port (loglik_oracle.c, dense x'Wx definition), one thread per node"}[kind]
    sample = rows_total
    if kind in ("mixture", "blasso"):
        # calibrate on a small sample, then take as many rows as (warmup + steps) passes of the budget allow
        ir = prm["item_rows"]
        cal_rows = min(rows_total, max(ir, (1_000_000 if kind == "mixture" else 10_000) // ir * ir))
        cal = build_fn(cal_rows)
        cal()
        t0 = time.perf_counter()
        cal()
        per_row = (time.perf_counter() - t0) / cal_rows
        afford = budget_s / max(1, warmup + steps) / per_row
        if afford < rows_total:
            sample = int(max(cal_rows, min(rows_total, afford) // ir * ir))
        fn = cal if sample == cal_rows else build_fn(sample)
    else:
        fn = build_fn(rows_total)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    done = 0
    while done < steps and (done == 0 or time.perf_counter() - t0 < 2.0 * budget_s):
        fn()
        done += 1
    dt = (time.perf_counter() - t0) / done
    scale = rows_total / sample
    whole = "all" if sample == rows_total else "%d of" % sample
    return dict(value=B / (dt * scale), unit="evals/s", cores=threads, kind="port",
                sample="%s %d rows x %d nodes per step, %d steps after %d warm-up, %.3f s/step measured%s; %s"
                       % (whole, rows_total, B, done, warmup, dt,
                          "" if scale == 1.0 else ", throughput scaled x%.1f to the full row count" % scale, what)), dt, done


def numpy_one_core(kind, prm, budget_s=3.0):
    """BASELINE.md section 3, baseline 2(a): the numpy statement of the reference's model code -- what a reference
    worker really executes per item (python.cc:185 -> pymixture.py:86-92 / pyblasso.py:135-141) -- on one core, on
    a bounded number of items, scaled to one whole-data evaluation of one node."""
    from oracle import oracle as orc
    if kind not in ("mixture", "blasso"):
        return None
    ir = prm["item_rows"]
    n_items_total = prm["rows"] // ir
    sh = make_shard(kind, dict(prm, rows=min(prm["rows"], 200 * ir)), 0, 1)
    th = make_thetas(kind, prm, sh, 1)[0]
    n_have = sh["data"].shape[0] // ir
    try:
        from threadpoolctl import threadpool_limits
        limit = threadpool_limits(limits=1)
    except Exception:
        limit = None
    try:
        if kind == "mixture":
            th2 = th.reshape(prm["K"], prm["d"])
            item = lambda i: orc.np_mixture_evaluate(sh["data"], th2, ir, i)
        else:
            item = lambda i: orc.np_blasso_evaluate(sh["data"], sh["y"], th, i, ir)
        item(0)
        t0, done = time.perf_counter(), 0
        while done < n_have and (done == 0 or time.perf_counter() - t0 < budget_s):
            item(done)
            done += 1
        per_item = (time.perf_counter() - t0) / done
    finally:
        if limit is not None:
            limit.restore_original_limits()
    return {"value": 1.0 / (per_item * n_items_total), "unit": "evals/s", "cores": 1,
            "sample": "%d items of %d rows timed, %.1f us per item, scaled to the %d items of one node" %
                      (done, ir, per_item * 1e6, n_items_total)}


def reference_mpirun_chain(levels=300):
    """BASELINE.md section 3, baseline 1(b), for C1: the reference's own parallel program -- JobTree::run_master on rank
    0, the registered normal sampler's work() -> worker<NormalSampler<2>>() on ranks 1..np-1, N = 10^6 points per
    worker (samplers/normal.cc) -- compiled in place from the reference sources with THREADS for ranks
    (`oracle/_ref/pf_ref mpirun`, thread-mailbox boost::mpi shim), np = all host threads.  Chain steps/s from the
    time stamps of the TSV it prints (first to last level, data-set construction excluded) and the stock
    `OK sample/s` figure (everything included).  None when the binary did not travel."""
    from oracle import oracle as orc
    if not orc.have_pf_ref():
        return None
    np_ranks = max(2, host_cores())
    try:
        p = subprocess.run([orc.REF_BIN, "mpirun", "-np", str(np_ranks), "normal", "-l", str(levels)],
                           capture_output=True, text=True, timeout=600)
    except (OSError, subprocess.TimeoutExpired):
        return None
    rows = [l.split("\t") for l in p.stdout.splitlines() if l[:1].isdigit()]
    ok = [l for l in p.stderr.splitlines() if l.startswith("OK sample/s")]
    if p.returncode != 0 or len(rows) < 3 or not ok:
        return None
    t0, t1 = float(rows[1][0]), float(rows[-1][0])
    return {"steps_per_sec": (int(rows[-1][1]) - int(rows[1][1])) / max(t1 - t0, 1e-9), "np": np_ranks,
            "levels": int(rows[-1][1]), "ok_line_samples_per_sec": float(ok[0].split("=")[1].split()[0]),
            "accept_rate": sum(int(r[2]) for r in rows[1:]) / max(1, len(rows) - 1),
            "kind": "reference source (master.cc, masterloop.cc, worker.hh, samplers/normal.cc) rebuilt against a shim, "
                    "threads for MPI ranks; not the stock Boost.MPI binary"}


# ------------------------------------------------------------------------------------------ GPU helpers
class Ctx:
    """torch / torch.distributed handles + this process's place in the job."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        assert self.world == args.gpus or self.world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
        # an explicit stream: handle 0 (the legacy default stream) means "the engine's own stream" in the C ABI
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def all_agree(self, ok):
        t = self.torch.tensor([1 if ok else 0], device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item()) == 1

    def identical_across_ranks(self, a):
        """True when every rank holds the same bits."""
        t = self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        if self.world == 1:
            return True
        got = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(got, t)
        return self.all_agree(all(bool(self.torch.equal(g.view(self.torch.int64), t.view(self.torch.int64))) for g in got))


def attach_group(ctx, eng, reduce="peer"):
    """Put `eng` into the job's group: NCCL communicator + (reduce == "peer") the in-kernel NVLink exchange.
    Returns the reduction in force ("peer" or "nccl") -- the same on every rank."""
    import fetching_b200 as fb
    from fetching_b200 import distributed as pfd
    if ctx.world == 1:
        return "none"
    ids = [fb.nccl_unique_id() if ctx.rank == 0 else None]
    ctx.dist.broadcast_object_list(ids, src=0)
    eng.comm_init(ids[0], ctx.rank, ctx.world)
    if reduce != "peer":
        return "nccl"
    try:
        pfd.init_engine_peers(eng)
        ok = True
    except fb.PfetchError as ex:  # e.g. peers hidden by CUDA_VISIBLE_DEVICES: cudaIpcOpenMemHandle fails
        print("peer exchange unavailable on rank %d (%s): NCCL all-reduce instead" % (ctx.rank, ex), file=sys.stderr)
        ok = False
    if ctx.all_agree(ok):
        return "peer"
    if ok:
        eng.set_option(fb.OPT_PEER_EXCHANGE, 0)
    return "nccl"


def time_engine(ctx, eng, thetas, B, K, W):
    """W warm-up + K timed steps on the device-pointer path (CUDA events on the launching stream, barrier +
    synchronize on both sides, max over ranks), then the same through the host-buffer C-ABI call issued from a C
    loop (pf_eval_frontier_repeat: what the reference's C++ master would pay per burst).
    -> dict(ms_total, launches, dev[2][B], host_s, host_q, e2e_s)"""
    torch = ctx.torch
    d_th = torch.from_numpy(np.ascontiguousarray(thetas)).cuda()
    d_out = torch.zeros(2, B, dtype=torch.float64, device="cuda")
    step = lambda: eng.eval_frontier_device(d_th.data_ptr(), d_out.data_ptr(), B, ctx.stream.cuda_stream)
    for _ in range(W):
        step()
    ctx.barrier()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record(ctx.stream)
    for _ in range(K):
        step()
    ev1.record(ctx.stream)
    ctx.barrier()
    ms_total = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count - launches0
    dev = d_out.cpu().numpy().copy()
    for _ in range(W):
        eng.eval_frontier(thetas)
    ctx.barrier()
    # the step's inputs start in PINNED host memory (the engine's staging buffer: what a master that assembles the
    # burst's proposals into one buffer hands over) and travel to the device inside the timed call
    s_host, q_host, t_e2e = eng.eval_frontier_repeat(thetas, K, pinned=True)
    torch.cuda.synchronize()
    return dict(ms_total=ms_total, launches=int(launches), dev=dev, host_s=s_host, host_q=q_host,
                e2e_s=ctx.max_over_ranks(t_e2e))


def chain_kwargs(kind, prm, shard, B, length):
    kw = dict(chain_length=length, frontier=B)
    if kind == "mixture":
        kw.update(dimension=prm["K"] * prm["d"])
    elif kind == "blasso":
        kw.update(dimension=prm["p"] + 1, initial_theta=np.concatenate([[1.0], shard["beta"]]))
    elif kind == "boltzmann":
        kw.update(dimension=prm["D"])
    return kw


def chain_run(ctx, eng, kw, **more):
    """pf_chain_run: the speculative tree + sampler on the host, every scheduling burst scored by one engine call.
    All ranks of a sharded group run the same deterministic chain, of the same length, in lock-step."""
    from fetching_b200 import chain as pf_chain
    ctx.barrier()
    r = pf_chain.run_chain(eng, **kw, collect_rows=False, **more)
    ctx.torch.cuda.synchronize()
    st = r.stats
    return {"steps_per_sec": r.chain_steps_per_sec, "levels": int(st["levels"]), "engine_calls": int(st["engine_calls"]),
            "node_passes": int(st["nodes_scored"]), "node_item_evals": int(st["items_scored"]),
            "dropped_early": int(st["abandoned"]), "frontier": kw["frontier"],
            "accept_rate": float(st["accepted"]) / max(1, int(st["levels"]) - 1), "seconds": st["seconds"],
            "eval_seconds": st["eval_seconds"],
            # how far the closest accept / reject decision was from flipping (absolute, and relative to |sum|)
            "min_margin": st["min_margin"], "min_margin_rel": st["min_margin_rel"]}


def tol_of(dtype):
    return 1e-5 if dtype == "f32" else 1e-10


def sum_err(kind, got, want):
    """Relative error of per-node sums.  The dense Boltzmann energy x'Wx is a sum of ~D^2 terms of either sign that
    can land anywhere near zero, so its error is taken relative to max(|want|, 1) (as tests/test_gpu_parity.py does)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    floor = 1.0 if kind == "boltzmann" else 1e-300
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor)))


def small_config(ctx, name, B, dtype, K, W, chain_levels):
    """One BASELINE config on ONE GPU (this rank's), checked against the oracle over the whole data set."""
    desc, kind, prm = WORKLOADS[name]
    prm = dict(prm)
    shard = make_shard(kind, prm, 0, 1)
    thetas = make_thetas(kind, prm, shard, B)
    eng = make_engine(kind, prm, shard, dtype, ctx.local_rank)
    rows = shard["data"].shape[0]
    m = time_engine(ctx, eng, thetas, B, K, W)
    want = oracle_frontier(kind, prm, shard, thetas)
    err = max(sum_err(kind, m["dev"][0], want[0]), sum_err(kind, m["dev"][1], want[1]),
              sum_err(kind, m["host_s"], want[0]))
    assert err < tol_of(dtype), "%s: GPU vs oracle relative error %.3e" % (name, err)
    kern_ms = m["ms_total"] / K
    out = {"workload": desc, "frontier_nodes": B, "ms_per_step": kern_ms, "evals_per_sec": B / (kern_ms * 1e-3),
           "e2e_evals_per_sec": B * K / m["e2e_s"], "e2e_us_per_call": m["e2e_s"] / K * 1e6, "steps": K,
           "gpu_launches": m["launches"], "oracle_rel_err": err,
           "roofline": roofline_record(name, kind, prm, rows, B, dtype, kern_ms, ctx.local_rank, 1)}
    if chain_levels:
        out["chain"] = chain_run(ctx, eng, chain_kwargs(kind, prm, shard, B, chain_levels))
    eng.close()
    if name == "normal1e6":
        ref = reference_mpirun_chain()
        if ref:
            out["reference_cpu_chain"] = ref
    return out


def node_split_extra(ctx, K, W):
    """Node split (small N): the 10^6-row mixture replicated on every GPU, a frontier of 64 nodes divided over the
    ranks, sums gathered inside the frontier kernels over NVLink.  Checked against the same engine un-split."""
    import fetching_b200 as fb
    desc, kind, prm = WORKLOADS["mixture1e6"]
    B = 64
    shard = make_shard(kind, prm, 0, 1)
    thetas = make_thetas(kind, prm, shard, B)
    eng = make_engine(kind, prm, shard, "f64", ctx.local_rank)
    single = time_engine(ctx, eng, thetas, B, K, W)  # every rank alone, all 64 nodes
    if attach_group(ctx, eng, "peer") != "peer":
        eng.close()
        return {"unavailable": "peer exchange not attachable on this box"}
    eng.set_option(fb.OPT_NODE_SPLIT, 1)
    m = time_engine(ctx, eng, thetas, B, K, W)
    err = max(rel_err(m["dev"][0], single["dev"][0]), rel_err(m["dev"][1], single["dev"][1]),
              rel_err(m["host_s"], single["dev"][0]), rel_err(m["host_q"], single["dev"][1]))
    same = ctx.identical_across_ranks(m["dev"])
    assert err < 1e-12 and same, "node split disagrees with the un-split engine: %.3e (identical on all ranks: %s)" % (err, same)
    eng.close()
    ms, ms1 = m["ms_total"] / K, single["ms_total"] / K
    return {"workload": desc, "frontier_nodes": B, "nodes_per_gpu": -(-B // ctx.world), "ms_per_step": ms,
            "evals_per_sec": B / (ms * 1e-3), "e2e_evals_per_sec": B * K / m["e2e_s"],
            "one_gpu_ms_per_step": ms1, "speedup_vs_one_gpu": ms1 / ms, "steps": K,
            "vs_unsplit_rel_err": err, "ranks_bit_identical": same}


def lasso_extra(ctx, K, W, reduce):
    """Lasso C4 row-sharded over the job's GPUs (12.5 k rows per GPU at N = 8: ~17 us of HBM time per call, so the
    exchange and the call overhead are what is measured), checked against the oracle on every rank's shard."""
    desc, kind, prm = WORKLOADS["blasso"]
    B = 8
    shard = make_shard(kind, prm, ctx.rank, ctx.world)
    thetas = make_thetas(kind, prm, shard, B)
    eng = make_engine(kind, prm, shard, "f64", ctx.local_rank)
    red = attach_group(ctx, eng, reduce)
    m = time_engine(ctx, eng, thetas, B, K, W)
    want = oracle_frontier(kind, prm, shard, thetas)
    want = ctx.sum_over_ranks(np.stack(want))
    err = max(rel_err(m["dev"][0], want[0]), rel_err(m["dev"][1], want[1]), rel_err(m["host_s"], want[0]))
    same = ctx.identical_across_ranks(m["dev"])
    assert err < 1e-10 and same, "lasso row shards: relative error %.3e vs the oracle (identical on all ranks: %s)" % (err, same)
    eng.close()
    ms = m["ms_total"] / K
    return {"workload": desc, "frontier_nodes": B, "rows_per_gpu": int(shard["data"].shape[0]), "reduce": red,
            "ms_per_step": ms, "evals_per_sec": B / (ms * 1e-3), "e2e_evals_per_sec": B * K / m["e2e_s"], "steps": K,
            "oracle_rel_err": err, "ranks_bit_identical": same,
            "checksum": {"sum": [float(x) for x in m["dev"][0]]}}


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
    ap.add_argument("--chain-length", type=int, default=2000,
                    help="levels of the chain-steps/s measurement (the same on every rank)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip `configs` / `scale_extra` / `node_split` (sweeps and profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()

    desc, kind, prm = WORKLOADS[args.workload]
    prm = dict(prm)
    if args.rows:
        prm["rows"] = args.rows
    B = 64 if kind == "boltzmann" and args.frontier == 8 else args.frontier
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    default_line = args.workload == "mixture1e8" and args.dtype == "f64" and not args.rows and args.split == "rows"
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
        base, sec_per_step, done = cpu_leg(kind, prm, B, K, W, budget_s=200.0)
        line = {"impl": "reference", "metric": "loglik_evals_per_sec", "value": base["value"], "unit": "evals/s",
                "n_gpus": args.gpus, "steps": done, "warmup": W, "ms_per_step": sec_per_step * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    # the CPU baseline first (rank 0 of a single-GPU run), before any GPU work: same leg as the reference arm,
    # bounded to ~15 s of CPU work
    base = None
    if world == 1 and not args.no_cpu:
        base, _, _ = cpu_leg(kind, prm, B, 3, 1, budget_s=8.0)
        one = numpy_one_core(kind, prm)
        if one:
            base["numpy_1core"] = one

    import fetching_b200 as fb
    ctx = Ctx(args)
    torch = ctx.torch

    node_split = world > 1 and args.split == "nodes"
    shard = make_shard(kind, prm, 0, 1) if node_split else make_shard(kind, prm, rank, world)
    rows_local = shard["data"].shape[0]
    thetas = make_thetas(kind, prm, shard, B)
    eng = make_engine(kind, prm, shard, args.dtype, ctx.local_rank)
    assert B <= eng.max_frontier, "frontier larger than one kernel pass; use <= %d" % eng.max_frontier
    if world > 1:
        args.reduce = attach_group(ctx, eng, args.reduce)
        if args.reduce == "nccl":
            config["parallelism"] = "row-shard x%d + NCCL all-reduce of 2*B doubles" % world
    if node_split:
        assert args.reduce == "peer" and world <= B <= 64, "node split: peers attached and N <= B <= 64"
        eng.set_option(fb.OPT_NODE_SPLIT, 1)
        config["parallelism"] = "node split x%d (replicated data, %d nodes per GPU, in-kernel NVLink gather)" % (
            world, -(-B // world))

    # ---- parity, part 1: a window of EVERY rank's shard against the oracle (summed over the ranks like the engine
    # sums it: the range call of a peer group exchanges too).  Whole data set for the small single-GPU workloads.
    parity = {"tolerance": tol_of(args.dtype)}
    if not node_split:
        if kind in ("mixture", "blasso"):
            n_win = min(1000 if kind == "mixture" else 10, eng.data_size)
            win_rows = n_win * prm["item_rows"]
            got = eng.eval_frontier_range(thetas, 0, n_win)
        else:
            n_win, win_rows = eng.data_size, None
            got = eng.eval_frontier(thetas)
        want = ctx.sum_over_ranks(np.stack(oracle_frontier(kind, prm, shard, thetas, win_rows)))
        err = max(sum_err(kind, got[0], want[0]), sum_err(kind, got[1], want[1]))
        parity["oracle_window"] = {"items_per_rank": int(n_win), "ranks": world, "rel_err": err}
        assert err < parity["tolerance"], "rank %d: window of %d items vs the oracle: %.3e" % (rank, n_win, err)
    shard.pop("data")  # the engine owns the device copy; free the host rows

    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()
    m = time_engine(ctx, eng, thetas, B, K, W)
    clock_info = clocks.stop() if rank == 0 else None
    ms_total, result_dev = m["ms_total"], m["dev"]
    # ---- parity, part 2: host path == device path; every rank holds the same bits
    assert np.allclose(m["host_s"], result_dev[0], rtol=1e-12, atol=0), "host and device paths disagree"
    parity["host_path_equals_device_path"] = bool(np.array_equal(m["host_s"], result_dev[0]) and
                                                  np.array_equal(m["host_q"], result_dev[1]))
    parity["ranks_bit_identical"] = ctx.identical_across_ranks(result_dev)
    assert parity["ranks_bit_identical"], "the reduced sums differ between ranks"

    # ---- dominant kernel: the step is that one launch (the exchange happens inside it), so the
    # CUDA-event time per step is its launch duration (an upper bound for N > 1)
    kern_ms = ms_total / K

    # ---- chain level
    ckw = chain_kwargs(kind, prm, shard, B, args.chain_length)
    chain_info = chain_run(ctx, eng, ckw)  # nodes scored whole: one kernel pass per scheduling burst
    if eng.data_size >= 16 and world == 1:
        # predictive mode (the reference's -b / -y 0): a probe batch of 1/16 of the items, then the rest,
        # only for the nodes the tree still believes in
        chain_info["predictive"] = chain_run(ctx, eng, ckw, batch_items=max(1, eng.data_size // 16), update_style=0)

    # ---- the cases where launch + finalize + exchange overhead shows (same engine, a frontier of ONE node)
    extras = {}
    if default_line and not args.no_extras:
        m1 = time_engine(ctx, eng, thetas[:1], 1, K, W)
        assert rel_err(m1["dev"][0], result_dev[0][:1]) < 1e-12, "B = 1 disagrees with node 0 of the frontier"
        ms1 = m1["ms_total"] / K
        extras["scale_extra"] = {
            "mixture1e8_B1": {"frontier_nodes": 1, "ms_per_step": ms1, "evals_per_sec": 1.0 / (ms1 * 1e-3),
                              "e2e_evals_per_sec": K / m1["e2e_s"], "steps": K,
                              "roofline": roofline_record(args.workload, kind, prm, rows_local, 1, args.dtype, ms1,
                                                          ctx.local_rank, world)}}
    eng.close()
    if default_line and not args.no_extras:
        Ke = max(K, 50)
        extras["scale_extra"]["blasso_B8_rowshard"] = lasso_extra(ctx, Ke, W, args.reduce)
        extras["scale_extra"]["limiter"] = ("per call: launch + TMA prologue + last-CTA finalize + (N > 1) one NVLink "
                                            "store/flag round trip; see DESIGN.md section 5")
        if world > 1:
            extras["node_split"] = node_split_extra(ctx, Ke, W)
        else:
            cfgs = {}
            for name, b, levels in (("normal1e6", 8, 20000), ("boltzmann", 64, 2000), ("mixture1e6", 8, 20000),
                                    ("blasso", 8, 2000)):
                cfgs[name] = small_config(ctx, name, b, "f64", Ke, W, levels)
            # the PF_F32 lasso: the tcgen05 kernel (3xTF32, operands and accumulators in TMEM) under the driver's clock
            cfgs["blasso_f32_tcgen05"] = small_config(ctx, "blasso", 8, "f32", Ke, W, 0)
            cfgs["blasso_f32_tcgen05_B64"] = small_config(ctx, "blasso", 64, "f32", Ke, W, 0)
            extras["configs"] = cfgs

    if rank == 0:
        B_kernel = -(-B // world) if node_split else B  # nodes one GPU's kernel scores
        evals = B * K / (ms_total * 1e-3)
        line = {
            "metric": "loglik_evals_per_sec", "value": evals, "unit": "evals/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": config,
            "chain_steps_per_sec": chain_info["steps_per_sec"],
            "item_evals_per_sec": evals * prm.get("rows", prm.get("D", 1)),
            "clocks": clock_info,
            "e2e": {"value": B * K / m["e2e_s"], "unit": "evals/s", "h2d_bytes_per_step": int(thetas.nbytes),
                    "d2h_bytes_per_step": int(2 * B * 8), "ms_per_step": m["e2e_s"] / K * 1e3},
            "gpu_launches": int(m["launches"]) * world,
            "roofline": roofline_record(args.workload, kind, prm, rows_local, B_kernel, args.dtype, kern_ms,
                                        ctx.local_rank, world, clock_info.get("sm_mhz") if clock_info else None),
            "checksum": {"sum": [float(x) for x in result_dev[0]], "sum_sq": [float(x) for x in result_dev[1]]},
            "parity": parity,
            "chain": chain_info,
        }
        line.update(extras)
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

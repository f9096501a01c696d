#!/usr/bin/env python
"""torchrun --nproc-per-node N tests/multigpu_check.py   (run by tests/test_gpu_multigpu.py when the box has >= 2 GPUs)

Row-sharded engines against a single engine holding the whole data set on the rank's GPU, for the
mixture and the lasso, with both reductions: the in-kernel peer exchange over NVLink (CUDA IPC handles,
pf_engine_peer_attach) and the NCCL all-reduce (pf_engine_comm_init).  Prints one line per
check; exit code 0 only if every rank agrees to 1e-12 with the single-GPU totals."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import fetching_b200 as fb
from fetching_b200 import distributed as pfd, models

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


# mixture: 2000 items of 1000 rows
data, pos = models.mixture_dataset(2_000_000, 8, 2, seed=7)
thetas = models.mixture_thetas(pos, 16, seed=3)
lo, hi = pfd.shard_rows(len(data), 1000, rank, world)
eng = fb.FrontierEngine.mixture(data[lo:hi], 8, 1000, device=local)
pfd.init_engine_comm(eng)
s, q = eng.eval_frontier(thetas)                      # host path: H2D, kernel, ncclAllReduce, D2H
pfd.init_engine_peers(eng)                            # from here on: reduced inside the frontier kernel
l0 = eng.launch_count
for _ in range(5):
    sp, qp = eng.eval_frontier(thetas)                # host path: H2D, ONE kernel, completion flag
eng.peer_status()
ok &= eng.launch_count - l0 == 5
ok &= rel(sp, s) < 1e-13 and rel(qp, q) < 1e-13
tp = torch.from_numpy(np.stack([sp, qp])).cuda()
tp0 = tp.clone()
dist.broadcast(tp0, src=0)
ok &= bool(torch.equal(tp, tp0))                      # rank-order fold: identical bits on every rank
print("rank %d mixture peer exchange vs NCCL: rel %.2e / %.2e, bit-identical across ranks %s" %
      (rank, rel(sp, s), rel(qp, q), bool(torch.equal(tp, tp0))), flush=True)
s, q = sp, qp
d_th = torch.from_numpy(thetas).cuda()
d_out = torch.zeros(2, 16, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
eng.eval_frontier_device(d_th.data_ptr(), d_out.data_ptr(), 16, st.cuda_stream)   # device path
st.synchronize()
full = fb.FrontierEngine.mixture(data, 8, 1000, device=local)
fs, fq = full.eval_frontier(thetas)
e1, e2 = rel(s, fs), rel(q, fq)
same = np.array_equal(d_out.cpu().numpy()[0], s)
print("rank %d mixture shard rows [%d,%d): rel err sum %.2e sum_sq %.2e device==host %s" % (rank, lo, hi, e1, e2, same), flush=True)
ok &= e1 < 1e-12 and e2 < 1e-12 and same
# every rank must hold the same totals after the all-reduce
t = torch.from_numpy(np.stack([s, q])).cuda()
t0 = t.clone()
dist.broadcast(t0, src=0)
ok &= bool(torch.equal(t, t0))
eng.close(); full.close()

# lasso: 40 items of 500 rows, p = 96
X, y, beta = models.blasso_dataset(20000, 96, seed=1)
th = models.blasso_thetas(beta, 9, seed=1)
lo, hi = pfd.shard_rows(20000, 500, rank, world)
eng = fb.FrontierEngine.blasso(X[lo:hi], y[lo:hi], item_rows=500, device=local)
pfd.init_engine_peers(eng)                            # peer exchange only: no NCCL communicator at all
s, q = eng.eval_frontier(th)
eng.peer_status()
full = fb.FrontierEngine.blasso(X, y, item_rows=500, device=local)
fs, fq = full.eval_frontier(th)
e1, e2 = rel(s, fs), rel(q, fq)
print("rank %d lasso   shard rows [%d,%d): rel err sum %.2e sum_sq %.2e" % (rank, lo, hi, e1, e2), flush=True)
ok &= e1 < 1e-12 and e2 < 1e-12
eng.close(); full.close()

# node split (small data set replicated on every GPU; nodes divided over the ranks; all-gather)
data, pos = models.mixture_dataset(100_000, 8, 2, seed=9)
for B in (64, 13, 3):
    thetas = models.mixture_thetas(pos, B, seed=B)
    eng = fb.FrontierEngine.mixture(data, 8, 1000, device=local)
    single = eng.eval_frontier(thetas)
    pfd.init_engine_comm(eng)
    eng.set_option(fb.OPT_NODE_SPLIT, 1)
    s, q = eng.eval_frontier(thetas)                  # NCCL all-gather behind the slice kernel
    e1, e2 = rel(s, single[0]), rel(q, single[1])
    pfd.init_engine_peers(eng)                        # in-kernel gather over peer memory where it applies
    l0 = eng.launch_count
    sp, qp = eng.eval_frontier(thetas)
    fused = eng.launch_count - l0 == 1 and world <= B
    ok &= rel(sp, s) < 1e-13 and rel(qp, q) < 1e-13
    if rank == 0:
        print("node split B=%d over %d ranks: rel err sum %.2e sum_sq %.2e; peer gather vs NCCL gather %.1e (one kernel: %s)" %
              (B, world, e1, e2, rel(sp, s), fused), flush=True)
    ok &= e1 < 1e-12 and e2 < 1e-12
    t = torch.from_numpy(np.stack([s, q])).cuda()
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    ok &= bool(torch.equal(t, t0))
    eng.close()

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU CHECK", "OK" if int(flag.item()) == 1 else "FAILED", "world", world, flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)

"""Host-buffer call time with the resident frontier kernel on / off (PF_OPT_RESIDENT), C loop (pf_eval_frontier_repeat).
PF_TRACE=1 adds the in-kernel time marks of one call per configuration."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, fetching_b200 as fb
from fetching_b200 import models
from fetching_b200._lib import OPT_RESIDENT

def run(name, make, th):
    for res in (1, 0):
        with make() as e:
            e.set_option(OPT_RESIDENT, res)
            if os.environ.get("PF_TRACE"):
                print("---- %s B=%d resident=%d" % (name, th.shape[0], res), file=sys.stderr, flush=True)
                for _ in range(3):
                    e.eval_frontier(th)
                continue
            e.eval_frontier_repeat(th, 50)
            best = min(e.eval_frontier_repeat(th, 2000)[2] / 2000 for _ in range(5))
            st = e.resident_stats
            print("%s B=%d resident=%d: %.2f us per host-buffer call  (stats %s, launches %d)" % (
                name, th.shape[0], res, best * 1e6, st, e.launch_count), flush=True)

nd = models.normal_dataset(1_000_000)
for B in (1, 8, 32, 64):
    run("normal1e6", lambda: fb.FrontierEngine.normal(nd), models.normal_thetas(B, seed=1))
data, pos = models.mixture_dataset(1_000_000, 8, 2, seed=0)
for B in (1, 8, 32, 64):
    run("mixture1e6", lambda: fb.FrontierEngine.mixture(data, 8, 1000), models.mixture_thetas(pos, B, seed=1))

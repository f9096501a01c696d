"""PF_TRACE=1 python scripts/trace_probe.py -- in-kernel time marks (ns) of one host-path call per configuration."""
import os, sys
os.environ.setdefault("PF_TRACE", "1")
sys.path.insert(0, '.')
import numpy as np, fetching_b200 as fb
from fetching_b200 import models
def go(name, eng, th):
    for _ in range(3): eng.eval_frontier(th)
    sys.stderr.write("== %s B=%d\n" % (name, th.shape[0])); sys.stderr.flush()
    eng.eval_frontier(th); eng.eval_frontier(th)
with fb.FrontierEngine.normal(models.normal_dataset(1_000_000)) as e: go("normal1e6", e, models.normal_thetas(8, seed=1))
data, pos = models.mixture_dataset(1_000_000, 8, 2, seed=0)
with fb.FrontierEngine.mixture(data, 8, 1000) as e: go("mixture1e6", e, models.mixture_thetas(pos, 8, seed=1))
with fb.FrontierEngine.mixture(data[:4000], 8, 1000) as e: go("mixture4000", e, models.mixture_thetas(pos, 8, seed=1))
W, X = models.boltzmann_dataset(1024, 64, seed=1)
with fb.FrontierEngine.boltzmann_dense(W) as e: go("boltzmann", e, X)
Xl, y, beta = models.blasso_dataset(100_000, 1000, seed=0)
with fb.FrontierEngine.blasso(Xl, y, item_rows=1000) as e: go("blasso", e, models.blasso_thetas(beta, 8, seed=1))

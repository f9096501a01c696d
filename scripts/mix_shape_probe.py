"""python scripts/mix_shape_probe.py K d rows B -- device time per call of the mixture kernel for another shape (PFETCH_LIB selects a variant)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, fetching_b200 as fb
from fetching_b200 import models
K, d, rows, B = (int(float(x)) for x in sys.argv[1:5])
data, pos = models.mixture_dataset(rows, K, d, seed=0)
with fb.FrontierEngine.mixture(data, K, 1000) as e:
    th = models.mixture_thetas(pos, B, seed=1)
    e.eval_frontier_repeat(th, 3)
    t = min(e.eval_frontier_repeat(th, 10)[2] / 10 for _ in range(3))
    s, q = e.eval_frontier(th)
    print("K=%d d=%d rows=%d B=%d: %.4f ms per host-buffer call; sums[0] %.10g" % (K, d, rows, B, t * 1e3, s[0]))

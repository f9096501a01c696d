import sys; sys.path.insert(0,'.')
import numpy as np, fetching_b200 as fb
from fetching_b200 import models
from fetching_b200._lib import OPT_INLINE_THETA
def run(name, eng, th):
    for inl in (1,0):
        eng.set_option(OPT_INLINE_THETA, inl)
        eng.eval_frontier_repeat(th, 50)
        best=min(eng.eval_frontier_repeat(th, 500)[2]/500 for _ in range(5))
        print("%s B=%d inline=%d: %.2f us per host-buffer call (C loop)"%(name, th.shape[0], inl, best*1e6))
nd = models.normal_dataset(1_000_000)
with fb.FrontierEngine.normal(nd) as e:
    for B in (1,8,64): run("normal1e6", e, models.normal_thetas(B, seed=1))
data,pos = models.mixture_dataset(1_000_000, 8, 2, seed=0)
with fb.FrontierEngine.mixture(data, 8, 1000) as e:
    for B in (1,8,24): run("mixture1e6", e, models.mixture_thetas(pos, B, seed=1))

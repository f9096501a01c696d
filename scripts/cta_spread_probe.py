"""PF_TRACE=11 python scripts/cta_spread_probe.py [rows] [B] -- when every CTA of a mixture call published its record (us):
the first wave of CTAs (one per SM) against the second (its SM mate)."""
import os, sys
os.environ["PF_TRACE"] = "11"
sys.path.insert(0, '.')
import numpy as np, fetching_b200 as fb
from fetching_b200 import models
from fetching_b200._lib import OPT_RESIDENT
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
data, pos = models.mixture_dataset(rows, 8, 2, seed=0)
with fb.FrontierEngine.mixture(data, 8, 1000) as e:
    e.set_option(OPT_RESIDENT, 0)
    th = models.mixture_thetas(pos, B, seed=1)
    for _ in range(3): e.eval_frontier(th)
    sys.stderr.write("== mixture rows=%d B=%d\n" % (rows, B)); sys.stderr.flush()
    e.eval_frontier(th)

"""PF_TRACE=1 PFETCH_LIB=.../libpfetch_tl.so python scripts/tc_timeline.py [B] -- per-step timeline of CTA 0 of the tcgen05 lasso kernel
(library built with -DPF_TC_TIMELINE)."""
import os, sys
os.environ.setdefault("PF_TRACE", "1")
sys.path.insert(0, '.')
import numpy as np, fetching_b200 as fb
from fetching_b200 import models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Xl, y, beta = models.blasso_dataset(100_000, 1000, seed=0)
with fb.FrontierEngine.blasso(Xl, y, item_rows=18000, dtype=fb.F32) as e:
    e.set_option(fb.OPT_MATVEC_TENSOR, 2)
    th = models.blasso_thetas(beta, B, seed=1)
    for _ in range(3): e.eval_frontier(th)
    sys.stderr.write("== blasso f32 tcgen05 B=%d\n" % B); sys.stderr.flush()
    e.eval_frontier(th)

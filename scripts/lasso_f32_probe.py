"""python scripts/lasso_f32_probe.py [B ...] -- PF_F32 lasso C4 (n = 1e5, p = 1000): tcgen05 3xTF32 path vs CUDA-core FP32 path."""
import sys
sys.path.insert(0, ".")
import numpy as np
import bench, fetching_b200 as fb

desc, kind, prm = bench.WORKLOADS["blasso"]
class A: gpus = 1
ctx = bench.Ctx(A)
sh = bench.make_shard(kind, prm, 0, 1)
eng = bench.make_engine(kind, prm, sh, "f32", 0)
floor_ms = (prm["rows"] * prm["p"] * 4 + prm["rows"] * 8) / 6542.1e9 * 1e3
for B in [int(x) for x in sys.argv[1:]] or [8, 16, 32, 64]:
    th = bench.make_thetas(kind, prm, sh, B)
    out = []
    for tensor in (2, 0):
        eng.set_option(fb.OPT_MATVEC_TENSOR, tensor)
        m = bench.time_engine(ctx, eng, th, B, 50, 5)
        out.append((m["ms_total"] / 50, m["e2e_s"] / 50 * 1e6, m["dev"][0].copy()))
    print("lasso f32 B=%2d: tcgen05 %.4f ms/step (%.2fx the HBM floor %.4f), e2e %.1f us | CUDA cores %.4f ms/step | rel diff %.1e"
          % (B, out[0][0], out[0][0] / floor_ms, floor_ms, out[0][1], out[1][0],
             float(np.max(np.abs(out[0][2] - out[1][2]) / np.abs(out[1][2])))))

#!/bin/bash
# Round benchmark sweep (1 GPU): every BASELINE configuration at frontier sizes 1 / 8 / 64, FP64 (+ the FP32 corner).
# usage: scripts/bench_sweep.sh OUT.log   -- "== workload B=.. dtype" header lines followed by bench.py's JSON line
out=${1:-gpurun_out/bench_all.log}
: > "$out"
run() { echo "== $1 B=$2 ${3:-f64}" >> "$out"; timeout 600 python bench.py --workload "$1" --frontier "$2" --dtype "${3:-f64}" --steps 30 --warmup 5 2>/dev/null | tail -1 >> "$out"; }
for w in normal1e6 mixture1e6 blasso; do for b in 1 8 64; do run $w $b; done; done
run blasso 16
run blasso 32
run boltzmann 64
run mixture1e8 1
run mixture1e8 64
run mixture1e8 8 f32
run mixture1e8 8
python - "$out" <<'PY'
import json, sys
print("| workload | B | dtype | ms/step | evals/s (HBM-resident) | e2e evals/s | HBM achieved (frac) | FP64 pipe frac | chain steps/s | CPU port evals/s |")
print("|---|---:|---|---:|---:|---:|---|---:|---:|---|")
hdr = None
for l in open(sys.argv[1]):
    if l.startswith("=="):
        hdr = l.split()
    elif l.startswith("{"):
        d = json.loads(l)
        f = (d["roofline"].get("fp64_pipe") or {}).get("frac")
        cb = d.get("cpu_baseline", {})
        print("| %s | %d | %s | %.4f | %.0f | %.0f | %.0f GB/s (%.3f) | %s | %.0f | %.2f (%s thr) |" % (
            hdr[1], d["config"]["frontier_nodes"], d["dtype"], d["ms_per_step"], d["value"], d["e2e"]["value"],
            d["roofline"]["achieved"], d["roofline"]["frac"], "%.2f" % f if f else "-", d["chain"]["steps_per_sec"],
            cb.get("value", float("nan")), cb.get("cores", "?")))
PY

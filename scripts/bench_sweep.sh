#!/bin/bash
# Round benchmark sweep (1 GPU): every BASELINE configuration at several frontier sizes, FP64 (+ the FP32 corners).
# usage: scripts/bench_sweep.sh OUT.log   -- "== workload B=.. dtype" header lines followed by bench.py's JSON line
out=${1:-gpurun_out/bench_all.log}
: > "$out"
run() { echo "== $1 B=$2 ${3:-f64}" >> "$out"; timeout 600 python bench.py --no-extras --no-cpu --workload "$1" --frontier "$2" --dtype "${3:-f64}" --steps 30 --warmup 5 2>/dev/null | tail -1 >> "$out"; }
for w in normal1e6 mixture1e6 blasso; do for b in 1 8 64; do run $w $b; done; done
run blasso 16
run blasso 32
run blasso 1 f32
run blasso 8 f32
run blasso 32 f32
run blasso 64 f32
run boltzmann 64
run mixture1e8 1
run mixture1e8 16
run mixture1e8 32
run mixture1e8 64
run mixture1e8 8 f32
run mixture1e8 8
python - "$out" <<'PY'
import json, sys
print("| workload | B | dtype | ms/step | evals/s (HBM-resident) | e2e evals/s | binding unit: achieved / peak (frac) | HBM GB/s (frac) | chain steps/s | min margin / \|sum\| |")
print("|---|---:|---|---:|---:|---:|---|---|---:|---:|")
hdr = None
for l in open(sys.argv[1]):
    if l.startswith("=="):
        hdr = l.split()
    elif l.startswith("{"):
        d = json.loads(l)
        r = d["roofline"]
        print("| %s | %d | %s | %.4f | %.0f | %.0f | %s: %.3g / %.3g (%.2f) | %.0f (%.3f) | %.0f | %.1e |" % (
            hdr[1], d["config"]["frontier_nodes"], d["dtype"], d["ms_per_step"], d["value"], d["e2e"]["value"],
            r["bound"], r["achieved"], r["peak"], r["frac"], r["hbm"]["achieved"], r["hbm"]["frac"],
            d["chain"]["steps_per_sec"], d["chain"]["min_margin_rel"]))
PY

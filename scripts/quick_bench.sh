#!/bin/bash
# usage: scripts/quick_bench.sh "workload:B[:dtype]" ...   -- one line per configuration (device ms, HBM GB/s, e2e)
for spec in "$@"; do
  IFS=: read -r w b dt <<< "$spec"
  timeout 300 python bench.py --workload "$w" --frontier "$b" --dtype "${dt:-f64}" --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
f=d['roofline'].get('fp64_pipe') or {}
print('%-11s B=%-3d %s ms=%.4f hbm=%.0f GB/s (%.3f) fp64=%.2f e2e_ms=%.4f chain=%.0f steps/s' % ('$w', d['config']['frontier_nodes'], d['dtype'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], f.get('frac', float('nan')), d['e2e']['ms_per_step'], d['chain']['steps_per_sec']))"
done

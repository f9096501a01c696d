#!/bin/bash
# usage (GPU box): scripts/variant_bench.sh [rows] [B] -- device-timed ms/step of every fetching_b200/lib/libpfetch*.so
# (experimental variants built with `python -m fetching_b200.build --out=libpfetch_<name>.so -D...`), mixture K=8 d=2
rows=${1:-32000000}; B=${2:-8}
for lib in fetching_b200/lib/libpfetch*.so; do
  PFETCH_LIB=$PWD/$lib python bench.py --no-extras --no-cpu --rows $rows --frontier $B --steps 10 --warmup 3 --chain-length 20 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-28s B=%d rows=%d  %.4f ms/step  (x%.2f to 1e8 rows: %.3f ms)  oracle err %.1e' % ('$(basename $lib)', d['config']['frontier_nodes'], d['config']['rows_total'], d['ms_per_step'], 1e8/d['config']['rows_total'], d['ms_per_step']*1e8/d['config']['rows_total'], d['parity']['oracle_window']['rel_err']))"
done

#!/bin/bash
# ncu evidence for the round (run on the GPU box under gpurun; every bench command first exits 0 without ncu).
# Outputs under gpurun_out/: launches_*.csv (launch lists) and prof_*.ncu-rep (one full capture per dominant kernel).
# gpurun returns at most 64 MiB and a pointwise capture is ~17-20 MB: `scripts/profile_round.sh small` takes the two
# small-model pointwise captures in a second call.
set -u
o=gpurun_out
if [ "${1:-main}" = "small" ]; then
  for w in normal1e6 mixture1e6; do
    python bench.py --workload $w --frontier 8 --steps 5 --warmup 3 > /dev/null 2>&1 || exit 1
    ncu --set full --clock-control none -k regex:pw_frontier -s 4 -c 1 -f -o $o/prof_${w}_B8 \
        python bench.py --workload $w --frontier 8 --steps 5 --warmup 3 > /dev/null 2>&1
  done
else
  python bench.py --steps 5 --warmup 3 > $o/bench_check.json 2>/dev/null || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_mixture1e8.csv \
      python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:pw_frontier -s 4 -c 1 -f -o $o/prof_mixture1e8_B8 \
      python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
  python bench.py --workload blasso --frontier 8 --steps 5 --warmup 3 > /dev/null 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $o/launches_blasso.csv \
      python bench.py --workload blasso --frontier 8 --steps 5 --warmup 3 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:mv_frontier -s 4 -c 1 -f -o $o/prof_blasso_B8 \
      python bench.py --workload blasso --frontier 8 --steps 5 --warmup 3 > /dev/null 2>&1
  ncu --set full --clock-control none -k regex:mv_frontier -s 4 -c 1 -f -o $o/prof_blasso_B64 \
      python bench.py --workload blasso --frontier 64 --steps 5 --warmup 3 > /dev/null 2>&1
fi
ls -la $o/*.ncu-rep $o/launches_*.csv 2>/dev/null
du -sh $o

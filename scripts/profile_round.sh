#!/bin/bash
# ncu evidence for the round (run on the GPU box under gpurun; every bench command first exits 0 without ncu).
# usage: scripts/profile_round.sh main|small|lasso   -- gpurun returns at most 64 MiB per call, hence the three parts.
# Outputs under gpurun_out/: r2_launches_*.csv (launch lists of the default command) and r2_prof_*.ncu-rep.
set -u
o=gpurun_out
B="python bench.py --no-extras --no-cpu --steps 5 --warmup 3 --chain-length 20"
cap() { # name kernel-regex bench-args...
  local name=$1 k=$2; shift 2
  $B "$@" > /dev/null 2>&1 || { echo "bench failed: $*"; return 1; }
  ncu --set full --clock-control none -k regex:$k -s 4 -c 1 -f -o $o/r2_prof_$name $B "$@" > /dev/null 2>&1
}
case "${1:-main}" in
main)
  python bench.py --steps 5 --warmup 3 > $o/r2_bench_check.json 2>/dev/null || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/r2_launches_default.csv \
      python bench.py --steps 5 --warmup 3 > /dev/null 2>&1
  $B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pw_frontier -s 4 -c 1 -f \
      -o $o/r2_prof_mixture1e8_B8 $B > /dev/null 2>&1
  cap mixture1e8_B8_f32 pw_frontier --dtype f32
  ;;
small)
  cap mixture1e6_B8 pw_frontier --workload mixture1e6
  cap normal1e6_B8 pw_frontier --workload normal1e6
  cap boltzmann_B64 mv_frontier --workload boltzmann
  ;;
lasso)
  cap blasso_B8 mv_frontier --workload blasso
  cap blasso_B64 mv_frontier --workload blasso --frontier 64
  cap blasso_f32_B64_tcgen05 mv_tc32 --workload blasso --frontier 64 --dtype f32
  cap blasso_f32_B8_tcgen05 mv_tc32 --workload blasso --frontier 8 --dtype f32
  ;;
esac
ls -la $o/r2_prof_*.ncu-rep $o/r2_launches_*.csv 2>/dev/null
du -sh $o

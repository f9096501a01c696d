#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into small committed summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_x.ncu-rep profiles/r1_x_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = defaultdict(lambda: [0, 0.0, "", ""])
    order = []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])[:110]
        if name not in agg:
            order.append(name)
        a = agg[name]
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
        a[2], a[3] = r[gi], r[bi]
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), from `%s`\n\n" % src)
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share | grid | block |\n|---|---:|---:|---:|---|---|\n")
        for name in sorted(order, key=lambda n: -agg[n][1]):
            a = agg[name]
            f.write("| `%s` | %d | %.1f | %.1f%% | %s | %s |\n" % (name, a[0], a[1] / 1e3, 100 * a[1] / total, a[2], a[3]))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary, from `%s`\n\n" % src)
        for r in rows[2:]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % r[hdr.index("Kernel Name")][:120])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio$", h)
                if m:
                    try:
                        stalls.append((float(r[i]), m.group(1)))
                    except ValueError:
                        pass
            f.write("\nTop warp stall reasons (warps stalled per issue-active cycle): " +
                    ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]) + "\n\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

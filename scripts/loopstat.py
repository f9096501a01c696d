#!/usr/bin/env python3
"""usage: scripts/loopstat.py [lib.so] [kernel-name regex]

Instruction mix of the hot row loops of a frontier kernel, from the SASS of the built library: every innermost loop
(backward branch without a call or a barrier inside) with at least 40 FP64 instructions is listed with its FP64 /
other instruction counts and the dispatch-cycle model of DESIGN.md section 4 (an FP64 instruction holds the
dispatch port for 2 cycles, any other for 1).  Default kernel: the FP64 mixture kernel K = 8, d = 2, one node per lane
(it contains the factored loop and the clamped-distance loop, each in a 4-row and a 2-row version)."""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1] if len(sys.argv) > 1 else "fetching_b200/lib/libpfetch.so"
pat = sys.argv[2] if len(sys.argv) > 2 else r"pw_frontier_kernel2INS_12MixtureModelILi8ELi2EdLb0EEEdLi1E"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if not re.search(pat, name):
        continue
    ins = []
    for l in f.split("\n"):
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print("kernel", name, "-", len(ins), "instructions")
    census = Counter(re.sub(r"^@!?U?P\d\s+", "", x).split()[0].split(".")[0] for _, x in ins)
    print("  whole kernel:", {k: census[k] for k in ("UBLKCP", "UTMALDG", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "LDS", "MUFU", "BAR") if census[k]})
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:[A-Z0-9!]+,\s*)?0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            if any(re.search(r"\b(CALL|BAR|SYNCS)\b", x) for _, x in body):
                continue
            nd = sum(1 for _, x in body if re.search(r"\bD(FMA|ADD|MUL)\b", x))
            if nd >= 40:
                loops.append((tgt, a, body, nd))
    # innermost only
    loops = [l for l in loops if not any(o is not l and o[0] >= l[0] and o[1] <= l[1] for o in loops)]
    for tgt, a, body, nd in loops:
        c = Counter()
        for _, x in body:
            c[re.sub(r"^@!?U?P\d\s+", "", x).split()[0].split(".")[0]] += 1
        n = len(body)
        lds_tab = c["LDS"]
        rows = max(1, round(sum(1 for _, x in body if re.search(r"\bI2F|LDS\.128", x)) / 1))  # informational only
        print("  loop 0x%04x-0x%04x: %d instructions, FP64 %d, other %d, model cycles %d (FP64 share of dispatch %.2f)"
              % (tgt, a, n, nd, n - nd, 2 * nd + n - nd, 2 * nd / (2 * nd + n - nd)))
        print("    ", dict(c.most_common()))

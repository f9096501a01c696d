"""Count FP64 instructions by distinct register-pair sources (register-file read cycles) in a kernel's SASS.

A DFMA/DADD/DMUL whose three sources are distinct 64-bit register pairs needs 3 register-file cycles instead of the
pipe's 2 (measured: scripts/fp64_probe.py, 42.8 vs 60 DFMA/clk/SM); uniform registers, constants, immediates, RZ and
operands held in the reuse cache (`.reuse` on the same slot of the previous instruction) are free.

usage: python scripts/rf_pressure.py <obj> <kernel-name-substring> [--loop]   (--loop: only the largest backward-branch body)
"""
import re, subprocess, sys

def main():
    obj, pat = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    hits = [f for f in funcs if pat in f.split("\n", 1)[0]]
    for f in hits:
        name = f.split("\n", 1)[0]
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        if "--loop" in sys.argv:
            best = None
            for addr, t in ins:
                m = re.search(r"BRA\s+(?:U,)?\s*`?\(?0x([0-9a-f]+)", t)
                if m and int(m.group(1), 16) < addr:
                    span = (int(m.group(1), 16), addr)
                    n64 = sum(1 for a, tt in ins if span[0] <= a <= span[1] and re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL)", tt))
                    if best is None or n64 > best[0]:
                        best = (n64, span)
            if best:
                ins = [(a, t) for a, t in ins if best[1][0] <= a <= best[1][1]]
        hist = {}
        prev_reuse = {}
        total = 0
        for addr, t in ins:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t2.split()[0]
            args = [a.strip() for a in t2[len(op):].split(",")]
            srcs = args[1:]
            cur_reuse = {}
            if re.match(r"D(FMA|ADD|MUL)", op):
                regs = set()
                for slot, s in enumerate(srcs):
                    m = re.match(r"[-|]*\s*R(\d+)(\.reuse)?", s)
                    if not m:
                        continue
                    r = int(m.group(1))
                    if m.group(2):
                        cur_reuse[slot] = r
                    if prev_reuse.get(slot) == r:
                        continue
                    regs.add(r)
                hist[len(regs)] = hist.get(len(regs), 0) + 1
                total += 1
            else:
                for slot, s in enumerate(srcs):
                    m = re.match(r"[-|]*\s*R(\d+)\.reuse", s)
                    if m:
                        cur_reuse[slot] = int(m.group(1))
            prev_reuse = cur_reuse
        cyc = sum(max(2, k) * v for k, v in hist.items())
        print(f"{name[:110]}\n  instrs {len(ins)}  fp64 {total}  by distinct reg sources {dict(sorted(hist.items()))}  "
              f"rf-limited cycles {cyc} vs pipe {2 * total}  (cap {2 * total / max(cyc, 1):.2f})")

main()

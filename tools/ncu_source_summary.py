#!/usr/bin/env python
"""Per-region summary of an `ncu --page source --csv` dump of one kernel: contiguous SASS regions by execution count with
their stall-sample breakdown, converted to cycles per execution (samples of the region / samples per issue).
  ncu -i rep.ncu-rep --page source --csv > src.csv;  python tools/ncu_source_summary.py src.csv [min_exec] [listing.txt]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
R = ["stall_selected", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_branch_resolving", "stall_no_inst",
     "stall_sleep", "stall_membar", "stall_math", "stall_dispatch", "stall_not_selected", "stall_mio", "stall_lg"]


def num(r, k):
    try:
        return int(r[ix[k]] or 0)
    except (ValueError, KeyError):
        return 0


regs, cur = [], None
for n, r in enumerate(data):
    e = num(r, "Instructions Executed")
    if cur and cur["e"] == e:
        cur["b"] = n
    else:
        cur = {"e": e, "a": n, "b": n}
        regs.append(cur)
for g in regs:
    g["n"] = g["b"] - g["a"] + 1
    g["s"] = {k: sum(num(r, k) for r in data[g["a"]:g["b"] + 1]) for k in R}
    g["tot"] = sum(num(r, "# Samples") for r in data[g["a"]:g["b"] + 1])
sel = sum(g["s"]["stall_selected"] for g in regs if g["e"] >= min_exec)
ins = sum(g["n"] * g["e"] for g in regs if g["e"] >= min_exec)
spc = sel / max(ins, 1)  # samples per issued warp instruction = samples per cycle of one warp
print(f"samples per issue-cycle: {spc:.3e}")
for g in regs:
    if g["e"] >= min_exec and g["n"] >= 6:
        cyc = g["tot"] / spc / g["e"]
        top = ", ".join(f"{k[6:]}={v / spc / g['e']:.0f}" for k, v in sorted(g["s"].items(), key=lambda kv: -kv[1])[:5] if v)
        print(f"rows {g['a']:5d}-{g['b']:5d} exec={g['e']:8d} instr={g['n']:4d} cycles/exec={cyc:7.0f}  [{top}]")
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for n, r in enumerate(data):
            e = num(r, "Instructions Executed")
            if e >= min_exec:
                f.write(f"{n:5d} e={e:7d} {num(r, '# Samples') / spc / e:7.1f} w={num(r, 'stall_wait') / spc / e:5.1f} "
                        f"l={num(r, 'stall_long_sb') / spc / e:5.1f} s={num(r, 'stall_short_sb') / spc / e:5.1f} {r[ix['Source']].strip()[:90]}\n")

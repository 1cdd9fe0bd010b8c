#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode of one kernel in an ncu report (source page).
usage: tools/ncu_opcodes.py <report.ncu-rep> [top N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
agg = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    op = r[ci["Source"]].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    o = o.rstrip(";")
    base = o.split(".")[0]
    ins = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    agg[base] += ins; smp[base] += s; tot += ins; tots += s
print("total warp-instr %d  samples %d" % (tot, tots))
for o, n in agg.most_common(top):
    print("%-12s %12d %6.2f%%   samples %6.2f%%" % (o, n, 100.0 * n / tot, 100.0 * smp[o] / max(tots, 1)))

#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS listing by CUDA source line, using nvdisasm -g line markers.

usage: tools/ncu_by_line.py <report.ncu-rep> <object-or-cubin> <kernel mangled-name substring> [top N]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    obj = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout.splitlines()
# address -> (file, line) inside the wanted function
addr2line = {}
infunc = False
cur = ("?", 0)
for ln in dis:
    if ln.startswith(".text."):
        infunc = kname in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None:
        base = a
    off = a - base
    key = addr2line.get(off, (("?", 0), ""))[0]
    ins = int(r[ci["Instructions Executed"]] or 0)
    smp = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    thr = int(r[ci["Thread Instructions Executed"]] or 0)
    for k, v in enumerate((ins, smp, thr)):
        agg[key][k] += v
        tot[k] += v
print("total warp-instr %d  samples %d  thread-instr %d" % tuple(tot))
print("%-28s %12s %7s %9s %7s" % ("file:line", "warp-instr", "%", "samples", "%"))
for key, (ins, smp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-28s %12d %6.2f%% %9d %6.2f%%" % ("%s:%d" % key, ins, 100.0 * ins / max(tot[0], 1), smp,
                                               100.0 * smp / max(tot[1], 1)))

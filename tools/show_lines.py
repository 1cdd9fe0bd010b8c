#!/usr/bin/env python
"""Print a tools/ncu_by_line.py listing with the source text of every line next to it.
usage: tools/show_lines.py profiles/<tag>_<kernel>_by_line.txt [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cache = {}
for ln in open(sys.argv[1]).read().splitlines()[2:2 + n]:
    key = ln.split()[0]
    f, l = key.rsplit(":", 1)
    path = os.path.join(ROOT, "relagn_b200", "csrc", f)
    txt = ""
    if os.path.exists(path):
        src = cache.setdefault(path, open(path).read().splitlines())
        if 0 < int(l) <= len(src):
            txt = src[int(l) - 1].strip()[:110]
    print("%-24s %s  | %s" % (key, " ".join("%9s" % x for x in ln.split()[1:]), txt))

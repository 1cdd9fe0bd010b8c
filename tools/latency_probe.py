"""Single-object latency (BASELINE config 1: one default relagn set): host call to host result, and the kernels'
share with per-kernel CUDA-event brackets."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relagn_b200 import _capi, tables as T  # noqa: E402

ctx = _capi.Context(0)
ctx.set_grid(np.geomspace(1e-4, 1e3, 1001))
T.load_default(ctx)
p = np.array([[1e8, 100, -1, 0, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0]])
res = {}
for rel in (True, False):
    for comps in (False, True):
        for _ in range(20):
            ctx.eval_batch_host(p, rel=rel, components=comps)
        t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            ctx.eval_batch_host(p, rel=rel, components=comps)
        res["host_call_ms_rel%d_comps%d" % (rel, comps)] = (time.perf_counter() - t0) / n * 1e3
ctx.set_profile(True)
ctx.reset_stats()
for _ in range(50):
    ctx.eval_batch_host(p, rel=True)
s = ctx.stats()
res["kernel_ms_per_call"] = {k: s[k] / 50 for k in ("ms_setup", "ms_nthcomp", "ms_hist", "ms_spectrum")}
print(json.dumps(res))

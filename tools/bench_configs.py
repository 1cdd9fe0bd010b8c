"""Timings of the other BASELINE.json configurations (device-resident, CUDA events); prints one JSON object.
C1: single default relagn set (latency), C3: relqso 100 x 100 mass-mdot grid, C5: hi-res 5000 bins x 2034 annuli."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relagn_b200.batch import BatchEvaluator  # noqa: E402
from tests.util import c3_params, RELAGN_DEF  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
dev = torch.device("cuda", 0)
g1000 = np.geomspace(1e-4, 1e3, 1001)
ev = BatchEvaluator(g1000, model="relagn", device=0)
p1 = torch.tensor([RELAGN_DEF], dtype=torch.float64, device=dev)
res["C1_relagn_default_P1_rel_ms"] = timed(lambda: ev.evaluate_device(p1, rel=True))
res["C1_relagn_default_P1_nonrel_ms"] = timed(lambda: ev.evaluate_device(p1, rel=False))
evq = BatchEvaluator(g1000, model="relqso", device=0)
p3 = torch.from_numpy(c3_params(100)).to(dev)
ms = timed(lambda: evq.evaluate_device(p3, rel=True), reps=3)
res["C3_relqso_grid_1e4_rel_ms"] = ms
res["C3_relqso_grid_1e4_spectra_per_s"] = 1e4 / (ms * 1e-3)
evq.ctx.set_profile(True); evq.ctx.reset_stats(); evq.evaluate_device(p3, rel=True); torch.cuda.synchronize()
s = evq.ctx.stats(); evq.ctx.set_profile(False)
res["C3_kernels_ms"] = {k: s[k] for k in ("ms_setup", "ms_nthcomp", "ms_hist", "ms_spectrum")}
g5000 = np.geomspace(1e-4, 1e3, 5001)
ev5 = BatchEvaluator(g5000, model="relagn", device=0, dr_dex=700.0)
p5 = torch.tensor([[1e8, 100, -1, 0.998, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, 3, 1, 10, 0]] * 8, dtype=torch.float64, device=dev)
ms = timed(lambda: ev5.evaluate_device(p5, rel=True), reps=3, warm=1)
res["C5_hires_5000x2034_P8_ms"] = ms
res["C5_hires_spectra_per_s"] = 8 / (ms * 1e-3)
ev5.ctx.set_profile(True); ev5.ctx.reset_stats(); ev5.evaluate_device(p5, rel=True); torch.cuda.synchronize()
s = ev5.ctx.stats(); ev5.ctx.set_profile(False)
res["C5_kernels_ms"] = {k: s[k] for k in ("ms_setup", "ms_nthcomp", "ms_hist", "ms_spectrum")}
print(json.dumps(res))

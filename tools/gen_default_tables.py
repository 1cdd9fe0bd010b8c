"""Ray-trace the default Kerr transfer tables on the device (rg_tables_generate) and store them in the package
cache relagn_b200/_tables/ (and a copy under gpurun_out/ so the file travels back from a gpurun box)."""
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relagn_b200 import _capi, tables as T  # noqa: E402

ctx = _capi.Context(0)
t0 = time.time()
bad = T.load_default(ctx)
print("default tables: %d nodes, %d unresolved points, %.1f s" % (T.DEFAULT["spins"].size * T.DEFAULT["thetas"].size,
                                                                   bad, time.time() - t0))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for f in os.listdir(T.CACHE_DIR):
    shutil.copy(os.path.join(T.CACHE_DIR, f), os.path.join(ROOT, "gpurun_out", f))
    print("cached", f)

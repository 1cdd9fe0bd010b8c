"""Small evaluations of every path for compute-sanitizer (memcheck / racecheck / initcheck): the large-batch path with
both interpolation variants, the small-batch path, components, relqso, the kyconv seam."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from relagn_b200.batch import BatchEvaluator

P = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ev = BatchEvaluator(bench.ebins(), model="relagn", device=0)
ev.ensure_tables()
p = bench.c4_params(P, 5)
a, _ = ev.evaluate(p)
a = a.copy()
os.environ["RG_INTERP_GLOBAL"] = "1"
b, _ = ev.evaluate(p)
del os.environ["RG_INTERP_GLOBAL"]
print("large path:", a.shape, "finite", bool(np.isfinite(a).all()), "global-table variant equal:", bool(np.array_equal(a, b)))
c, _ = ev.evaluate(p[:3])
print("small path equal to the large one:", bool(np.array_equal(c, a[:3])))
d, _ = ev.evaluate(p[:200], components=True)
print("components:", d.shape, bool(np.isfinite(d).all()))

"""The reference's own Python CPU path, timed on the host cores.

What runs is the UNMODIFIED scotthgn/RELAGN `relagn.py` + `pyNTHCOMP.py`
(/root/reference/src/python_version/), copied verbatim by `install_from()` -- called from __graft_entry__.build() in
the build container, where /root/reference exists -- into the git-ignored directory baseline/_ref/, which travels to
the GPU box with the snapshot (it is not in .gpurunignore) and never enters the repository history.  The files import
`xspec` (HEASOFT) and `astropy`, neither of which is installed or installable offline; tests/golden/ref_stubs.py
provides in-memory stand-ins (CODATA constants, the unit conversions the file performs, and a kyconv seam that
raises).  KYCONV unavailable offline; baseline = non-relativistic AGNSED-equivalent path: the timed call is
`relagn(**pars).new_ear(grid); get_totSED(rel=False)` (relagn.py:1552-1577 -> do_nonrel*Spec :1030-1059, :1127-1156,
:1387-1401), on one core (the reference is single-threaded) and as one process per host core.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline_reference block; never imported by
relagn_b200/.
"""
import contextlib
import io
import os
import shutil
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("relagn.py", "pyNTHCOMP.py")
KEYS = ["M", "dist", "log_mdot", "a", "cos_inc", "kTe_hot", "kTe_warm", "gamma_hot", "gamma_warm", "r_hot", "r_warm",
        "log_rout", "fcol", "h_max", "z"]


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def install_from(reference_root="/root/reference"):
    """Copy the two reference files (unmodified) into baseline/_ref/.  Returns True when they are in place."""
    src = os.path.join(reference_root, "src", "python_version")
    if all(os.path.exists(os.path.join(src, f)) for f in FILES):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    return available()


def _load():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_stubs
    ref_stubs.install(None)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import relagn as ref
    return ref


def _run(args):
    """Worker: evaluate get_totSED(rel=False) for every row of params; returns (n, seconds, checksum)."""
    params, grid = args
    ref = _load()
    chk = 0.0
    t0 = time.perf_counter()
    for p in params:
        with contextlib.redirect_stdout(io.StringIO()):  # the class prints WARNING lines when it clamps radii
            obj = ref.relagn(**dict(zip(KEYS, [float(v) for v in p])))
            obj.new_ear(np.array(grid))
            spec = obj.get_totSED(rel=False)
        chk += float(spec[len(spec) // 2])
    return len(params), time.perf_counter() - t0, chk


def _spawn(params, grid, tmpdir, tag):
    """One worker process (its own interpreter: the stubs go into sys.modules there, not here)."""
    import subprocess
    path = os.path.join(tmpdir, "job_%s.npz" % tag)
    np.savez(path, params=params, grid=grid)
    return subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", path], stdout=subprocess.PIPE,
                            stderr=subprocess.DEVNULL, text=True, cwd=ROOT)


def _collect(proc, timeout):
    import json
    out, _ = proc.communicate(timeout=timeout)
    return json.loads(out.strip().splitlines()[-1])


def time_nonrel(params, grid, n_single=48, n_per_proc=24, procs=None, timeout=600):
    """Spectra/s of the reference's non-relativistic path on 1 core and on `procs` cores (one process each).
    params: [P, 15] C4 sets; bounded samples: n_single sets on one core, n_per_proc sets per process."""
    import tempfile
    procs = procs or os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        r1 = _collect(_spawn(params[:n_single], grid, tmp, "single"), timeout)
        t0 = time.perf_counter()
        ps = [_spawn(params[(k * n_per_proc) % len(params):][:n_per_proc], grid, tmp, "w%d" % k) for k in range(procs)]
        res = [_collect(p, timeout) for p in ps]
        wall = time.perf_counter() - t0
    nN = sum(r["n"] for r in res)
    tN = max(r["seconds"] for r in res)  # the slowest worker's loop (imports and process start-up excluded)
    return {"single_core": {"value": r1["n"] / r1["seconds"], "sets": r1["n"], "seconds": r1["seconds"]},
            "all_cores": {"value": nN / tN, "sets": nN, "seconds": tN, "processes": procs, "wall_with_startup": wall}}


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--worker":
        import json
        job = np.load(sys.argv[2])
        n, dt, chk = _run((job["params"], job["grid"]))
        print(json.dumps({"n": int(n), "seconds": dt, "checksum": chk}))
    else:
        print(install_from())

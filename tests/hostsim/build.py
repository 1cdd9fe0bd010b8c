"""Build the TEST-ONLY host simulation of the kernels (see cuda_shim.h): the product's .cu sources compiled
with g++ against the CUDA shim -> tests/hostsim/_build/librelagn_hostsim.so.  Never loaded by relagn_b200/."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "relagn_b200", "csrc")
OUT = os.path.join(HERE, "_build", "librelagn_hostsim.so")


def build(force=False, opt="-O1"):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu"))) + [os.path.join(HERE, "cuda_shim.cpp")]
    deps = srcs + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "cuda_shim.h"), os.path.join(ROOT, "include", "relagn_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(HERE, "_build", os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["g++", "-std=c++17", opt, "-g", "-fPIC", "-ffp-contract=off", "-DRG_HOSTSIM", "-I", HERE, "-I", CSRC,
               "-x", "c++", "-c", s, "-o", o, "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
               "-Wno-unused-variable"]
        procs.append((s, subprocess.Popen(cmd)))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError("hostsim compile failed: " + s)
    subprocess.check_call(["g++", "-shared", "-o", OUT] + objs + ["-lm"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))

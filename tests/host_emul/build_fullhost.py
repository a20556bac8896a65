"""Builds tests/host_emul/_build/libphb200_hostemul.so: the whole product library compiled for the host
(TEST INFRASTRUCTURE, see fullhost/cuda_runtime.h).  PHB200_TEST_HOST_EMUL=1 makes tests/conftest.py point the
ctypes binding at it, so `pytest -m gpu` can be exercised where there is no GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "fullhost"))
from cu2cpp import convert  # noqa: E402

SRC = ["api", "assembly", "solver", "sparse", "comm", "timestep", "mfg", "incomp"]
OUT = os.path.join(HERE, "_build", "libphb200_hostemul.so")


def build(force=False):
    csrc = os.path.join(ROOT, "phasta_b200", "csrc")
    gen = os.path.join(HERE, "_build", "fullhost_src")
    os.makedirs(gen, exist_ok=True)
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".h", ".cuh"))]
    deps += [os.path.join(HERE, "fullhost", f) for f in os.listdir(os.path.join(HERE, "fullhost"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".h", ".cuh")):
            text = convert(open(os.path.join(csrc, f)).read())
            text = text.replace('"../../include/phb200.h"', '"%s"' % os.path.join(ROOT, "include", "phb200.h"))
            open(os.path.join(gen, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w").write(text)
    objs = []
    procs = []
    for s in SRC + ["shim_runtime"]:
        src = os.path.join(gen, s + ".cpp") if s != "shim_runtime" else os.path.join(HERE, "fullhost", "shim_runtime.cpp")
        obj = os.path.join(gen, s + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", "-O1", "-g", "-fPIC", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
                                       "-DPHB_HOST_FULL=1", "-D__CUDACC__=1", "-I", os.path.join(HERE, "fullhost"), "-I", gen,
                                       "-c", src, "-o", obj]))
    if any(p.wait() for p in procs):
        raise RuntimeError("host-emulation build failed")
    subprocess.check_call(["g++", "-shared", "-o", OUT] + objs + ["-ldl", "-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force=True))

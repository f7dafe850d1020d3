"""Build libhelfemqc_b200.so in-tree with nvcc for sm_100a.

    python -m helfem_b200.build [--force]

The library is the product: C++ host setup + CUDA kernels + the C ABI declared
in include/helfem_b200.h.  It travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhelfemqc_b200.so")
SOURCES = ["fem.cpp", "special.cpp", "atomic_setup.cpp", "diatomic_setup.cpp", "grid_setup.cpp", "comm.cpp", "engine.cu", "grid.cu", "solver.cu", "tei_device.cu", "capi.cpp"]
HEADERS = ["fem.h", "special.h", "tables.h", "engine.h", "grid.h", "comm.h", "tei_device.h", "kernels.cuh", "xc_builtin.cuh", "../../include/helfem_b200.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-fopenmp,-O3", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu" if src.endswith(".cu") else "c++"]
        if not src.endswith(".cu"):
            cmd = [NVCC] + FLAGS + ["-x", "c++"]
        cmd += ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out.strip():
            print(out)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lcudart", "-lgomp", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Compile the CUDA sources under csrc/ into libhbp_b200.so (sm_100a only).

    python -m human_body_proportion_estimation_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is kept in-tree next to
the sources (git-ignored, shipped to the GPU box by gpurun).  cudart is linked
statically and the driver API is resolved at run time, so the library also
loads on a box without a GPU (symbol checks in the CPU test tier).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libhbp_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fvisibility=hidden", "-DHBP_BUILD"]


def nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p) and f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def up_to_date():
    stamp = LIB + ".stamp"
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _digest()


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [cc] + ARCH + FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    cmd = [cc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(LIB + ".stamp", "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

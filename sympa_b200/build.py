"""Builds libsympa_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library has a plain C ABI and is loaded with ctypes).

    python -m sympa_b200.build [--force]

One translation unit per matrix size n (pair_kernels_n.cu -DSYMPA_TU_N=n) plus the ABI file,
compiled in parallel; objects go to build/obj (git-ignored), the .so next to this file.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_PATH = os.environ.get("SYMPA_B200_LIB") or os.path.join(PKG_DIR, "libsympa_b200.so")
MAX_N = 10

EXTRA_FLAGS = os.environ.get("SYMPA_NVCC_EXTRA", "").split()

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + EXTRA_FLAGS


def _nvcc():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (set NVCC or CUDA_HOME)")
    return cand


def _sources_digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/sympa_b200.h"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(args):
    src, obj, defs, verbose = args
    cmd = [_nvcc()] + NVCC_FLAGS + defs + ["-Xptxas", "-v", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj


def build_library(force=False, jobs=None, verbose=False):
    """Compile (if sources changed) and return the path of libsympa_b200.so."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "digest.txt")
    digest = _sources_digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    tasks = [(os.path.join(CSRC, "sympa_b200.cu"), os.path.join(OBJ_DIR, "sympa_b200.o"), [], verbose)]
    for n in range(1, MAX_N + 1):
        tasks.append((os.path.join(CSRC, "pair_kernels_n.cu"), os.path.join(OBJ_DIR, "pair_kernels_%d.o" % n),
                      ["-DSYMPA_TU_N=%d" % n], verbose))
    jobs = jobs or min(len(tasks), os.cpu_count() or 4)
    with concurrent.futures.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(_compile, tasks))
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(p)

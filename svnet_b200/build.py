"""Build libsvnet_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m svnet_b200.build [--force]

One object per source under svnet_b200/build/ (recompiled when the source or a header is newer), then one
link step.  The shared object is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsvnet_b200.so")
SOURCES = ["misc.cu", "knn.cu", "knn_tc.cu", "gate.cu", "edge_xyz.cu", "edge.cu", "edge_fast.cu", "edge_tc.cu", "edge_fp_tc.cu", "rows.cu",
           "rows_fast.cu", "binlinear_tc.cu", "gemm.cu", "gemm_tc.cu", "gemm_tcgen05.cu", "gemm_tc3.cu", "head.cu", "seg_head.cu", "model.cu", "collective.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "svnet_b200.h"), os.path.abspath(__file__)]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("SVNET_NVCC_EXTRA", "").split()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or force or _stale(LIB, objs):
        run([nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB] + objs + ["-ldl"])
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))

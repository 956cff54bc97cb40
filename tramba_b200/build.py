"""Build libtramba_b200.so (C-ABI, sm_100a only) in-tree with nvcc.

    python -m tramba_b200.build [--force] [--verbose]

The library links the CUDA runtime statically and has no torch / Python dependency; the
Python host loads it with ctypes (tramba_b200/_lib.py).  Built artefacts are git-ignored but
travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libtramba_b200.so")
SOURCES = ["capi.cu", "sscan_fwd.cu", "sscan_bwd.cu", "ss2d_fused_fwd.cu", "ss2d_fused_bwd.cu", "cross_scan.cu",
           "dct2d.cu", "dct2d_tc.cu", "layernorm2d.cu", "dwconv2d.cu", "proj_gather.cu", "orders.cpp"]
HEADERS = ["common.cuh", "sscan.cuh", "ss2d_fused.cuh", os.path.join("..", "..", "include", "tramba_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"] + \
        os.environ.get("TRAMBA_NVCC_EXTRA", "").split()        # e.g. -DTRAMBA_PACK=0 for an A/B build


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def sources_newer_than_lib():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return _stale(LIB, deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, os.path.splitext(os.path.basename(s))[0] + ".o")
        if force or _stale(obj, [src] + hdrs):
            if s.endswith(".cu"):
                cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            else:  # host-only translation unit
                cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed: " + " ".join(cmd))
    objs = [os.path.join(OBJ, os.path.splitext(os.path.basename(s))[0] + ".o") for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + [
               "-Xcompiler", "-fPIC", "-cudart", "static",
               "-Xlinker", "--exclude-libs,ALL"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

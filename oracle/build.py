"""Compile the oracle's C restatement (gcc, OpenMP).  Output: oracle/_build/libsscan_ref.so.

There is no `oracle/_ref`: the reference is pure Python (zero .c/.cu/.cpp files)
and its recurrence lives in an un-vendored CUDA extension, so nothing of it compiles here.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libsscan_ref.so")


def build(force=False):
    src = os.path.join(HERE, "sscan_ref.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
           "-fno-fast-math", src, "-o", OUT, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))

"""Builds the sm_100a shared library (libosb_b200.so) in-tree with nvcc.

    python optimization-solvers_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  -fmad=false: the reference (Rust) never contracts a*b+c, and
the active set is defined by exact float equality, so fused multiply-adds exist only where the
kernels spell them out.  -lineinfo keeps the ncu source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libosb_b200.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "engine.cu", "objectives.cu", "vec_kernels.cu", "qn_kernels.cu", "qn_small.cu", "qn_device.cu", "newton.cu", "logistic.cu",
           "batched.cu", "dist.cu", "qn_iter.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "--extended-lambda", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-ccbin", "/usr/bin/g++"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "optsolv_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    procs, objs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or hdr_time > os.path.getmtime(o):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT + ".tmp"] + objs + ["-ldl", "-ccbin", "/usr/bin/g++"]
        subprocess.check_call(cmd)
        os.replace(OUT + ".tmp", OUT)  # (a snapshot of the tree never sees a half-written library)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

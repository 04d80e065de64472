#!/usr/bin/env python3
"""Builds the sm_100a engine (era_zkevm_circuits_b200/libzkc_b200.so) and the CPU oracle
(oracle/liborc.so, test infrastructure).  nvcc cross-compiles without a GPU."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "era_zkevm_circuits_b200", "csrc")
LIB = os.path.join(ROOT, "era_zkevm_circuits_b200", "libzkc_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-cudart", "static"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_engine(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    objdir = os.path.join(ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer(o, deps):
            cmd = ["nvcc", *NVCC_FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    if force or procs or _newer(LIB, objs):
        subprocess.check_call(["nvcc", *NVCC_FLAGS, "-shared", "-o", LIB, *objs])
    return LIB


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return os.path.join(ROOT, "oracle", "liborc.so")


if __name__ == "__main__":
    build_engine(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle()
    print("built", LIB)

"""Builds libtamago_b200.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc.

python -m tamago_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtamago_b200.so")
SOURCES = ["tg_engine.cu", "tg_record.cpp"]
HEADERS = ["tg_common.cuh", "tg_detmath.cuh", "tg_board.cuh", "tg_tree.cuh", "tg_search.cuh", "tg_block.cuh", "tg_dualnet.cuh",
           os.path.join("..", "..", "include", "tamago_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              # search math must round like numpy/torch scalar arithmetic: no FMA contraction (the DualNet
              # kernels use explicit fmaf where they want it)
              "-fmad=false", "--expt-relaxed-constexpr", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

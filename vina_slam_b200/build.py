"""Build libvina_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Two kinds of translation units:
  * scan_kernels.cu / map_kernels.cu  -> -fmad=false: reference operation order, one rounding per op
  * iekf_kernel.cu / vn_ctx.cu / host -> FMA allowed; decision-bearing expressions use explicit
                                         single-rounding intrinsics (vn_math.cuh)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OUT = os.path.join(HERE, "libvina_b200.so")
OBJ = os.path.join(HERE, "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = (["-DVINA_SPLIT_TRACE"] if os.environ.get("VINA_SPLIT_TRACE") else []) + (["-DVINA_LOOP_TRACE"] if os.environ.get("VINA_LOOP_TRACE") else []) + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function,-Wno-unknown-pragmas",
          "-I", os.path.join(HERE, "..", "include")]

UNITS = [
    (os.path.join(CSRC, "scan_kernels.cu"), ["-fmad=false"]),
    (os.path.join(CSRC, "map_kernels.cu"), ["-fmad=false"]),
    (os.path.join(CSRC, "front_kernels.cu"), ["-fmad=false"]),
    (os.path.join(CSRC, "shard_kernels.cu"), ["-fmad=false"]),
    (os.path.join(CSRC, "iekf_kernel.cu"), ["-DIEKF_THREADS=" + os.environ.get("VINA_IEKF_THREADS", "832"),
                                            "-DIEKF_BLOCKS_PER_SM=" + os.environ.get("VINA_IEKF_BLOCKS_PER_SM", "1")]),
    (os.path.join(CSRC, "ba_kernels.cu"), []),
    (os.path.join(CSRC, "vn_ctx.cu"), []),
    (os.path.join(HOST, "vina_pipeline.cpp"), ["-x", "cu"]),
    (os.path.join(HOST, "vina_ba.cpp"), ["-x", "cu"]),
    (os.path.join(HOST, "vina_sync.cpp"), ["-x", "cu"]),
    (os.path.join(HOST, "vina_decode.cpp"), ["-x", "cu"]),
]
HEADERS = [os.path.join(CSRC, f) for f in ("vn_types.cuh", "vn_math.cuh", "vn_kernels.cuh", "vn_ctx.h")] + [
    os.path.join(HOST, "vina_ba.h"),
    os.path.join(HERE, "..", "include", "vina_b200.h")]


def _newer(src_list, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for src, extra in UNITS:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer([src] + HEADERS, obj):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _newer(objs, OUT):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libvelvet_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m velvet_b200.build [--force] [--verbose]

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no other targets, no PTX fallback
  -fmad=false                               no FMA contraction: fp32 results follow the written operation
                                            order and match the CPU oracle (which uses -ffp-contract=off)
  -lineinfo                                 ncu --import-source maps SASS back to these sources
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libvelvet_b200.so")
INCLUDE = os.path.normpath(os.path.join(HERE, "..", "include"))

SOURCES = ["radix_sort.cu", "seam_kernels.cu", "fused_kernels.cu", "setup_kernels.cu", "input_kernels.cu", "dd_peer.cu", "tile_plan.cpp", "grid_plan.cpp", "solver.cu", "capi.cu"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# Experiments only (never the default, never used by tests): VELVET_VARIANT=fast builds a second library with FMA
# contraction and approximate division/sqrt to measure what bit-exactness with the CPU oracle costs.
VARIANT = os.environ.get("VELVET_VARIANT", "")
if VARIANT:
    LIB = os.path.join(LIBDIR, f"libvelvet_b200_{VARIANT}.so")
    OBJDIR = os.path.join(HERE, f"build_{VARIANT}")
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    *(["-fmad=true", "-prec-div=false", "-prec-sqrt=false"] if VARIANT.startswith("fast") else ["-fmad=false"]),
    *([f"-DVT_IT_REGCAP_BLOCKS={VARIANT[-1]}"] if VARIANT[-2:-1] == "b" else []),
    *os.environ.get("VELVET_VARIANT_DEFS", "").split(),
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-I", INCLUDE,
]


FAST_FLAGS = [f for f in FLAGS if f != "-fmad=false"] + ["-fmad=true", "-prec-div=false", "-prec-sqrt=false", "-DVT_FAST_MATH=1"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "velvet_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs, jobs = [], []
    units = [(src, src + ".o", FLAGS) for src in SOURCES]
    # the float kernels a second time with relaxed math (namespace fast_math, see fused_kernels.cuh)
    units.append(("fused_kernels.cu", "fused_kernels_fast.cu.o", FAST_FLAGS))
    for src, objname, flags in units:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, objname)
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [NVCC, *flags, "-x", "cu", "-c", path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose or r.stderr.strip():
            sys.stderr.write(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

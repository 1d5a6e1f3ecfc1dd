"""In-tree build of libpolyphemus_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m polyphemus_b200.build [--force]

The shared library is written next to the sources (polyphemus_b200/lib/) so that it travels with the
repository snapshot to the GPU box; it is git-ignored, never pip-installed.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libpolyphemus_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "polyphemus_b200.h")

SOURCES = ["common.cu", "graph_build.cu", "csr.cu", "aggregate.cu", "agg_bwd_tc.cu", "bn.cu", "loss.cu", "chord.cu", "pool.cu", "dataset.cu", "rows.cu", "gemm_check.cu", "gemm_tcgen05.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".cuh", ".h"))] + [HEADER]
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS) + os.environ.get("PB_EXTRA_NVCC", "").split()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        path = os.path.join(SRC_DIR, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc, *flags, "-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

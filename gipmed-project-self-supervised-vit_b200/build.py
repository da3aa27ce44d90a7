"""Builds csrc/*.cu into the in-tree C-ABI shared library ``libb200ssl.so`` for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box; the resulting ``.so``
is git-ignored but travels to the GPU box with the repo snapshot.  Sources are compiled to objects
in parallel and re-compiled only when they (or a header) are newer than the object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200ssl.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; b200ssl has no CPU fallback and cannot be built without CUDA")
    return cand


def _newest_header() -> float:
    ts = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        if not os.path.isdir(root):
            continue
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                ts = max(ts, os.path.getmtime(os.path.join(root, f)))
    return ts


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr_ts = _newest_header()
    include = os.path.join(os.path.dirname(HERE), "include")

    def compile_one(src: str):
        obj = os.path.join(BUILD, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_ts)):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-I", CSRC, "-I", include, "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs) or 1)) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    logs = "".join(log for _, log in results)
    if logs:
        with open(os.path.join(BUILD, "ptxas.log"), "a") as f:
            f.write(logs)
    if verbose and logs:
        print(logs)
    relink = force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

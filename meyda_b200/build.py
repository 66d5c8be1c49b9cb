"""Builds meyda_b200/_lib/libmeyda_b200.so (CUDA kernels + C ABI) with nvcc for
sm_100a, in-tree, so the library travels with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libmeyda_b200.so")
SOURCES = ["capi.cu", "kernel_generic.cu", "kernel_warp.cu", "kernel_warp_mf.cu", "kernel_exact_warp.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--use_fast_math=false", "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the Meyda B200 path needs the CUDA toolkit to build")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "meyda_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    def compile_one(src):
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(LIB_DIR, src.replace(".cu", ".ptxas.txt")), "w") as f:
            f.write(r.stderr)
        return src, obj, r

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:  # one nvcc per source file
        results = list(pool.map(compile_one, SOURCES))
    objs = []
    for src, obj, r in results:
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

"""Build ``libtreedet.so`` (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

Usage: ``python -m treedetection_b200.build [--force]``.  nvcc cross-compiles
without a GPU.  The shared object lands next to the sources
(``treedetection_b200/csrc/libtreedet.so``): it is git-ignored but travels to the
GPU box with the working tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libtreedet.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # parity: every float operation is rounded as written; fused multiply-adds
    # appear only where the code spells them out (__fmaf_rn / fma)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--extended-lambda",
    "-Xcompiler", "-fPIC,-O3,-fno-fast-math,-ffp-contract=off",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    inc = os.path.join(os.path.dirname(HERE), "include")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if os.path.isdir(inc):
        hs += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    return sorted(hs)


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp.txt")
    hdr_digest = _digest(_headers())
    nvcc = _nvcc()
    inc = os.path.join(os.path.dirname(HERE), "include")

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        tag = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".sha")
        dig = _digest([src]) + hdr_digest
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == dig:
            return obj, "", False
        cmd = [nvcc, *NVCC_FLAGS, "-I", CSRC, "-I", inc, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(tag, "w") as f:
            f.write(dig)
        return obj, r.stderr, True

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [r[0] for r in results]
    rebuilt = any(r[2] for r in results)
    if verbose:
        for _, log, did in results:
            if did and log:
                sys.stderr.write(log)
    link_digest = _digest(objs)
    if rebuilt or force or not os.path.exists(LIB) or not os.path.exists(stamp) or open(stamp).read() != link_digest:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(link_digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print(path)

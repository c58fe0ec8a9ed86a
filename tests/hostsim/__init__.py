"""TEST INFRASTRUCTURE: builds tests/hostsim/hostsim.cpp with g++ (host simulation of the
sequential per-item device algorithms) and exposes it through ctypes."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "treedetection_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libhostsim.so")

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HERE, "hostsim.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", src, "-I", CSRC,
               "-o", LIB]
        subprocess.run(cmd, check=True)
    _lib = C.CDLL(LIB)
    return _lib

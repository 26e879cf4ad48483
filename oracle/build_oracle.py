"""TEST INFRASTRUCTURE ONLY — builds the C++ CPU oracle into oracle/lib/liboracle_cpu.so (git-ignored, travels
with gpurun).  The reference itself is Rust with un-vendored git dependencies and cannot be compiled in this image
(no cargo/rustc), so there is no oracle/_ref/ build; see DESIGN.md."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp")
LIB = os.path.join(HERE, "lib", "liboracle_cpu.so")
# portable x86-64-v3-ish flags instead of -march=native: the .so built here also runs on the GPU box's host CPU
FLAGS = ["-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-mavx2", "-mbmi2", "-madx", "-Wall", "-Wno-unused-function"]


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    cmd = ["g++", *FLAGS, os.path.join(SRC, "capi.cpp"), "-o", LIB]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

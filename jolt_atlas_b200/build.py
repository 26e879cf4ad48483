"""Build recipe: nvcc -> in-tree shared libraries (sm_100a only, -lineinfo so ncu source pages map back).

  jolt_atlas_b200/lib/libjolt_atlas_b200.so   kernels + C ABI (include/jolt_atlas_b200.h)

Run `python -m jolt_atlas_b200.build` (or __graft_entry__.build()).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libjolt_atlas_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-fopenmp,-mbmi2,-madx",   # mulx / adcx / adox for the host-side Montgomery glue (every x86-64 server CPU since 2015)
    "--expt-relaxed-constexpr",
] + (["-DJA_MUL_CALL"] if os.environ.get("JA_MUL_CALL") else [])


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def sources() -> list[str]:
    out = []
    for d, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp")):
                out.append(os.path.join(d, f))
    out.append(os.path.join(ROOT, "include", "jolt_atlas_b200.h"))
    return sorted(out)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sources()
    if not force and _newer(LIB, srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cus = [s for s in srcs if s.endswith((".cu", ".cpp"))]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for cu in cus:
        obj = os.path.join(HERE, "build", os.path.basename(cu) + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", cu, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cu, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cu, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {cu}:\n{out}")
        if verbose and out:
            print(out, file=sys.stderr)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lgomp", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

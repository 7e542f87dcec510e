"""In-tree build of libgears_physics_cuda.so (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libgears_physics_cuda.so")
SRC = os.path.join(HERE, "csrc", "gp_api.cu")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # Rust never contracts a*b+c; bit parity needs the same here
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return [SRC] + sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + \
        [os.path.join(ROOT, "include", "gears_physics_cuda.h")]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("GP_NVCC_EXTRA", "").split()  # e.g. -DGP_COOP_THREADS=640 (tuning experiments)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC, "-ldl"]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    subprocess.check_call(cmd, env=env)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))

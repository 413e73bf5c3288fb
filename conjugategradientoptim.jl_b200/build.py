"""Builds libcgoptim.so (hand-written sm_100a CUDA kernels + the C ABI of include/cgoptim.h).

In-tree build with nvcc; no torch extension machinery, no JIT cache: the .so travels with the
repository snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libcgoptim.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",          # Julia does not contract a*b+c: no FMA anywhere
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(_HERE, "..", "include", "cgoptim.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("CGO_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", os.path.join(_HERE, "..", "include"), "-o", LIB] + sources() + ["-ldl"]
    env = dict(os.environ)
    # /opt/gcc's wrapper lacks some runtime specs; the distro compiler is the supported host cc
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or p.returncode != 0:
        print(p.stdout)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed building libcgoptim.so")
    with open(os.path.join(_HERE, "csrc", "ptxas_report.txt"), "w") as f:
        f.write(p.stdout)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose=True))

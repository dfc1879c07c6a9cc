"""Build recipe for libwf_b200.so — explicit nvcc, sm_100a only, in-tree output.

    python -m weldformfem_b200.build [--force]

The kernel source is compiled twice (strict: -fmad=false, fast: -fmad=true) into one library;
see csrc/wf_kernels.cu.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libwf_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + ARCH


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "wf_engine.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    # use the system g++ as host compiler (the image's /opt/gcc lacks libgomp specs; not needed here but keep one toolchain)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    jobs = [
        (["-fmad=false", "-DWF_NS=wf_strict"], "wf_kernels.cu", "wf_kernels_strict.o"),
        (["-fmad=true", "-DWF_NS=wf_fast"], "wf_kernels.cu", "wf_kernels_fast.o"),
        ([], "wf_engine.cu", "wf_engine.o"),
        (["-fmad=false"], "wf_contact.cu", "wf_contact.o"),
        ([], "wf_mesh.cpp", "wf_mesh.o"),
    ]
    procs = []
    for flags, src, obj in jobs:
        cmd = [nvcc] + ccbin + COMMON + flags + ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out)
        if verbose and out.strip():
            print(out)
    link = [nvcc] + ccbin + ARCH + ["-shared", "-o", LIB] + [os.path.join(OBJ, j[2]) for j in jobs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

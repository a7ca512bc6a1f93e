"""Builds libphylocsf_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m phylocsf_b200.build [--force]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libphylocsf_b200.so")
CLI = os.path.join(HERE, "bin", "PhyloCSF")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2", "-shared",
] + os.environ.get("PCSF_NVCC_EXTRA", "").split()


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def deps():
    out = [os.path.join(HERE, "..", "include", "phylocsf_b200.h")]
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".hpp", ".h", ".inc", ".cpp"))]
    return out


def stale():
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(CLI))
    return any(os.path.getmtime(d) > t for d in deps())


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    subprocess.check_call(cmd)
    # the drop-in command line: C++ host over the C ABI, finds the library next to itself
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-pthread", "-o", CLI, os.path.join(CSRC, "host", "cli_main.cpp"),
                           "-L" + HERE, "-lphylocsf_b200", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

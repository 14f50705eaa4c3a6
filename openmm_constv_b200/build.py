"""Builds the in-tree native libraries with nvcc for sm_100a (cross-compiles without a GPU).

    python -m openmm_constv_b200.build [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    openmm_constv_b200/libconstv_b200.so   the C-ABI library (include/constv_b200.h)
    openmm_constv_b200/plugin/lib/...      the OpenMM plugin libraries over it (plugin/Makefile)
    openmm_constv_b200/plugin/bin/Test*    their C++ tests
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libconstv_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # every fused multiply-add is an explicit fma
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2,-Wall",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_capi(force: bool = False, verbose: bool = False) -> str:
    sources = [os.path.join(CSRC, "constv_capi.cu")]
    deps = sources + [os.path.join(CSRC, "constv_kernels.cuh"), os.path.join(CSRC, "constv_drude_kernels.cuh"),
                      os.path.join(CSRC, "constv_drude_capi.inc"), os.path.join(CSRC, "constv_settle_kernels.cuh"),
                      os.path.join(CSRC, "constv_settle_capi.inc"),
                      os.path.join(ROOT, "include", "constv_b200.h"), os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        cmd = [nvcc(), *NVCC_FLAGS, "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               "-o", LIB, *sources]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(HERE, "build_ptxas.log")
        with open(log, "w") as fh:
            fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building libconstv_b200.so")
        if verbose:
            print(res.stderr)
    return LIB


PLUGIN_DIR = os.path.join(HERE, "plugin")


def build_plugin(force: bool = False, verbose: bool = False) -> str:
    """The OpenMM-facing C++ libraries (libOpenMMConstV, lib/plugins/libConstVPluginCUDA) and their
    C++ tests, via plugin/Makefile: against the OpenMM shim, or a real OpenMM when OPENMM_DIR is
    set in the environment."""
    cmd = ["make", "-C", PLUGIN_DIR, "-j8"]
    if force:
        subprocess.run(["make", "-C", PLUGIN_DIR, "clean"], capture_output=True, text=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("building the plugin libraries failed")
    if verbose:
        print(res.stdout)
    return os.path.join(PLUGIN_DIR, "lib")


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_capi(force, verbose)
    build_plugin(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)

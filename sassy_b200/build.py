"""Builds the native libraries in-tree (sassy_b200/lib/).

* ``libsassy_b200.so``     -- the product: CUDA kernels for sm_100a + C++ host + C ABI (nvcc)
* ``libsassy_b200_emu.so`` -- CPU-only emulation of the per-thread kernel logic, used by the
                              ``-m "not gpu"`` tests only (g++; never loaded by the package)

nvcc cross-compiles without a GPU, so this also runs in the CPU-only container.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")

CU_SOURCES = ["scan_kernels.cu", "post_kernels.cu", "engine.cu", "searcher.cu", "transport.cu", "peer_gather.cu"]
HEADERS = ["profile.h", "scan_core.cuh", "host_logic.h", "kernels.cuh", "engine.h", "searcher.h", "transport.h", "dna_pack.h",
           "peer_gather.h", "shard_merge.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]

LIB = os.path.join(LIBDIR, "libsassy_b200.so")
EMU = os.path.join(LIBDIR, "libsassy_b200_emu.so")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _hdrs():
    inc = os.path.join(os.path.dirname(HERE), "include")
    return [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(inc, "sassy.h"), os.path.join(inc, "sassy_gpu.h")]


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = _hdrs()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        srcp = os.path.join(CSRC, src)
        if force or _newer(obj, [srcp] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + ["-c", srcp, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(CU_SOURCES)) as ex:
        objs = list(ex.map(compile_one, CU_SOURCES))
    if force or _newer(LIB, objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
               "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


def build_variant(name: str, defines) -> str:
    """Kernel-tuning experiments: libsassy_b200_<name>.so compiled with extra -D flags
    (select it at run time with SASSY_B200_LIB=<path>)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, f"libsassy_b200_{name}.so")
    cmd = [NVCC] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-shared", "-o", out] + \
        [os.path.join(CSRC, c) for c in CU_SOURCES]
    subprocess.check_call(cmd)
    return out


def build_emu(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    src = os.path.join(CSRC, "emu.cpp")
    deps = [src] + [os.path.join(CSRC, h) for h in ("profile.h", "scan_core.cuh", "host_logic.h", "dna_pack.h", "shard_merge.h")]
    if force or _newer(EMU, deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-x", "c++", src, "-o", EMU])
    return EMU


def build_all(force: bool = False, verbose: bool = False):
    return build_lib(force, verbose), build_emu(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))

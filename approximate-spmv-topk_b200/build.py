"""In-tree build of libtopkspmv.so (sm_100a only) and of the host executable.

    python approximate-spmv-topk_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot; nothing is installed outside the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
LIB_DIR = HERE / "lib"
LIB = LIB_DIR / "libtopkspmv.so"
EXE = ROOT / "build" / "topk-spmv-b200"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = [HERE / "csrc" / "api.cu", HERE / "csrc" / "bscsr_api.cu"]
CPP_SOURCES = [HERE / "csrc" / "host_api.cpp"]
HEADERS = sorted(list((HERE / "csrc").glob("*.cuh")) + list((HERE / "csrc").glob("*.hpp")) +
                 list((HERE / "host").glob("*.hpp")) + [ROOT / "include" / "topkspmv.h"])


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    deps = CU_SOURCES + CPP_SOURCES + HEADERS + [Path(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [NVCC, "-O3", "-std=c++17", "-lineinfo", *ARCH, "--shared", "-Xcompiler", "-fPIC,-O3,-fvisibility=default",
           "-Xptxas", "-v" if verbose else "-warn-spills", "-cudart", "static",
           "-o", str(LIB), *map(str, CU_SOURCES), *map(str, CPP_SOURCES), "-lpthread"]
    print("[build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


def build_host_exe(force: bool = False) -> Path:
    EXE.parent.mkdir(exist_ok=True)
    src = HERE / "host" / "main_b200.cpp"
    deps = [src] + HEADERS + [LIB]
    if not force and not _stale(EXE, deps):
        return EXE
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O3", "-std=c++17", "-o", str(EXE), str(src), f"-L{LIB_DIR}", "-ltopkspmv",
           "-Wl,-rpath,$ORIGIN/../approximate-spmv-topk_b200/lib", "-lpthread"]
    print("[build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return EXE


if __name__ == "__main__":
    force = "--force" in sys.argv
    build_lib(force=force, verbose="-v" in sys.argv)
    if (HERE / "host" / "main_b200.cpp").exists():
        build_host_exe(force=force)
    print(LIB)

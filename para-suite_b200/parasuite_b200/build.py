"""In-tree build of the native libraries (explicit nvcc / g++ command lines, no build system).

  lib/libparasuite_b200.so : CUDA kernels (sm_100a) + C ABI + host batcher      (the product)
  lib/libps_synth.so       : seeded synthetic workload generator (bench/test input only)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

from . import abi

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX", "g++")
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def sources():
    c = abi.CSRC_DIR
    cu = [os.path.join(c, f) for f in sorted(os.listdir(c)) if f.endswith(".cu")]
    host = [os.path.join(c, f) for f in sorted(os.listdir(c)) if f.endswith(".cpp") and f != "synth.cpp"]
    hdr = [os.path.join(c, f) for f in sorted(os.listdir(c)) if f.endswith((".h", ".cuh"))]
    hdr.append(os.path.join(abi.INCLUDE_DIR, "parasuite_b200.h"))
    return cu, host, hdr


def build_all(force: bool = False, verbose: bool = True) -> None:
    os.makedirs(abi.LIB_DIR, exist_ok=True)
    cu, host, hdr = sources()
    out = abi.lib_path()
    if force or _newer(out, cu + host + hdr):
        cmd = [NVCC, "-O3", "-std=c++17", "-lineinfo", *GENCODE, "-Xcompiler", "-fPIC,-O3,-pthread", "-shared",
               "-I", abi.INCLUDE_DIR, "-I", abi.CSRC_DIR, "-Xptxas", "-v", "-o", out, *cu, *host, "-lz", "-lpthread"]
        _run(cmd, verbose)
    synth_src = os.path.join(abi.CSRC_DIR, "synth.cpp")
    synth_out = os.path.join(abi.LIB_DIR, "libps_synth.so")
    if force or _newer(synth_out, [synth_src] + hdr):
        _run([CXX, "-O3", "-std=c++17", "-fPIC", "-pthread", "-shared", "-I", abi.INCLUDE_DIR, "-o", synth_out,
              synth_src], verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)

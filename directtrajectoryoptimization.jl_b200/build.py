"""Builds libdto.so in-tree: the C-ABI host runtime (csrc/dto_runtime.cpp, g++) linked with the
model-independent kernels (csrc/dto_kkt.cu: KKT consumer, csrc/dto_sqp.cu: solver bookkeeping; nvcc -gencode
arch=compute_100a,code=sm_100a)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
# DTO_LIB: use this prebuilt runtime instead (A/B experiments with kernel variants, tools/ab_kkt.sh); never built here
LIB = os.environ.get("DTO_LIB") or os.path.join(PKG_DIR, "libdto.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in ("dto_runtime.cpp", "dto_model_abi.h", "dto_kkt_host.inc", "dto_kkt_dev.h", "dto_kkt.cu", "dto_sqp_host.inc",
                                            "dto_sqp_dev.h", "dto_sqp.cu")]
    srcs.append(os.path.join(PKG_DIR, "..", "include", "dto.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build_runtime(force: bool = False, verbose: bool = False) -> str:
    if os.environ.get("DTO_LIB"):
        return LIB
    if not force and not _stale():
        return LIB
    objs = []
    for name in ("dto_kkt", "dto_sqp"):
        obj = os.path.join(PKG_DIR, name + ".o")
        nv = [os.path.join(CUDA_HOME, "bin", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "static", "-c", os.path.join(CSRC, name + ".cu"), "-o", obj]
        r = subprocess.run(nv, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {name}.o (nvcc, sm_100a) failed:\n" + r.stderr[-4000:])
        objs.append(obj)
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Wno-unused-function",
           "-I", os.path.join(CUDA_HOME, "include"), os.path.join(CSRC, "dto_runtime.cpp"), *objs,
           "-o", LIB + ".tmp", "-L", os.path.join(CUDA_HOME, "lib64"), "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libdto.so failed:\n" + r.stderr[-4000:])
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print("[dto] built", LIB)
    return LIB


if __name__ == "__main__":
    build_runtime(force=True, verbose=True)

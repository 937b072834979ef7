"""Build recipe for the native pieces (run by __graft_entry__.build()).

  gabby_b200/libb2l.so          CUDA engine + C-ABI (include/b2l.h), sm_100a only
  gabby_b200/libgabby_host.so   host C++ layer (safetensors, params, KV allocator, sampler,
                                Generator adapter) over the C-ABI (include/gabby_b200_host.h)

Everything is built IN-TREE so the .so files travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "g++"  # the image's $CXX points at a wrapper without OpenMP specs; use the system g++

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", ]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    out = os.path.join(PKG, "libb2l.so")
    csrc = os.path.join(PKG, "csrc")
    deps = glob.glob(os.path.join(csrc, "*")) + [os.path.join(ROOT, "include", "b2l.h")]
    if force or _newer(out, deps):
        srcs = sorted(glob.glob(os.path.join(csrc, "*.cu")))
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out] + srcs + ["-ldl"]
        subprocess.run(cmd, check=True)
    return out


def build_host(force: bool = False) -> str | None:
    hdir = os.path.join(PKG, "host")
    srcs = sorted(glob.glob(os.path.join(hdir, "*.cc")))
    if not srcs:
        return None
    out = os.path.join(PKG, "libgabby_host.so")
    deps = glob.glob(os.path.join(hdir, "*")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _newer(out, deps):
        # hidden visibility + -Bsymbolic: only the gb_* entry points (GB_API) are exported and the library's own references
        # bind inside it, so it can be linked next to gabby's gabby::inference:: / gabby::json:: symbols (integration/)
        cmd = [CXX, "-O2", "-std=c++20", "-fPIC", "-shared", "-Wall", "-fvisibility=hidden", "-fvisibility-inlines-hidden",
               "-Wl,-Bsymbolic", "-I" + os.path.join(ROOT, "include"), "-I" + hdir,
               "-o", out] + srcs + ["-L" + PKG, "-lb2l", "-Wl,-rpath,$ORIGIN", "-lpthread"]
        subprocess.run(cmd, check=True)
    return out


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    print("built", os.path.join(PKG, "libb2l.so"))

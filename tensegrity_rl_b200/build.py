"""Builds libtsg.so (hand-written CUDA, sm_100a only) in-tree with nvcc."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "tsg_api.cu")
DEPS = [SRC] + [os.path.join(PKG, "csrc", f) for f in ("tsg_core.cuh", "tsg_env.cuh", "tsg_host.h")] + [
    os.path.join(ROOT, "include", f) for f in ("tsg.h", "tsg_model.h")]
SO = os.path.join(PKG, "libtsg.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def is_stale():
    return not os.path.isfile(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS)


def build(force=False, verbose=False, extra=()):
    if not force and not is_stale():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    print("[tsg build]", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)

"""Builds libtsg.so (hand-written CUDA, sm_100a only) in-tree with nvcc."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "tsg_api.cu")
DEPS = [SRC] + [os.path.join(PKG, "csrc", f) for f in ("tb_simt.h", "tb_model.h", "tb_math.h", "tb_mpr.h", "tb_core.cuh", "tb_env.cuh")] + [
    os.path.join(ROOT, "include", f) for f in ("tsg.h", "tsg_model.h")]
SO = os.path.join(PKG, "libtsg.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


STAMP = SO + ".stamp"   # hash of sources + flags the existing .so was built from (travels with it to the GPU box)


def _digest(extra=()):
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + list(extra)).encode())
    for d in DEPS:
        h.update(open(d, "rb").read())
    return h.hexdigest()


def is_stale(extra=()):
    if not os.path.isfile(SO) or not os.path.isfile(STAMP):
        return True
    return open(STAMP).read().strip() != _digest(extra)


def build(force=False, verbose=False, extra=()):
    if not force and not is_stale(extra):
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    print("[tsg build]", " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    open(STAMP, "w").write(_digest(extra))
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)

"""Builds lib/libpiql_b200.so from csrc/*.cu with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("backup.cu", "tile.cu", "table.cu", "rollout.cu", "peer.cu", "api.cu")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("common.cuh", "algebra.cuh", "dtables.cuh", "internal.h")] + \
      [os.path.join(HERE, "..", "include", "piql_b200.h")]
OUT = os.path.join(HERE, "lib", "libpiql_b200.so")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in SRC + HDR)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: an experimental variant (tools/vi_ab.py) built beside the product library, selected with PIQL_B200_LIB."""
    if out is None and not (force or needs_build()):
        return OUT
    out = out or OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + SRC + ["-o", out]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""In-tree build of the CUDA extension (libb200env.so) for sm_100a.  nvcc cross-compiles
without a GPU; the built .so travels to the GPU box with the repo snapshot."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "csrc", "b200env.cu"), os.path.join(HERE, "csrc", "b200env_v2p.cu")]
HDRS = [os.path.join(os.path.dirname(HERE), "include", h) for h in ("b200env.h", "b200env_v2p.h", "b200ball.h")] + [os.path.join(HERE, "csrc", f) for f in ("packed.cuh", "packed3.cuh", "ballgen.cuh", "dyn_common.cuh")]
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.environ.get("B200ENV_LIB", os.path.join(LIB_DIR, "libb200env.so"))  # override: A/B kernel variants

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177,550"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SRCS + HDRS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRCS
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))

"""TEST INFRASTRUCTURE: builds the CPU lane emulator of the packed articulated step (tests/emu/emu_packed.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vid2player3d_b200", "csrc")


def build(variant="packed", flags=()):
    """variant: packed | packed3; flags: extra -D options (e.g. ("-DPK_CONTACT_COMPACT=1",)) - they become part of the library name"""
    tag = variant + "".join("_" + f.lstrip("-D").replace("=", "") for f in flags)
    out = os.path.join(HERE, "build", f"libemu_{tag}.so")
    srcs = [os.path.join(HERE, "emu_packed.cpp"), os.path.join(HERE, "cuda_compat.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in srcs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O1", "-std=c++17", "-pthread", "-shared", "-fPIC", "-x", "c++", "-Wno-unknown-pragmas", "-ffp-contract=off",
           f"-DEMU_{variant.upper()}=1", *flags, "-o", out, os.path.join(HERE, "emu_packed.cpp")]
    subprocess.check_call(cmd)
    return out

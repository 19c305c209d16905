"""Build libpbr.so in-tree with nvcc for sm_100a (no GPU needed; nvcc cross-compiles).

    python -m pyremot_b200.build [--force]

The device code (csrc/pbr_device.cuh) is embedded into the library as text and compiled at run time by
NVRTC together with the model's rate expressions; this script only builds the host side + embeds the text.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpbr.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _embed():
    for name in ("pbr_device", "pbr_math"):
        src = open(os.path.join(CSRC, name + ".cuh")).read()
        assert ')PBRSRC"' not in src
        out = os.path.join(CSRC, name + "_src.inc")
        txt = 'R"PBRSRC(' + src + ')PBRSRC"\n'
        if not os.path.exists(out) or open(out).read() != txt:
            with open(out, "w") as f:
                f.write(txt)


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, "pbr_api.cu"), os.path.join(CSRC, "pbr_device.cuh"), os.path.join(CSRC, "pbr_math.cuh"),
            os.path.join(os.path.dirname(HERE), "include", "pbr.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    _embed()
    nvcc = os.path.join(CUDA_HOME, "bin", "nvcc")
    cmd = [nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           os.path.join(CSRC, "pbr_api.cu"), "-o", LIB + ".tmp",
           "-lnvrtc", "-ldl", "-Xlinker", f"-rpath={CUDA_HOME}/lib64"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose=True)
    print(LIB)

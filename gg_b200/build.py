"""Builds gg_b200/libggcuda.so for sm_100a with nvcc (cross-compiles without a GPU).

    python -m gg_b200.build [--force] [--verbose]

Objects land in gg_b200/csrc/build/, the shared library next to this file (git-ignored, but it
travels with the repo snapshot to the GPU box).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libggcuda.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]

# (source, extra flags). pipeline.cu feeds integer decisions from float32 math that must match
# the Go reference bit-for-bit, so FMA contraction is off there. fine.cu: the reference's area
# formula cancels catastrophically for near-vertical segments ((b + 0.5(d^2-c^2) - xmin)/(xmax-xmin)),
# so a contracted evaluation differs from the Go one by up to ~13/255 on such pixels; contraction is
# off there too and the compositing code asks for FMA explicitly (fmaf) where 1 ulp does not matter.
UNITS = [
    ("pipeline.cu", ["-fmad=false"]),
    ("fine.cu", ["-fmad=false"]),   # (-ftz=true was tried for cheaper divisions: 3 % faster, 17 wrong pixels on the 4K scene -- not worth it)
    ("api.cu", []),
    ("host_scene.cpp", []),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "ggcuda.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    nvcc = _nvcc()
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(bdir, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(OUT, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

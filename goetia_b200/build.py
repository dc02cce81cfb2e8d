"""Build goetia_b200/libgoetia_b200.so (the C-ABI library) for sm_100a with nvcc, in-tree.

    python -m goetia_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the source snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "capi.cu")
SRC_HOST = os.path.join(HERE, "csrc", "pack_simd.cpp")  # host-only helpers (AVX2 where the CPU has it)
DEPS = [SRC, SRC_HOST] + [os.path.join(HERE, "csrc", f) for f in ("kernels.cuh", "sketch.cuh", "sketch_host.inc", "bucket.cuh",
                                                          "bucket_host.inc", "fastx_host.inc")] + [os.path.join(ROOT, "include", "goetia_b200.h")]
OUT = os.path.join(HERE, "libgoetia_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-shared",
]
LIBS = ["-lz"]  # gz-transparent FASTX front end (the reference links zlib for the same purpose)


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC, SRC_HOST] + LIBS
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)

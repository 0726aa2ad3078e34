"""Builds psxavenc_b200/libpsxav_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the shared library links the CUDA runtime statically and
has no Python or torch dependency. Usage: python -m psxavenc_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpsxav_b200.so")
SOURCES = ["bs_encode.cu", "adpcm_encode.cu", "capi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "psxav_b200.h")]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return LIB


def build_variant(path, defines):
    """A/B builds for tools/ab.sh: the same sources with extra -D switches, written to `path`
    (loaded through PSXB200_LIB=...). Not used by the product."""
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + ["-o", path] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return path


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

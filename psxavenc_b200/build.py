"""Builds psxavenc_b200/libpsxav_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the shared library links the CUDA runtime statically and
has no Python or torch dependency. Sources are compiled to objects in parallel (build/ is
git-ignored) and linked. Usage: python -m psxavenc_b200.build [--force] [-v]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpsxav_b200.so")
SOURCES = ["bs_encode.cu", "adpcm_encode.cu", "color_convert.cu", "capi_bs.cu", "capi_audio.cu", "capi_multi.cu", "capi_color.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]
LDFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
           [os.path.join(ROOT, "include", "psxav_b200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps)


def _compile(sources, objdir, defines, force, verbose):
    os.makedirs(objdir, exist_ok=True)
    headers = _headers()
    jobs = []
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src[:-3] + ".o")
        if force or _newer(obj, [path] + headers):
            cmd = [NVCC] + CFLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, path]
            jobs.append(cmd)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for res in pool.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
            if verbose or res.returncode:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode:
                raise subprocess.CalledProcessError(res.returncode, res.args)
    return [os.path.join(objdir, s[:-3] + ".o") for s in sources], bool(jobs)


def build(force=False, verbose=False):
    objs, rebuilt = _compile(SOURCES, os.path.join(HERE, "build"), [], force, verbose)
    if rebuilt or _newer(LIB, objs):
        subprocess.run([NVCC] + LDFLAGS + ["-o", LIB] + objs, check=True)
    return LIB


def build_variant(path, defines):
    """A/B builds (tools/ab_variants.sh): the same sources with extra -D switches, written to
    `path` (loaded through PSXB200_LIB=...). Not used by the product."""
    objdir = os.path.join(HERE, "build", "variant_" + os.path.basename(path).replace(".", "_"))
    objs, _ = _compile(SOURCES, objdir, defines, True, False)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    subprocess.run([NVCC] + LDFLAGS + ["-o", path] + objs, check=True)
    return path


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

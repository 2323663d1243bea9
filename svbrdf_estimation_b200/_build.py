"""Builds ``libsvbrdf_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python -m svbrdf_estimation_b200._build [--force] [--verbose]

The library has a pure C ABI (include/svbrdf_b200.h) and links only the CUDA runtime, so plain
``nvcc -shared`` is enough; ``torch.utils.cpp_extension`` is deliberately not involved.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsvbrdf_b200.so")
SOURCES = ["kernels.cu", "host_ctx.cu"]
HEADERS = [os.path.join(CSRC, "shading.cuh"), os.path.join(CSRC, "pixel_ops.cuh"), os.path.join(CSRC, "internal.h"),
           os.path.join(ROOT, "include", "svbrdf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the CUDA sources into ``LIB`` if it is missing or older than its inputs."""
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB + ".tmp"] \
        + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%d): %s" % (proc.returncode, " ".join(cmd)))
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

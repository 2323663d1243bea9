"""Builds ``libsvbrdf_b200.so`` in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python -m svbrdf_estimation_b200._build [--force] [--verbose]

The library has a pure C ABI (include/svbrdf_b200.h) and links only the CUDA runtime, so plain
``nvcc -shared`` is enough; ``torch.utils.cpp_extension`` is deliberately not involved.  A hash of the
sources and flags is compiled in (``svbrdf_b200_build_id()``); ``is_stale()`` compares it with the sources on
disk, so an edited kernel can never be served by an old binary.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libsvbrdf_b200.so")
SOURCES = ["kernels.cu", "host_ctx.cu", "scene_sampler.cpp"]
HEADERS = [os.path.join(CSRC, "shading.cuh"), os.path.join(CSRC, "pixel_ops.cuh"), os.path.join(CSRC, "internal.h"),
           os.path.join(ROOT, "include", "svbrdf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-ffp-contract=off", "--cudart", "static"]
ID_MARKER = "SVB_BUILD_ID:"


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def source_id(extra_flags=()):
    """sha256 over the CUDA/C++ sources, the headers and the compiler flags (first 16 hex digits)."""
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(list(NVCC_FLAGS) + list(extra_flags)).encode())
    return h.hexdigest()[:16]


def binary_id(path=LIB):
    """The build id compiled into an existing library (read from the file, nothing is loaded); None if absent."""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError:
        return None
    i = data.find(ID_MARKER.encode())
    if i < 0:
        return None
    j = i + len(ID_MARKER)
    return data[j:j + 16].decode("ascii", "replace")


def is_stale():
    return binary_id() != source_id()


def compile_library(out, extra_flags=(), verbose=False):
    """nvcc: every source to an object file (in parallel), then one -shared link."""
    nvcc = find_nvcc()
    flags = list(NVCC_FLAGS) + list(extra_flags) + ['-DSVB_BUILD_ID="%s%s"' % (ID_MARKER, source_id(extra_flags))]
    tmp = out + ".build.%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    try:
        procs = []
        for s in SOURCES:
            obj = os.path.join(tmp, os.path.splitext(s)[0] + ".o")
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, s)]
            procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs = []
        for cmd, obj, p in procs:
            text, _ = p.communicate()
            if verbose or p.returncode != 0:
                sys.stderr.write(text)
            if p.returncode != 0:
                raise RuntimeError("nvcc failed (%d): %s" % (p.returncode, " ".join(cmd)))
            objs.append(obj)
        cmd = [nvcc] + flags + ["-shared", "-o", out + ".tmp"] + objs
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout)
            raise RuntimeError("nvcc link failed (%d): %s" % (p.returncode, " ".join(cmd)))
        os.replace(out + ".tmp", out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def build(force=False, verbose=False):
    """Compile the CUDA sources into ``LIB`` if it is missing or was built from other sources.  Safe to call from
    several processes at once (one builds, the others wait on a file lock and find a current library)."""
    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or is_stale():
                compile_library(LIB, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""TEST / BENCH INFRASTRUCTURE.  Compiles the UNMODIFIED reference modules of the hot path to Python bytecode:

    /root/reference/development/multiImage_pytorch/{utils,environment,renderers,losses}.py  ->  oracle/_ref/*.bin

``oracle/_ref/`` is git-ignored (no reference source or derivative enters the history) but travels to the GPU box with
the snapshot, like the built ``.so`` files, so ``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the
reference's own code on the box's host cores (``kind: "reference"``) instead of the restatement in
``oracle/reference_port.py`` (``kind: "port"``).  Run by ``__graft_entry__.build()`` when ``/root/reference`` exists;
``oracle/ref_loader.py`` imports the result.  The bytecode is specific to the interpreter version that wrote it (the
GPU box runs the same image); a mismatch makes the loader fall back to the port and say so.
"""
import os
import py_compile
import sys

REFERENCE = "/root/reference/development/multiImage_pytorch"
MODULES = ("utils", "environment", "renderers", "losses")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build(reference=REFERENCE, out=OUT):
    """-> list of written files ([] when the reference tree is not present, e.g. on the GPU box)."""
    if not os.path.isdir(reference):
        return []
    os.makedirs(out, exist_ok=True)
    written = []
    for name in MODULES:
        src = os.path.join(reference, name + ".py")
        dst = os.path.join(out, name + ".bin")
        py_compile.compile(src, cfile=dst, dfile="reference/%s.py" % name, doraise=True, optimize=0)
        written.append(dst)
    with open(os.path.join(out, "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    return written


if __name__ == "__main__":
    print("\n".join(build()) or "reference tree not found: nothing built")

"""Copy recipe for the reference arm of bench.py: puts an UNMODIFIED copy of the reference's `sympa` package
into oracle/_ref/ so that the GPU box - where /root/reference does not exist - can time the real thing on its
host cores (bench.py --impl reference, cpu_baseline.kind = "reference").

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (the sources never enter the repository's
history) but NOT gpurun-ignored, so the copy travels with the snapshot like a built .so.  The reference is pure
Python: there is nothing to compile.  It is imported through oracle/refstub (geoopt / tensorboardX stand-ins and
the torch.symeig shim of oracle/import_reference.py); no file of the copy is edited.

    python oracle/make_ref.py        # needs /root/reference; run by __graft_entry__.build() when it is present
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SYMPA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref():
    src = os.path.join(SRC, "sympa")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DST, "sympa")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    cmp = filecmp.dircmp(src, dst, ignore=["__pycache__"])
    assert not cmp.diff_files and not cmp.left_only, "copy differs from the reference"
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of /root/reference/sympa made by oracle/make_ref.py (git-ignored; measurement only).\n")
    return dst


if __name__ == "__main__":
    out = make_ref()
    print(out if out else f"no reference tree at {SRC}: nothing copied")
    sys.exit(0)

"""Recipe for ``oracle/_ref``: byte-compile the UNMODIFIED reference so that it can travel to the GPU box (TEST ORACLE ONLY).

The reference is pure Python (five files under ``/root/reference/models``; no native code to compile), and ``/root/reference``
does not exist on the GPU box.  ``stage()`` compiles each file where it lies to a sourceless ``.pyc`` under
``oracle/_ref/models/`` -- build outputs only, no reference source enters the repository (``oracle/_ref/`` and ``*.pyc`` are
git-ignored but not gpurun-ignored, like the product's ``.so``).  ``oracle/ref_import.py`` then imports the real modules from
there (behind the same stub modules as in the authoring container), which lets ``bench.py --impl reference`` time the
reference's own forward on the box's host cores (``cpu_baseline.kind = "reference"``) and the GPU-box tests cross-check the
restated oracle against it.

Run:  python -m oracle.stage_ref        (also called by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import os
import py_compile
import sys

REFERENCE_ROOT = os.environ.get("RCN_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ("networks", "LiteISP", "groupmix", "tcm", "raw2bit")


def staged() -> bool:
    return all(os.path.isfile(os.path.join(OUT, "models", m + ".pyc")) for m in MODULES)


def stage(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds the compiled reference afterwards."""
    src_dir = os.path.join(REFERENCE_ROOT, "models")
    if not os.path.isdir(src_dir):
        return staged()
    os.makedirs(os.path.join(OUT, "models"), exist_ok=True)
    for m in MODULES:
        src = os.path.join(src_dir, m + ".py")
        dst = os.path.join(OUT, "models", m + ".pyc")
        if force or not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/models/{m}.py", doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as f:
        f.write(f"{sys.version_info.major}.{sys.version_info.minor}\n")
    return staged()


if __name__ == "__main__":
    print("oracle/_ref staged:", stage(force="--force" in sys.argv))

"""TEST INFRASTRUCTURE ONLY -- recipe for ``oracle/_ref``: the reference's own numba kernels.

The reference (qiboteam/qibojit) is Python: its hot path is ``src/qibojit/custom_operators/gates.py``
and ``ops.py`` (numpy + numba only).  ``pip install /root/reference`` is
impossible offline (its build backend poetry-core and its dependency qibo are absent), so this
recipe takes the kernel modules where they lie under /root/reference and places them, unmodified,
under ``oracle/_ref/qibojit/custom_operators/`` -- a git-ignored build output (like a compiled
``.so``) that travels to the GPU box with the snapshot.  ``oracle/numba_ref.py`` imports them
through a package stub (``import qibojit`` itself needs qibo).

Run in the build container (needs /root/reference):   python oracle/make_ref.py
``__graft_entry__.build()`` runs it whenever /root/reference is present.
"""

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/qibojit/custom_operators"
REF_DST = os.path.join(HERE, "_ref", "qibojit", "custom_operators")
FILES = ("__init__.py", "gates.py", "ops.py")


def available():
    return all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)


def build(force=False):
    """Place the reference's kernel modules under oracle/_ref (no-op without /root/reference)."""
    if not os.path.isdir(REF_SRC):
        return available()
    os.makedirs(REF_DST, exist_ok=True)
    for f in FILES:
        dst = os.path.join(REF_DST, f)
        if force or not os.path.exists(dst):
            shutil.copyfile(os.path.join(REF_SRC, f), dst)
    with open(os.path.join(HERE, "_ref", "README"), "w") as fh:
        fh.write("Build output of oracle/make_ref.py: unmodified kernel modules of the reference "
                 "(qibojit v0.1.17, custom_operators/).  Not product source; git-ignored.\n")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference)")

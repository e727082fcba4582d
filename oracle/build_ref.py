#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own hot-path module, UNMODIFIED, for the CPU arm of bench.py.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python: its hot path is one file, ``src/model.py``
(imports torch and loguru only).  This script copies that file, byte for byte, from where it lies under
``/root/reference`` into ``oracle/_ref/src/model.py`` and records its SHA-256 next to it.  ``oracle/_ref/`` is
git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels to the GPU box like the
built ``.so`` files; ``bench.py --impl reference`` and the ``cpu_baseline`` leg import it from there
(``cpu_baseline.kind == "reference"``) and fall back to the oracle port (``"port"``) when it is absent.
Run by ``__graft_entry__.build()`` whenever ``/root/reference`` exists (the build container); never at run time.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MAU_REFERENCE_ROOT", "/root/reference")
FILES = ["src/model.py"]


def build(verbose=False):
    if not os.path.isdir(REF_ROOT):
        return None
    out = os.path.join(HERE, "_ref")
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(out, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("\n".join(manifest))
    return out


if __name__ == "__main__":
    r = build(verbose=True)
    print(r if r else f"{REF_ROOT} not present: nothing built", file=sys.stderr)

"""TEST / BASELINE INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference for the GPU box.

``/root/reference`` exists only in the build container.  To time the reference's own NumPy path
beside the GPU in the same run (``bench.py``: ``cpu_baseline_reference`` and ``--impl reference``)
and to drive the reference's own agents against the drop-in module on a GPU box
(``tests/test_gpu_reference_loop.py``), this script copies the three files of the reference that
the episode loop imports -- ``drone_env.py``, ``utils.py``, ``SAC_agents.py`` -- byte for byte into
``oracle/_ref/`` and byte-compiles them.  ``oracle/_ref/`` is git-ignored (never part of the
history, never part of the product) but travels with the gpurun snapshot like a built ``.so``.

Nothing here is imported by the product package.  ``__graft_entry__.build()`` runs it when
``/root/reference`` is present; on the GPU box the staged files are used as they are.
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DRONESTEP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("drone_env.py", "utils.py", "SAC_agents.py")


def staged() -> bool:
    return all(os.path.isfile(os.path.join(OUT, f)) for f in FILES)


def make(force: bool = False) -> str | None:
    """Copy + byte-compile; returns the staging directory, or None when the reference is absent."""
    if not os.path.isfile(os.path.join(REF_ROOT, FILES[0])):
        return OUT if staged() else None
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for f in FILES:
        src, dst = os.path.join(REF_ROOT, f), os.path.join(OUT, f)
        with open(src, "rb") as fh:
            digest = hashlib.sha256(fh.read()).hexdigest()
        manifest[f] = digest
        if force or not os.path.isfile(dst) or hashlib.sha256(open(dst, "rb").read()).hexdigest() != digest:
            shutil.copyfile(src, dst)
        py_compile.compile(dst, cfile=dst + "c", doraise=True)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "sha256": manifest,
                   "note": "unmodified copies; git-ignored; baseline / test infrastructure only"}, fh, indent=1)
    return OUT


if __name__ == "__main__":
    print(make(force=True))

"""Recipe for ``oracle/_ref/``: a checkout of the UNMODIFIED reference detector for the GPU box.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; the GPU box gets
the repo snapshot.  ``oracle/_ref/`` is git-ignored (the reference's sources never enter this
repository's history) but NOT gpurun-ignored, so what this recipe puts there travels with the
snapshot like a built ``.so`` does.  It holds, byte for byte:

    oracle/_ref/nbm_model/**.py     the reference package (detector network, run_detection, nbm_detect)
    oracle/_ref/bird_dict.json      species -> id map read by run_detection.py:70-73

and a ``MANIFEST.json`` with the sha256 of every file, so a test can tell that the copy on the
box is the one that was taken here.  The reference is pure Python: there is nothing to compile.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s reference/baseline legs use it (through
``oracle/ref_shims.py``); the product package finds the detector network by PYTHONPATH, as any
user of the reference would provide it.

    python -m oracle.build_ref            # (re)create oracle/_ref from /root/reference
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(src: str = SRC, dst: str = DST, verbose: bool = False) -> str | None:
    """Copy the reference package when ``src`` exists; otherwise leave ``dst`` as it is (the GPU box uses
    the copy that came with the snapshot).  Returns ``dst`` or None when neither exists."""
    if not os.path.isfile(os.path.join(src, "nbm_model", "run_detection.py")):
        return dst if os.path.isfile(os.path.join(dst, "MANIFEST.json")) else None
    manifest = {}
    tmp = dst + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    for root, dirs, files in os.walk(os.path.join(src, "nbm_model")):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for name in files:
            if not name.endswith(".py"):
                continue
            p = os.path.join(root, name)
            rel = os.path.relpath(p, src)
            os.makedirs(os.path.dirname(os.path.join(tmp, rel)), exist_ok=True)
            shutil.copyfile(p, os.path.join(tmp, rel))
            manifest[rel] = _sha(p)
    shutil.copyfile(os.path.join(src, "bird_dict.json"), os.path.join(tmp, "bird_dict.json"))
    manifest["bird_dict.json"] = _sha(os.path.join(src, "bird_dict.json"))
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(dst, ignore_errors=True)
    os.rename(tmp, dst)
    if verbose:
        print(f"{dst}: {len(manifest)} files from {src}")
    return dst


def verify(dst: str = DST) -> bool:
    """True when every file of the manifest is present with the recorded sha256."""
    mp = os.path.join(dst, "MANIFEST.json")
    if not os.path.isfile(mp):
        return False
    with open(mp) as f:
        manifest = json.load(f)
    return all(os.path.isfile(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h for rel, h in manifest.items())


if __name__ == "__main__":
    out = build(verbose=True)
    sys.exit(0 if out and verify(out) else 1)

"""Recipe that populates ``baseline/_ref/`` with the UNMODIFIED reference tree.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  ``baseline/_ref`` is git-ignored (reference sources never enter this
repository's history) but not gpurun-ignored, so the copy travels to the GPU box, where ``/root/reference`` does not
exist.  ``pip install --target baseline/_ref /root/reference`` is not applicable: the reference is a flat directory of
scripts with no build target (its ``pyproject.toml`` only configures black / isort), and its solver dependency
``torchdiffeq==0.2.2`` is in neither the image nor ``/opt/wheelhouse``.  So the "install" is a file copy of the Python
sources, shell scripts, configs and published results (notebooks are skipped), plus a manifest of SHA-256 digests that
``tests/test_reference_install.py`` checks to prove the files are byte-identical to the source tree.

    python -m oracle.install_reference [--src /root/reference]

``__graft_entry__.build()`` calls :func:`install` whenever the source tree is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
SKIP_EXT = (".ipynb", ".pyc")


def _digest(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(src: str = "/root/reference", dest: str = DEST, verbose: bool = True) -> bool:
    """Copy the reference tree; returns False (and does nothing) when ``src`` does not exist."""
    if not os.path.isfile(os.path.join(src, "model.py")):
        return False
    manifest = {}
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__")]
        for name in files:
            if name.endswith(SKIP_EXT):
                continue
            s = os.path.join(base, name)
            rel = os.path.relpath(s, src)
            d = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.isfile(d) and _digest(d) == _digest(s)):
                shutil.copyfile(s, d)
            manifest[rel] = _digest(d)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("[reference] {} files copied unmodified to {}".format(len(manifest), dest), flush=True)
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    sys.exit(0 if install(src) else 1)

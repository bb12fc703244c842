"""Stage the UNMODIFIED reference files of the hot path into the git-ignored ``baseline/_ref/``.

TEST / MEASUREMENT INFRASTRUCTURE (SURVEY.md §8c last row).  ``/root/reference``
exists only in the build container; ``gpurun`` snapshots ``/root/repo`` (git-ignored files included),
so a byte-for-byte copy under ``baseline/_ref/`` is what lets the GPU box

  * time the reference's own ``words_loss`` / ``GlobalAttentionGeneral`` (CPU arm of ``bench.py`` with
    ``cpu_baseline.kind = "reference"``, and the eager same-GPU arm with ``cfg.CUDA = True``),
  * run the reference's own ``models.Gen`` / ``train.Trainer.DAMSM_loss`` with ``eegan_b200.install()``
    routing ``miscc.DAMSM_losses`` and ``sync_batchnorm`` to this package (bench ``--config c4``,
    tests/test_gpu_reference_in_loop.py).

Nothing is modified and nothing of it enters the git history (``.gitignore``: ``baseline/_ref/``); the
product package never imports it.  ``python oracle/stage_reference.py`` (also run by
``__graft_entry__.build()`` when ``/root/reference`` is mounted) writes the files plus a manifest with
their SHA-256 so that a test can hold the staged copy to the mounted original.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DEFAULT = os.environ.get("EEGAN_REFERENCE_ROOT", "/root/reference")
DST_DEFAULT = os.path.join(ROOT, "baseline", "_ref")

# the path's own files and the modules that call it (train.py -> models.py / DAMSM.py);
# datasets.py only because train.py imports it at module level
FILES = [
    "miscc/__init__.py", "miscc/config.py", "miscc/DAMSM_losses.py", "miscc/utils.py",
    "sync_batchnorm/__init__.py", "sync_batchnorm/batchnorm.py", "sync_batchnorm/comm.py", "sync_batchnorm/replicate.py",
    "models.py", "DAMSM.py", "train.py", "datasets.py",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(src: str = SRC_DEFAULT, dst: str = DST_DEFAULT, quiet: bool = False) -> bool:
    """Copy FILES from ``src`` to ``dst``; returns False (and does nothing) when ``src`` is not mounted."""
    if not os.path.isfile(os.path.join(src, "miscc", "DAMSM_losses.py")):
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files into %s" % (len(FILES), dst))
    return True


def staged(dst: str = DST_DEFAULT) -> bool:
    return os.path.isfile(os.path.join(dst, "miscc", "DAMSM_losses.py")) and os.path.isfile(os.path.join(dst, "MANIFEST.json"))


if __name__ == "__main__":
    ok = stage()
    if not ok:
        print("reference not mounted at %s: nothing staged" % SRC_DEFAULT)
    sys.exit(0)

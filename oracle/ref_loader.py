"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so this module
is used only by ``oracle/make_golden.py`` (fixture generation, committed script) and by the
CPU tests that are skipped when the mount is absent.  No reference source is copied.

The reference needs ``easydict`` (not installed, no network): a minimal attribute-dict
stand-in is placed in ``sys.modules`` before import (miscc/config.py:7 only uses
``EasyDict()`` attribute get/set and ``type(x) is edict``).
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("EEGAN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "miscc", "DAMSM_losses.py"))


class _AttrDict(dict):
    """Tiny easydict.EasyDict stand-in: dict with attribute access, nested dicts wrapped."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _install_stubs():
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _AttrDict
        sys.modules["easydict"] = m


_cache = {}


def load_reference():
    """Returns a namespace with the reference modules: .losses, .config, .sync_batchnorm."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    # The reference package names (miscc, sync_batchnorm) must resolve to the reference
    # tree, not to anything of ours; our product package is `eegan_b200`, so no clash.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses = importlib.import_module("miscc.DAMSM_losses")
        config = importlib.import_module("miscc.config")
        syncbn = importlib.import_module("sync_batchnorm")
    config.cfg.CUDA = False
    ns = types.SimpleNamespace(losses=losses, config=config, cfg=config.cfg, sync_batchnorm=syncbn)
    _cache["ns"] = ns
    return ns

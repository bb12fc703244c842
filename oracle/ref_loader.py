"""Import the UNMODIFIED reference: from /root/reference (build container) or from the byte-for-byte
staged copy under the git-ignored ``baseline/_ref/`` (``oracle/stage_reference.py``; that copy travels
to the GPU box with the gpurun snapshot).

TEST / MEASUREMENT INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` (fixture generation, committed
script), by the tests that hold the oracle / the CUDA path to the live reference, and by the
reference arms of ``bench.py``.  No reference source is committed; the product never imports this.

The reference needs ``easydict`` (not installed, no network): a minimal attribute-dict stand-in is
placed in ``sys.modules`` before import (miscc/config.py:7 only uses ``EasyDict()`` attribute get/set
and ``type(x) is edict``).  ``train.py`` additionally imports ``tensorboardX.SummaryWriter`` and
``datasets.py`` imports ``nltk`` at module level; inert stand-ins are installed for those too (nothing
of them is on the measured path).
"""
import importlib
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(ROOT, "baseline", "_ref")


def _pick_root():
    env = os.environ.get("EEGAN_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", STAGED_ROOT]:
        if cand and os.path.isfile(os.path.join(cand, "miscc", "DAMSM_losses.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "miscc", "DAMSM_losses.py"))


def reference_models_available() -> bool:
    return reference_available() and all(os.path.isfile(os.path.join(REFERENCE_ROOT, f)) for f in ("models.py", "DAMSM.py", "train.py"))


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


def _install_stubs(train_too=False):
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _AttrDict
        sys.modules["easydict"] = m
    if not train_too:
        return
    try:
        import tensorboardX  # noqa: F401
    except ImportError:
        m = types.ModuleType("tensorboardX")

        class SummaryWriter:  # train.py:17, :140 — only written to when a Trainer logs
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

        m.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = m
    try:
        import nltk  # noqa: F401
    except ImportError:
        m = types.ModuleType("nltk")
        t = types.ModuleType("nltk.tokenize")

        class RegexpTokenizer:  # datasets.py:16 — used only when captions are tokenised from text files
            def __init__(self, pattern):
                import re
                self._re = re.compile(pattern)

            def tokenize(self, s):
                return self._re.findall(s)

        t.RegexpTokenizer = RegexpTokenizer
        m.tokenize = t
        sys.modules["nltk"] = m
        sys.modules["nltk.tokenize"] = t


_cache = {}
_REF_TOP = ("models", "DAMSM", "train", "datasets")


def load_reference():
    """Returns a namespace with the reference modules: .losses, .config, .sync_batchnorm."""
    if "ns" in _cache:
        return _cache["ns"]
    if not reference_available():
        raise RuntimeError("reference neither mounted at /root/reference nor staged under baseline/_ref")
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
    ns = types.SimpleNamespace(losses=losses, config=config, cfg=config.cfg, sync_batchnorm=syncbn, root=REFERENCE_ROOT)
    _cache["ns"] = ns
    return ns


def _import_top(names):
    import contextlib
    import io
    mods = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with contextlib.redirect_stdout(io.StringIO()):
            for n in names:
                mods[n] = importlib.import_module(n)
    return mods


def load_reference_models():
    """The reference's own models.py / DAMSM.py / train.py on its own sync_batchnorm and losses:
    namespace with .models (Gen, ATTR_Enhance, affine_ssa ...), .DAMSM (CNN_ENCODER, RNN_ENCODER),
    .train (Trainer.DAMSM_loss ... as static methods) on top of load_reference()'s fields."""
    if "models" in _cache:
        return _cache["models"]
    ns = load_reference()
    if not reference_models_available():
        raise RuntimeError("reference models.py / DAMSM.py / train.py not available under %s" % REFERENCE_ROOT)
    _install_stubs(train_too=True)
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        mods = _import_top(("models", "DAMSM", "train"))
    finally:
        sys.argv = argv
    out = types.SimpleNamespace(**vars(ns))
    out.models, out.DAMSM, out.train = mods["models"], mods["DAMSM"], mods["train"]
    _cache["models"] = out
    return out


def load_reference_installed():
    """The SAME reference files imported a second time with ``eegan_b200.install()`` active, i.e. what a user
    gets from the zero-edit route of INTEGRATION.md: ``models.Gen`` builds eegan_b200's SynchronizedBatchNorm2d,
    ``train.Trainer.DAMSM_loss`` calls eegan_b200's words_loss / sent_loss.  The module table is restored
    afterwards, so the un-installed reference (load_reference_models) stays usable next to it."""
    if "installed" in _cache:
        return _cache["installed"]
    ref = load_reference_models()
    import eegan_b200
    keys = ("miscc.DAMSM_losses", "sync_batchnorm", "sync_batchnorm.batchnorm", "sync_batchnorm.replicate",
            "sync_batchnorm.comm") + _REF_TOP
    saved = {k: sys.modules.get(k) for k in keys}
    miscc = sys.modules["miscc"]
    saved_attr = getattr(miscc, "DAMSM_losses", None)
    try:
        for k in keys[1:]:
            sys.modules.pop(k, None)
        eegan_b200.install()
        argv, sys.argv = sys.argv, [sys.argv[0]]
        try:
            mods = _import_top(("models", "DAMSM", "train"))
        finally:
            sys.argv = argv
        out = types.SimpleNamespace(models=mods["models"], DAMSM=mods["DAMSM"], train=mods["train"], cfg=ref.cfg,
                                    losses=sys.modules["miscc.DAMSM_losses"], sync_batchnorm=sys.modules["sync_batchnorm"])
    finally:
        for k in keys:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
        if saved_attr is not None:
            setattr(miscc, "DAMSM_losses", saved_attr)
    assert out.train.words_loss.__module__.startswith("eegan_b200"), "install() did not route train.py's words_loss"
    assert out.models.BatchNorm.__module__.startswith("eegan_b200"), "install() did not route models.py's SyncBN"
    _cache["installed"] = out
    return out

"""Run-time configuration read by the loss functions.

The reference reads the global ``miscc.config.cfg`` at call time (DAMSM_losses.py:14,145,
159,196,211,227).  When this package is used as a drop-in inside the reference tree that
same object is honoured; stand-alone, an object with the reference's defaults
(miscc/config.py:19,47-51) is used.
"""
from __future__ import annotations

import sys
from types import SimpleNamespace

cfg = SimpleNamespace(
    CUDA=True,
    TRAIN=SimpleNamespace(SMOOTH=SimpleNamespace(GAMMA1=5.0, GAMMA2=5.0, GAMMA3=10.0, LAMBDA=1.0)),
    TEXT=SimpleNamespace(EMBEDDING_DIM=256, WORDS_NUM=20),
)


def get_cfg():
    mod = sys.modules.get("miscc.config")
    if mod is not None and hasattr(mod, "cfg"):
        return mod.cfg
    return cfg


def gammas():
    s = get_cfg().TRAIN.SMOOTH
    return float(s.GAMMA1), float(s.GAMMA2), float(s.GAMMA3)

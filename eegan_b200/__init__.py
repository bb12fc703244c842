"""eegan_b200 — B200 (sm_100a) word-region attention + DAMSM losses + SyncBatchNorm for
qikizh/EE-GAN, behind the reference's own Python signatures.

    import eegan_b200
    eegan_b200.install()          # miscc.DAMSM_losses / sync_batchnorm now resolve here
    from miscc.DAMSM_losses import words_loss, sent_loss      # train.py:24, unchanged

Hot path: hand-written CUDA in eegan_b200/csrc -> libeegan_b200.so (C ABI:
include/eegan_b200.h), loaded with ctypes.  No CPU path, no other backend.
"""
from __future__ import annotations

import sys

from . import damsm_losses  # noqa: F401
from .damsm_losses import (GlobalAttentionGeneral, cosine_similarity, func_attention, sent_loss,  # noqa: F401
                           sent_similarity, words_loss, words_similarity)
from .attr_enhance import ATTR_Enhance, attr_enhance  # noqa: F401
from .encoder import EmbFeatures, conv1x1_features, fuse_emb_features  # noqa: F401
from .evaluation import r_precision  # noqa: F401
from .ssa import affine_ssa, fuse_affine_ssa, ssa_modulate  # noqa: F401

__version__ = "0.1.0"


def install(losses: bool = True, sync_batchnorm: bool = True, distributed: bool = True) -> None:
    """Route the reference's import names to this package (drop-in boundary):
    ``miscc.DAMSM_losses`` (train.py:24) and ``sync_batchnorm`` (models.py:8-10, train.py:25).

    ``distributed=True``: under torchrun (one process per GPU, process group initialised) ``words_loss`` /
    ``sent_loss`` called with the rank's shard return the loss of the GLOBAL batch (eegan_b200.sharded) — what
    the reference computes on GPU 0 after nn.DataParallel gathers (train.py:195, 419-435); with a single process
    nothing changes."""
    damsm_losses.AUTO_SHARD = bool(distributed)
    if losses:
        sys.modules["miscc.DAMSM_losses"] = damsm_losses
        parent = sys.modules.get("miscc")
        if parent is not None:
            setattr(parent, "DAMSM_losses", damsm_losses)
    if sync_batchnorm:
        from . import sync_batchnorm as sbn
        sys.modules["sync_batchnorm"] = sbn

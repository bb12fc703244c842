"""Evaluation-side consumers of the matching scores.

``r_precision`` replaces the arithmetic of ``Tester.cal_sim_one_by_one`` (test.py:306-336): the
reference scores one generated image at a time against its R_val = 100 candidate sentence codes
(candidate 0 is the ground-truth caption) with a 1 x 100 ``torch.mm``, two norms, a clamp and a
host-synchronising ``argmax``; here the whole batch is one kernel launch (eegan_rprecision) and
nothing synchronises.  Producing the codes (the DAMSM encoders, the mismatched-caption sampler of
the dataset) stays with the caller, as in the reference.
"""
from __future__ import annotations

import torch

from . import _lib


def r_precision(cnn_code, rnn_codes, eps=1e-8, return_scores=False):
    """cnn_code [B, D]; rnn_codes [B, R_val, D] with the matching sentence at index 0 (test.py:321).

    Returns ``hits`` (bool [B], test.py:329-330: ``R_hits[R_cnt] = 1`` iff argmax == 0) and the
    argmax index per image; with ``return_scores`` also scores0 [B, R_val] (test.py:327).  No
    gradient (the reference runs this under evaluation)."""
    _lib.require_cuda(cnn_code, rnn_codes)
    if cnn_code.dim() != 2 or rnn_codes.dim() != 3 or rnn_codes.shape[0] != cnn_code.shape[0] \
            or rnn_codes.shape[2] != cnn_code.shape[1]:
        raise ValueError("r_precision: expected cnn_code [B,D] and rnn_codes [B,R_val,D], got %s and %s"
                         % (tuple(cnn_code.shape), tuple(rnn_codes.shape)))
    L = _lib.lib()
    cnn = _lib.f32c(cnn_code.detach())
    rnn = _lib.f32c(rnn_codes.detach())
    B, Rv, D = rnn.shape
    best = torch.empty(B, dtype=torch.int32, device=cnn.device)
    hit = torch.empty(B, dtype=torch.uint8, device=cnn.device)
    scores = torch.empty(B, Rv, dtype=torch.float32, device=cnn.device) if return_scores else None
    with torch.cuda.device(cnn.device):
        _lib.check(L.eegan_rprecision(_lib.ptr(cnn), _lib.ptr(rnn), B, Rv, D, float(eps), _lib.ptr(scores),
                                      _lib.ptr(best), _lib.ptr(hit), _lib.stream_ptr()), "rprecision")
    hits = hit.bool()
    return (hits, best, scores) if return_scores else (hits, best)

"""``ATTR_Enhance`` (models.py:146-180) on the B200 library — SURVEY.md §8f rank 2.

Self-attention of the sentence code over its attribute codes: ``combine = [sent ; attrs]`` (1 + attr_num tokens),
``q, k, v = Linear(combine)``, ``a = softmax(q k^T, -1) * 1/sqrt(ntf)`` (the scale is applied AFTER the softmax,
models.py:166 — reproduced), ``attn_attrs = a v``, ``attn_sent = attn_attrs[:, 0]``.  Its outputs feed ``sent_loss``
(the attribute loss, train.py:432) and ``Gen``.  The module keeps the reference's sub-module names
(``attr_query`` / ``attr_key`` / ``attr_value``: ``nn.Linear`` parameter holders), so a reference ``state_dict`` loads
unchanged; the arithmetic runs in three forward and four backward launches of the library (``eegan_attr_enhance_*``):
cat, one batched small GEMM for the three projections, per-sample attention; and their transposes.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib

__all__ = ["ATTR_Enhance", "attr_enhance"]


class _AttrEnhanceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sent, attrs, Wq, bq, Wk, bk, Wv, bv, norm_fact):
        L = _lib.lib()
        B, D = sent.shape
        A = attrs.shape[1]
        Tk = A + 1
        f32 = dict(dtype=torch.float32, device=sent.device)
        out = torch.empty(B, Tk, D, **f32)
        qkv = torch.empty(3, B * Tk, D, **f32)
        p = torch.empty(B, Tk, Tk, **f32)
        combine = torch.empty(B * Tk, D, **f32)
        with torch.cuda.device(sent.device):
            _lib.check(L.eegan_attr_enhance_fwd(_lib.ptr(sent), _lib.ptr(attrs), _lib.ptr(Wq), _lib.ptr(bq), _lib.ptr(Wk),
                                                _lib.ptr(bk), _lib.ptr(Wv), _lib.ptr(bv), B, D, A, norm_fact, _lib.ptr(out),
                                                _lib.ptr(qkv), _lib.ptr(p), _lib.ptr(combine), _lib.stream_ptr()), "attr_enhance_fwd")
        ctx.save_for_backward(combine, qkv, p, Wq, Wk, Wv)
        ctx.norm, ctx.dims = norm_fact, (B, D, A)
        return out

    @staticmethod
    def backward(ctx, d_attrs_out):
        # attn_sent is sliced out of attn_attrs by the caller: its gradient arrives inside d_attrs_out
        combine, qkv, p, Wq, Wk, Wv = ctx.saved_tensors
        L = _lib.lib()
        B, D, A = ctx.dims
        Tk = A + 1
        need = ctx.needs_input_grad
        f32 = dict(dtype=torch.float32, device=combine.device)
        d_attrs_out = _lib.f32c(d_attrs_out)
        g = torch.empty(3, B * Tk, D, **f32)
        dtok = torch.empty(B * Tk, D, **f32)
        d_sent = torch.empty(B, D, **f32) if need[0] else None
        d_attrs = torch.empty(B, A, D, **f32) if need[1] else None
        want_w = need[2] or need[4] or need[6]
        dWq, dWk, dWv = (torch.empty_like(Wq), torch.empty_like(Wk), torch.empty_like(Wv)) if want_w else (None, None, None)
        dbq = torch.empty(D, **f32) if need[3] else None
        dbk = torch.empty(D, **f32) if need[5] else None
        dbv = torch.empty(D, **f32) if need[7] else None
        with torch.cuda.device(combine.device):
            _lib.check(L.eegan_attr_enhance_bwd(None, _lib.ptr(d_attrs_out), _lib.ptr(combine), _lib.ptr(qkv), _lib.ptr(p),
                                                _lib.ptr(Wq), _lib.ptr(Wk), _lib.ptr(Wv), B, D, A, ctx.norm, _lib.ptr(g),
                                                _lib.ptr(dtok), _lib.ptr(d_sent), _lib.ptr(d_attrs), _lib.ptr(dWq), _lib.ptr(dbq),
                                                _lib.ptr(dWk), _lib.ptr(dbk), _lib.ptr(dWv), _lib.ptr(dbv), _lib.stream_ptr()),
                       "attr_enhance_bwd")
        return (d_sent, d_attrs, dWq if need[2] else None, dbq, dWk if need[4] else None, dbk, dWv if need[6] else None, dbv, None)


def attr_enhance(sent, attrs, Wq, bq, Wk, bk, Wv, bv, norm_fact):
    """Functional form: returns (attn_sent [B, D], attn_attrs [B, 1 + attr_num, D])."""
    _lib.require_cuda(sent, attrs, Wq, Wk, Wv)
    c = _lib.f32c
    out = _AttrEnhanceFn.apply(c(sent), c(attrs), c(Wq), c(bq), c(Wk), c(bk), c(Wv), c(bv), float(norm_fact))
    return out[:, 0, :], out  # models.py:168: attn_sent is a view of attn_attrs


class ATTR_Enhance(nn.Module):
    """Drop-in for models.py:146-180 (same constructor, sub-module names, forward signature and return pair)."""

    def __init__(self, ntf=None):
        super().__init__()
        if ntf is None:
            from .config import get_cfg
            ntf = int(get_cfg().TEXT.EMBEDDING_DIM)
        self.attr_query = nn.Linear(ntf, ntf)
        self.attr_key = nn.Linear(ntf, ntf)
        self.attr_value = nn.Linear(ntf, ntf)
        self._norm_fact = 1 / math.sqrt(ntf)

    def forward(self, sent, attrs):
        """sent: bs x ntf;  attrs: bs x attr_num x ntf  ->  (attn_sent bs x ntf, attn_attrs bs x (1 + attr_num) x ntf)."""
        return attr_enhance(sent, attrs, self.attr_query.weight, self.attr_query.bias, self.attr_key.weight, self.attr_key.bias,
                            self.attr_value.weight, self.attr_value.bias, self._norm_fact)

    @staticmethod
    def attr_merge(attn_attrs):
        # models.py:171-180, "method 1": plain sum over the tokens
        return attn_attrs.sum(dim=1)

"""func_attention and cosine_similarity as stand-alone autograd ops
(miscc/DAMSM_losses.py:17-23, 25-63).  Kernels: eegan_func_attention_*, eegan_cosine_rows_*."""
from __future__ import annotations

import torch

from . import _lib


class _FuncAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, context, gamma1):
        L = _lib.lib()
        B, D, T = query.shape
        R = context.shape[2]
        need = L.eegan_func_attention_workspace_bytes(B, D, R, T)
        ws = torch.empty(need, dtype=torch.uint8, device=query.device)
        u = torch.empty(B, D, T, dtype=torch.float32, device=query.device)
        attn = torch.empty(B, T, R, dtype=torch.float32, device=query.device)
        with torch.cuda.device(query.device):
            _lib.check(L.eegan_func_attention_fwd(_lib.ptr(query), _lib.ptr(context), B, D, R, T, gamma1, _lib.ptr(u),
                                                  _lib.ptr(attn), _lib.ptr(ws), need, _lib.stream_ptr()),
                       "func_attention_fwd")
        ctx.save_for_backward(query, context, attn)
        ctx.ws, ctx.g1 = ws, gamma1
        return u, attn

    @staticmethod
    def backward(ctx, d_u, d_attn):
        query, context, attn = ctx.saved_tensors
        if ctx.ws is None:
            raise RuntimeError("eegan_b200.func_attention: backward called twice (its stash is consumed in place)")
        L = _lib.lib()
        B, D, T = query.shape
        R = context.shape[2]
        d_u = _lib.f32c(d_u) if d_u is not None else None
        d_attn = _lib.f32c(d_attn) if d_attn is not None else None
        dq, dc = torch.empty_like(query), torch.empty_like(context)
        with torch.cuda.device(query.device):
            _lib.check(L.eegan_func_attention_bwd(_lib.ptr(query), _lib.ptr(context), _lib.ptr(attn), _lib.ptr(d_u),
                                                  _lib.ptr(d_attn), B, D, R, T, ctx.g1, _lib.ptr(dq), _lib.ptr(dc),
                                                  _lib.ptr(ctx.ws), ctx.ws.numel(), _lib.stream_ptr()),
                       "func_attention_bwd")
        ctx.ws = None
        return dq, dc, None


def func_attention(query, context, gamma1):
    """query [B,D,T], context [B,D,H,W] -> (weightedContext [B,D,T], attn [B,T,H,W])."""
    _lib.require_cuda(query, context)
    B = query.shape[0]
    ih, iw = context.size(2), context.size(3)
    ctxf = _lib.f32c(context).reshape(B, context.shape[1], ih * iw)
    u, attn = _FuncAttentionFn.apply(_lib.f32c(query), ctxf, float(gamma1))
    return u, attn.view(B, -1, ih, iw)


class _CosineRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, eps):
        L = _lib.lib()
        rows, D = x1.shape
        out = torch.empty(rows, dtype=torch.float32, device=x1.device)
        norms = torch.empty(rows, 2, dtype=torch.float32, device=x1.device)
        with torch.cuda.device(x1.device):
            _lib.check(L.eegan_cosine_rows_fwd(_lib.ptr(x1), _lib.ptr(x2), rows, D, eps, _lib.ptr(out),
                                               _lib.ptr(norms), _lib.stream_ptr()), "cosine_rows_fwd")
        ctx.save_for_backward(x1, x2, out, norms)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, out, norms = ctx.saved_tensors
        L = _lib.lib()
        rows, D = x1.shape
        d1, d2 = torch.empty_like(x1), torch.empty_like(x2)
        g = _lib.f32c(g)
        with torch.cuda.device(x1.device):
            _lib.check(L.eegan_cosine_rows_bwd(_lib.ptr(x1), _lib.ptr(x2), _lib.ptr(out), _lib.ptr(norms), _lib.ptr(g),
                                               rows, D, ctx.eps, _lib.ptr(d1), _lib.ptr(d2), _lib.stream_ptr()),
                       "cosine_rows_bwd")
        return d1, d2, None


def cosine_rows(x1, x2, dim=1, eps=1e-8):
    """cosine similarity along ``dim`` followed by ``.squeeze()`` (DAMSM_losses.py:23)."""
    _lib.require_cuda(x1, x2)
    x1, x2 = torch.broadcast_tensors(x1, x2)
    a = _lib.f32c(x1.movedim(dim, -1))
    b = _lib.f32c(x2.movedim(dim, -1))
    lead = a.shape[:-1]
    out = _CosineRowsFn.apply(a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1]), float(eps))
    return out.reshape(lead).squeeze()

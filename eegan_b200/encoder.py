"""``CNN_ENCODER.emb_features`` (DAMSM.py:162, 229) on the B200 library — SURVEY.md §8f rank 1.

The reference projects the 768-channel 17x17 Inception map to ``nef`` = 256 channels with
``conv1x1(768, nef)`` (``nn.Conv2d(kernel_size=1, bias=False)``, DAMSM.py:23-26) and hands the result to
``words_loss`` as ``img_features``.  ``EmbFeatures`` is a drop-in for that layer — same ``weight`` parameter
([nef, 768, 1, 1], so a reference ``state_dict`` loads unchanged) — whose forward and both backward
contractions run as batched fp32-accurate 3xTF32 GEMMs on the tensor cores (``eegan_conv1x1_fwd/_bwd``).
``fuse_emb_features(cnn_encoder)`` swaps the layer of an existing ``CNN_ENCODER`` in place.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib

__all__ = ["EmbFeatures", "conv1x1_features", "fuse_emb_features"]


class _Conv1x1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        L = _lib.lib()
        B, Cin, H, W = x.shape
        Cout, R = w.shape[0], H * W
        y = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
        nb = L.eegan_conv1x1_workspace_bytes(B, Cin, Cout, R, 0)
        xp = torch.empty(nb, dtype=torch.uint8, device=x.device)  # x re-pitched for TMA: the stash of the backward
        with torch.cuda.device(x.device):
            _lib.check(L.eegan_conv1x1_fwd(_lib.ptr(x), _lib.ptr(w), B, Cin, Cout, R, _lib.ptr(y), _lib.ptr(xp), nb,
                                           _lib.stream_ptr()), "conv1x1_fwd")
        ctx.save_for_backward(xp, w)
        ctx.dims = (B, Cin, Cout, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, w = ctx.saved_tensors
        L = _lib.lib()
        B, Cin, Cout, H, W = ctx.dims
        R = H * W
        dy = _lib.f32c(dy)
        dx = torch.empty(B, Cin, H, W, dtype=torch.float32, device=dy.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        nb = L.eegan_conv1x1_workspace_bytes(B, Cin, Cout, R, 1)
        ws = torch.empty(nb, dtype=torch.uint8, device=dy.device)
        with torch.cuda.device(dy.device):
            _lib.check(L.eegan_conv1x1_bwd(_lib.ptr(xp), _lib.ptr(w), _lib.ptr(dy), B, Cin, Cout, R, _lib.ptr(dx), _lib.ptr(dw),
                                           _lib.ptr(ws), nb, _lib.stream_ptr()), "conv1x1_bwd")
        return dx, dw


def conv1x1_features(x, weight):
    """``F.conv2d(x, weight)`` for a 1x1 kernel without bias: x [B, Cin, H, W], weight [Cout, Cin(, 1, 1)]."""
    _lib.require_cuda(x, weight)
    if x.dim() != 4:
        raise ValueError("conv1x1_features expects x of shape [B, Cin, H, W]")
    w = _lib.f32c(weight).reshape(weight.shape[0], -1)
    if w.shape[1] != x.shape[1]:
        raise ValueError("conv1x1_features: weight has %d input channels, x has %d" % (w.shape[1], x.shape[1]))
    return _Conv1x1Fn.apply(_lib.f32c(x), w)


class EmbFeatures(nn.Module):
    """Drop-in for ``conv1x1(in_planes, out_planes)`` of DAMSM.py:23-26 (``CNN_ENCODER.emb_features``, :162)."""

    def __init__(self, in_planes=768, out_planes=256):
        super().__init__()
        self.in_channels, self.out_channels = in_planes, out_planes
        self.weight = nn.Parameter(torch.empty(out_planes, in_planes, 1, 1))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))  # nn.Conv2d's default initialisation

    def forward(self, x):
        return conv1x1_features(x, self.weight)

    def extra_repr(self):
        return "%d, %d, kernel_size=(1, 1), stride=(1, 1), bias=False" % (self.in_channels, self.out_channels)


def fuse_emb_features(cnn_encoder):
    """Replace ``cnn_encoder.emb_features`` (an ``nn.Conv2d`` 1x1 without bias) by an ``EmbFeatures`` sharing its
    weight tensor; returns the encoder.  ``init_trainable_weights`` (DAMSM.py:165-168) keeps working."""
    conv = cnn_encoder.emb_features
    if isinstance(conv, EmbFeatures):
        return cnn_encoder
    if not isinstance(conv, nn.Conv2d) or conv.kernel_size != (1, 1) or conv.bias is not None:
        raise TypeError("emb_features is expected to be conv1x1(…, bias=False)")
    m = EmbFeatures(conv.in_channels, conv.out_channels)
    m.weight = conv.weight
    cnn_encoder.emb_features = m
    return cnn_encoder

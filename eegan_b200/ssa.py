"""affine_ssa — the spatially-gated conditional affine of the generator (models.py:43-86), SURVEY.md §8f rank 3.

``feat -> SyncBN(affine=False) -> (gamma(cond) * mask + 1) * xhat + beta(cond) * mask``: 14 of the 24
SyncBN layers of ``Gen`` feed this modulation.  The reference normalises (three passes over the
activations), expands gamma / beta to the full tensor and runs four more elementwise passes; here
the modulation rides on the normalise pass (eegan_ssa_apply), and the backward is the two passes
batch norm needs anyway — a reduction that also yields d_gamma, d_beta and d_mask
(eegan_ssa_bwd_reduce), the cross-replica all-reduce of 2C floats, and one apply pass
(eegan_ssa_bwd_apply).  The two small MLPs that produce gamma / beta from the condition stay
``nn.Linear`` (cuBLAS) exactly as in the reference; parameter names match, so a reference
``state_dict`` loads unchanged.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .sync_batchnorm.batchnorm import CudaBNOps, SynchronizedBatchNorm2d, _group_size, _stats_with_count


class CudaSSAOps:
    """Device steps of the fused modulation (the product path; tests/ may swap a CPU stand-in)."""

    @staticmethod
    def apply(x3, mean, inv_std, gamma, beta, mask2, y):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_ssa_apply(_lib.ptr(x3), _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(gamma), _lib.ptr(beta),
                                                  _lib.ptr(mask2), N, C, HW, _lib.ptr(y), _lib.stream_ptr()), "ssa_apply")

    @staticmethod
    def fwd_fused(x3, gamma, beta, mask2, eps, momentum, rm, rv, y, work):
        """single replica: statistics + modulate in one library call (one launch for small maps); work [4C] keeps mean / inv_std"""
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_ssa_fwd_fused(_lib.ptr(x3), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mask2), N, C, HW, eps,
                                                      momentum, _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(y), _lib.ptr(work),
                                                      _lib.stream_ptr()), "ssa_fwd_fused")

    @staticmethod
    def bwd_reduce(x3, dy3, mean, inv_std, gamma, beta, mask2, red, dgamma, dbeta, dmask):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_ssa_bwd_reduce(_lib.ptr(x3), _lib.ptr(dy3), _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(gamma),
                                                       _lib.ptr(beta), _lib.ptr(mask2), N, C, HW, _lib.ptr(red), _lib.ptr(dgamma),
                                                       _lib.ptr(dbeta), _lib.ptr(dmask), _lib.stream_ptr()), "ssa_bwd_reduce")

    @staticmethod
    def bwd_apply(x3, dy3, mean, inv_std, gamma, mask2, red, count, count_dev, eps, clamp_mode, dx):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_ssa_bwd_apply(_lib.ptr(x3), _lib.ptr(dy3), _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(gamma),
                                                      _lib.ptr(mask2), _lib.ptr(red), float(count), _lib.ptr(count_dev), eps,
                                                      clamp_mode, N, C, HW, _lib.ptr(dx), _lib.stream_ptr()), "ssa_bwd_apply")


class _SSAFn(torch.autograd.Function):
    """training=True: batch statistics (all-reduced over ``group``), running stats updated;
    training=False: ``mean`` / ``inv_std`` are the given constants (eval, batchnorm.py:50-53)."""

    @staticmethod
    def forward(ctx, x3, gamma, beta, mask2, running_mean, running_var, eps, momentum, group, training, bn_ops, ops):
        N, C, HW = x3.shape
        world = _group_size(group) if training else 1
        if training and world == 1 and hasattr(ops, "fwd_fused") and bn_ops is CudaBNOps:
            work = torch.empty(4 * C, dtype=torch.float32, device=x3.device)
            y = torch.empty_like(x3)
            ops.fwd_fused(x3, gamma, beta, mask2, eps, momentum, running_mean, running_var, y, work)
            ctx.save_for_backward(x3, gamma, beta, mask2, work[2 * C:3 * C], work[3 * C:])
            ctx.count_dev = None
            ctx.cfg = (N * HW, eps, 0, group, ops, 1, training)
            return y
        if training:
            buf = torch.empty(2 * C + 2, dtype=torch.float32, device=x3.device)
            local = N * HW
            if world > 1:
                _stats_with_count(bn_ops, x3, buf, local)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
                count, count_dev = 0, buf[2 * C:]
            else:
                bn_ops.stats(x3, buf)
                count, count_dev = local, None
            clamp_mode = 1 if world > 1 else 0  # batchnorm.py:125 vs :50-53
            mean = torch.empty(C, dtype=torch.float32, device=x3.device)
            inv_std = torch.empty_like(mean)
            bn_ops.finalize(buf, C, count, count_dev, eps, momentum, clamp_mode, mean, inv_std, running_mean, running_var)
        else:
            count, count_dev, clamp_mode = 0, None, 0
            mean = running_mean.contiguous()
            inv_std = torch.rsqrt(running_var + eps)
        y = torch.empty_like(x3)
        ops.apply(x3, mean, inv_std, gamma, beta, mask2, y)
        ctx.save_for_backward(x3, gamma, beta, mask2, mean, inv_std)
        ctx.count_dev = count_dev
        ctx.cfg = (count, eps, clamp_mode, group, ops, world, training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x3, gamma, beta, mask2, mean, inv_std = ctx.saved_tensors
        count, eps, clamp_mode, group, ops, world, training = ctx.cfg
        N, C, HW = x3.shape
        dy = dy.contiguous()
        red = torch.empty(2 * C, dtype=torch.float32, device=x3.device)
        dgamma = torch.empty(N, C, dtype=torch.float32, device=x3.device)
        dbeta = torch.empty_like(dgamma)
        dmask = torch.empty(N, HW, dtype=torch.float32, device=x3.device)
        ops.bwd_reduce(x3, dy, mean, inv_std, gamma, beta, mask2, red, dgamma, dbeta, dmask)
        if training and world > 1:
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
        dx = torch.empty_like(x3)
        ops.bwd_apply(x3, dy, mean, inv_std, gamma, mask2, red if training else None, count, ctx.count_dev, eps,
                      clamp_mode, dx)
        return dx, dgamma, dbeta, dmask, None, None, None, None, None, None, None, None


def ssa_modulate(feat, weight, bias, semi_mask, norm, ops=CudaSSAOps):
    """``(weight * mask + 1) * norm(feat) + bias * mask`` (models.py:69, 80-86) in fused kernels.

    feat [N,C,H,W]; weight, bias [N,C] (or [C]); semi_mask [N,1,H,W]; ``norm`` an affine-free
    SynchronizedBatchNorm2d that owns eps / momentum / running statistics / process group."""
    if ops is CudaSSAOps:
        _lib.require_cuda(feat)
    N, C = feat.shape[0], feat.shape[1]
    if weight.dim() == 1:  # models.py:73-76
        weight = weight.unsqueeze(0)
    if bias.dim() == 1:
        bias = bias.unsqueeze(0)
    x3 = feat.contiguous().float().reshape(N, C, -1)
    HW = x3.shape[2]
    gamma = weight.float().expand(N, C).contiguous()
    beta = bias.float().expand(N, C).contiguous()
    if semi_mask.numel() != N * HW:
        raise ValueError("affine_ssa: semi_mask must be [N,1,H,W] matching feat (got %s for feat %s)"
                         % (tuple(semi_mask.shape), tuple(feat.shape)))
    mask2 = semi_mask.contiguous().float().reshape(N, HW)
    y = _SSAFn.apply(x3, gamma, beta, mask2, norm.running_mean, norm.running_var, norm.eps, norm.momentum,
                     norm.process_group, norm.training, norm._ops, ops)
    return y.view(feat.shape)


class affine_ssa(nn.Module):
    """Drop-in for models.py:43-86 (same constructor, submodule and parameter names, zero-initialised
    second linears :63-66)."""

    def __init__(self, num_features, ntf=256, norm_layer=SynchronizedBatchNorm2d):
        super().__init__()
        self.norm2d = norm_layer(num_features, affine=False)
        self.fc_gamma = nn.Sequential(OrderedDict([
            ("linear1", nn.Linear(ntf, 256)),
            ("relu1", nn.ReLU(inplace=True)),
            ("linear2", nn.Linear(256, num_features)),
        ]))
        self.fc_beta = nn.Sequential(OrderedDict([
            ("linear1", nn.Linear(ntf, 256)),
            ("relu1", nn.ReLU(inplace=True)),
            ("linear2", nn.Linear(256, num_features)),
        ]))
        self._initialize()

    def _initialize(self):
        nn.init.zeros_(self.fc_gamma.linear2.weight.data)
        nn.init.zeros_(self.fc_gamma.linear2.bias.data)
        nn.init.zeros_(self.fc_beta.linear2.weight.data)
        nn.init.zeros_(self.fc_beta.linear2.bias.data)

    def forward(self, feat, cond, semi_mask):
        weight = self.fc_gamma(cond)
        bias = self.fc_beta(cond)
        return ssa_modulate(feat, weight, bias, semi_mask, self.norm2d)


def fuse_affine_ssa(root: nn.Module) -> int:
    """Swap every reference ``affine_ssa`` (models.py:43-86) inside ``root`` — ``Gen`` holds 14 — for the fused module above,
    in place.  The new module ADOPTS the old one's sub-modules (``norm2d``, ``fc_gamma``, ``fc_beta``): same parameter and buffer
    tensors, same ``state_dict`` keys, optimiser state stays valid.  A ``norm2d`` that is the reference's own SyncBN (no
    ``eegan_b200.install()``) is re-housed in eegan_b200's class on the same running-statistics buffers.  Returns the count.

        netG = models.Gen(ngf, nz)           # the reference's generator, unchanged
        eegan_b200.fuse_affine_ssa(netG)     # 14 x (SyncBN + 6 elementwise passes) -> 14 x (stats pass + 1 apply pass)
    """
    n = 0
    for name, child in list(root.named_children()):
        if type(child).__name__ == "affine_ssa" and not isinstance(child, affine_ssa):
            new = affine_ssa.__new__(affine_ssa)
            nn.Module.__init__(new)
            norm = child.norm2d
            if not isinstance(norm, SynchronizedBatchNorm2d):
                if getattr(norm, "affine", False) or not hasattr(norm, "running_mean"):
                    raise TypeError("fuse_affine_ssa: %s.norm2d is not an affine-free batch norm" % name)
                mine = SynchronizedBatchNorm2d(norm.num_features, eps=norm.eps, momentum=norm.momentum, affine=False)
                mine._buffers["running_mean"] = norm.running_mean
                mine._buffers["running_var"] = norm.running_var
                if getattr(norm, "num_batches_tracked", None) is not None:
                    mine._buffers["num_batches_tracked"] = norm.num_batches_tracked
                norm = mine
            new.norm2d, new.fc_gamma, new.fc_beta = norm, child.fc_gamma, child.fc_beta
            new.train(child.training)
            setattr(root, name, new)
            n += 1
        else:
            n += fuse_affine_ssa(child)
    return n

"""SynchronizedBatchNorm{1,2,3}d — sync_batchnorm/batchnorm.py:37-315 of the reference.

B200 design: one process per GPU.  Each rank makes ONE pass over its activations for
[sum, square-sum] (eegan_syncbn_stats), the 2C+2 floats are all-reduced over NCCL on the
process group shared with the sharded DAMSM losses, then mean / inv_std follow the
reference's N-replica formula ``inv_std = clamp(var, eps) ** -0.5`` (batchnorm.py:113-125)
and one more pass normalises.  With a single replica (or in eval mode) the reference calls
F.batch_norm (batchnorm.py:50-53), i.e. ``1/sqrt(var + eps)``; the same kernels are used
with that formula.  The backward is two passes: [sum dy, sum dy*xhat] -> all-reduce -> dx.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch.nn.modules.batchnorm import _BatchNorm

from .. import _lib

__all__ = ["SynchronizedBatchNorm1d", "SynchronizedBatchNorm2d", "SynchronizedBatchNorm3d"]

_COUNT_SPLIT = 4096  # element counts travel as two fp32 so that their all-reduce stays exact


class CudaBNOps:
    """Device steps of SyncBN (the product path).  tests/ swap in a CPU stand-in built on the
    oracle to exercise the host / collective logic with gloo."""

    @staticmethod
    def stats(x3, out):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_stats(_lib.ptr(x3), N, C, HW, _lib.ptr(out), _lib.stream_ptr()), "syncbn_stats")

    @staticmethod
    def stats_counted(x3, out):
        """stats + the local element count as the exact pair {n // 4096, n % 4096} in out[2C:2C+2] (one launch more, no host writes)."""
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_stats_counted(_lib.ptr(x3), N, C, HW, _lib.ptr(out), _lib.stream_ptr()), "syncbn_stats_counted")

    @staticmethod
    def fwd_fused(x3, w, b, eps, momentum, rm, rv, y, work):
        """single replica: statistics + normalise in one library call (one launch for small maps); work [4C] keeps mean / inv_std"""
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_fwd_fused(_lib.ptr(x3), _lib.ptr(w), _lib.ptr(b), N, C, HW, eps, momentum, _lib.ptr(rm),
                                                         _lib.ptr(rv), _lib.ptr(y), _lib.ptr(work), _lib.stream_ptr()), "syncbn_fwd_fused")

    @staticmethod
    def bwd_fused(x3, dy3, work, w, eps, dx, red):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_bwd_fused(_lib.ptr(x3), _lib.ptr(dy3), _lib.ptr(work), _lib.ptr(w), N, C, HW, eps,
                                                         _lib.ptr(dx), _lib.ptr(red), _lib.stream_ptr()), "syncbn_bwd_fused")

    @staticmethod
    def finalize(stats, C, count, count_dev, eps, momentum, clamp_mode, mean, inv_std, rm, rv):
        with torch.cuda.device(stats.device):
            _lib.check(_lib.lib().eegan_syncbn_finalize(_lib.ptr(stats), C, float(count), _lib.ptr(count_dev), eps,
                                                        momentum, clamp_mode,
                                                        _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(rm), _lib.ptr(rv),
                                                        _lib.stream_ptr()), "syncbn_finalize")

    @staticmethod
    def apply(x3, mean, inv_std, w, b, y):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_apply(_lib.ptr(x3), _lib.ptr(mean), _lib.ptr(inv_std), _lib.ptr(w),
                                                     _lib.ptr(b), N, C, HW, _lib.ptr(y), _lib.stream_ptr()), "syncbn_apply")

    @staticmethod
    def bwd_reduce(x3, dy3, mean, inv_std, red):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_bwd_reduce(_lib.ptr(x3), _lib.ptr(dy3), _lib.ptr(mean), _lib.ptr(inv_std),
                                                          N, C, HW, _lib.ptr(red), _lib.stream_ptr()), "syncbn_bwd_reduce")

    @staticmethod
    def bwd_apply(x3, dy3, mean, inv_std, w, red, count, count_dev, eps, clamp_mode, dx):
        N, C, HW = x3.shape
        with torch.cuda.device(x3.device):
            _lib.check(_lib.lib().eegan_syncbn_bwd_apply(_lib.ptr(x3), _lib.ptr(dy3), _lib.ptr(mean), _lib.ptr(inv_std),
                                                         _lib.ptr(w), _lib.ptr(red), float(count), _lib.ptr(count_dev),
                                                         eps, clamp_mode,
                                                         N, C, HW, _lib.ptr(dx), _lib.stream_ptr()), "syncbn_bwd_apply")


def _stats_with_count(ops, x3, buf, local):
    """[sum, square-sum] into buf[:2C] and the exact count pair into buf[2C:] (device-side when the ops provide it)."""
    if hasattr(ops, "stats_counted"):
        ops.stats_counted(x3, buf)
    else:  # CPU stand-ins of the tests
        C = x3.shape[1]
        ops.stats(x3, buf)
        buf[2 * C] = float(local // _COUNT_SPLIT)
        buf[2 * C + 1] = float(local % _COUNT_SPLIT)


def _group_size(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class _SyncBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x3, weight, bias, running_mean, running_var, eps, momentum, group, ops):
        N, C, HW = x3.shape
        world = _group_size(group)
        if world == 1 and hasattr(ops, "fwd_fused"):  # one replica: nothing to wait for between statistics and normalisation
            work = torch.empty(4 * C, dtype=torch.float32, device=x3.device)
            y = torch.empty_like(x3)
            ops.fwd_fused(x3, weight, bias, eps, momentum, running_mean, running_var, y, work)
            ctx.save_for_backward(x3, weight, work)
            ctx.fused = True
            ctx.cfg = (N * HW, eps, 0, group, ops, 1, bias is not None)
            return y
        ctx.fused = False
        buf = torch.empty(2 * C + 2, dtype=torch.float32, device=x3.device)
        local = N * HW
        if world > 1:
            _stats_with_count(ops, x3, buf, local)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)  # batchnorm.py:102 (ReduceAddCoalesced)
            count, count_dev = 0, buf[2 * C:]  # the total stays on the device: no host sync per layer
        else:
            ops.stats(x3, buf)  # fills [0, 2C)
            count, count_dev = local, None
        clamp_mode = 1 if world > 1 else 0  # batchnorm.py:125 vs :50-53
        mean = torch.empty(C, dtype=torch.float32, device=x3.device)
        inv_std = torch.empty_like(mean)
        ops.finalize(buf, C, count, count_dev, eps, momentum, clamp_mode, mean, inv_std, running_mean, running_var)
        y = torch.empty_like(x3)
        ops.apply(x3, mean, inv_std, weight, bias, y)
        ctx.save_for_backward(x3, weight, mean, inv_std)
        ctx.count_dev = count_dev
        ctx.cfg = (count, eps, clamp_mode, group, ops, world, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.fused:
            x3, weight, work = ctx.saved_tensors
            count, eps, _, _, ops, _, has_bias = ctx.cfg
            C = x3.shape[1]
            red = torch.empty(2 * C, dtype=torch.float32, device=x3.device)
            dx = torch.empty_like(x3)
            ops.bwd_fused(x3, dy.contiguous(), work, weight, eps, dx, red)
            return dx, (red[C:] if weight is not None else None), (red[:C] if has_bias else None), None, None, None, None, None, None
        x3, weight, mean, inv_std = ctx.saved_tensors
        count, eps, clamp_mode, group, ops, world, has_bias = ctx.cfg
        N, C, HW = x3.shape
        dy = dy.contiguous()
        red = torch.empty(2 * C, dtype=torch.float32, device=x3.device)
        ops.bwd_reduce(x3, dy, mean, inv_std, red)
        d_w = red[C:].clone() if weight is not None else None  # local sums: parameter grads stay per replica
        d_b = red[:C].clone() if has_bias else None
        if world > 1:
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
        dx = torch.empty_like(x3)
        ops.bwd_apply(x3, dy, mean, inv_std, weight, red, count, ctx.count_dev, eps, clamp_mode, dx)
        return dx, d_w, d_b, None, None, None, None, None, None


class _BNEvalFn(torch.autograd.Function):
    """Eval mode (batchnorm.py:50-53 -> F.batch_norm with the running statistics), differentiable like the
    reference's: the statistics are constants, so dx = dy * w * inv_std, d_w = sum dy * xhat, d_b = sum dy."""

    @staticmethod
    def forward(ctx, x3, weight, bias, mean, inv_std, ops):
        y = torch.empty_like(x3)
        ops.apply(x3, mean, inv_std, weight, bias, y)
        ctx.save_for_backward(x3, weight, mean, inv_std)
        ctx.ops, ctx.has_bias = ops, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x3, weight, mean, inv_std = ctx.saved_tensors
        ops = ctx.ops
        N, C, HW = x3.shape
        dy = dy.contiguous()
        d_w = d_b = dx = None
        if (weight is not None and ctx.needs_input_grad[1]) or (ctx.has_bias and ctx.needs_input_grad[2]):
            red = torch.empty(2 * C, dtype=torch.float32, device=x3.device)
            ops.bwd_reduce(x3, dy, mean, inv_std, red)
            d_w = red[C:] if weight is not None else None
            d_b = red[:C] if ctx.has_bias else None
        if ctx.needs_input_grad[0]:
            zero = torch.zeros(2 * C, dtype=torch.float32, device=x3.device)  # no batch-statistics terms in eval mode
            dx = torch.empty_like(x3)
            ops.bwd_apply(x3, dy, mean, inv_std, weight, zero, 1, None, 0.0, 0, dx)
        return dx, d_w, d_b, None, None, None


class _SynchronizedBatchNorm(_BatchNorm):
    """batchnorm.py:37-125.  ``process_group`` (default: the world group) replaces the
    reference's SyncMaster / SlavePipe plumbing."""

    _ops = CudaBNOps

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, process_group=None):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self.process_group = process_group

    def forward(self, input):
        self._check_input_dim(input)
        if self._ops is CudaBNOps:
            _lib.require_cuda(input)
        shape = input.shape
        x3 = input.contiguous().float().reshape(shape[0], self.num_features, -1)
        if not self.training:  # batchnorm.py:50-53 eval branch: running statistics
            inv_std = torch.rsqrt(self.running_var + self.eps)
            y = _BNEvalFn.apply(x3, self.weight, self.bias, self.running_mean.contiguous(), inv_std, self._ops)
            return y.view(shape)
        y = _SyncBNFn.apply(x3, self.weight, self.bias, self.running_mean, self.running_var, self.eps,
                            self.momentum, self.process_group, self._ops)
        return y.view(shape)

    def __data_parallel_replicate__(self, ctx, copy_id):
        # kept for API compatibility (batchnorm.py:80-88); replicas are processes here
        pass


class SynchronizedBatchNorm1d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 2 and input.dim() != 3:
            raise ValueError("expected 2D or 3D input (got {}D input)".format(input.dim()))


class SynchronizedBatchNorm2d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 4:
            raise ValueError("expected 4D input (got {}D input)".format(input.dim()))


class SynchronizedBatchNorm3d(_SynchronizedBatchNorm):
    def _check_input_dim(self, input):
        if input.dim() != 5:
            raise ValueError("expected 5D input (got {}D input)".format(input.dim()))

"""Caption-row-sharded DAMSM losses for one process per GPU.

The reference computes the full-batch B x B grid on GPU 0 after nn.DataParallel gathers the
generator's output (train.py:195, 419-435).  Here rank g owns captions [g*b, (g+1)*b) and
the same slice of images:

  fwd: all-gather region features -> every rank holds [B, D, R]; the rank computes its column
       block m[:, caps_g] (B x b pairs); all-gather of the blocks -> the full grid everywhere;
       the O(B^2) cross-entropy tail is computed redundantly on every rank.
  bwd: dL/dm is known everywhere; the rank computes complete d_words for its captions and a
       partial d_img [B, D, R] from its columns; reduce-scatter(sum) -> d_img of its own images.

The result equals the single-device full-batch loss.  Collectives go through
torch.distributed (NCCL over NVLink/NVSwitch on the GPU box; gloo in the CPU tests, where
tests/ inject an oracle-based ``grid_fn`` to exercise this host logic without a GPU).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .config import gammas


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


class _AllGatherRows(torch.autograd.Function):
    """all-gather along dim 0.  backward: ``partial=True`` — every rank holds a partial
    gradient of the gathered tensor -> reduce-scatter(sum); ``partial=False`` — every rank
    holds the same complete gradient -> keep the own rows (no collective)."""

    @staticmethod
    def forward(ctx, x, group, partial):
        world, rank = _world(group)
        ctx.group, ctx.world, ctx.rank, ctx.partial = group, world, rank, partial
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n = g.shape[0] // ctx.world
        if not ctx.partial:
            return g[ctx.rank * n:(ctx.rank + 1) * n].contiguous(), None, None
        out = torch.empty((n,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None, None


class _AllGatherCols(torch.autograd.Function):
    """[B, b] column blocks -> [B, B].  Every rank evaluates the same loss on the full grid, so
    the gradient of its own block is a slice of the local gradient (no collective)."""

    @staticmethod
    def forward(ctx, block, group):
        world, rank = _world(group)
        ctx.rank, ctx.b = rank, block.shape[1]
        parts = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(parts, block.contiguous(), group=group)  # concatenated along dim 0
        parts = parts.view(world, block.shape[0], block.shape[1])
        return parts.permute(1, 0, 2).reshape(block.shape[0], world * block.shape[1]).contiguous()

    @staticmethod
    def backward(ctx, g):
        return g[:, ctx.rank * ctx.b:(ctx.rank + 1) * ctx.b].contiguous(), None


def _gather_ids(ids, device, group):
    if ids is None:
        return None
    t = torch.as_tensor(ids).reshape(-1).to(device=device, dtype=torch.int64)
    world, _ = _world(group)
    out = torch.empty(world * t.shape[0], dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def sharded_words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size, group=None,
                       grid_fn=None, ce_fn=None):
    """words_loss (DAMSM_losses.py:272-342) over the GLOBAL batch, from per-rank shards.

    Arguments are the rank's local shard with the reference's meaning (``batch_size`` is the
    local batch b; ``labels`` only selects loss/no-loss — the global labels are arange(B) as in
    train.py:93).  Returns (loss0, loss1, att_maps of the local captions).
    """
    from . import damsm_losses as dl
    grid_fn = grid_fn or dl.pair_grid
    world, rank = _world(group)
    if world == 1:
        return dl.words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size)
    b = batch_size
    img_all = _AllGatherRows.apply(img_features[:b].contiguous(), group, True)
    m_block, att = grid_fn(img_all, words_emb[:b], cap_lens, diag_offset=rank * b)
    att_maps = dl._LazyAttMaps(att, cap_lens, dl._spatial(img_features))
    if labels is None:
        return None, None, att_maps
    m_all = _AllGatherCols.apply(m_block, group)
    cls_all = _gather_ids(class_ids, m_all.device, group)
    lab_all = torch.arange(world * b, device=m_all.device, dtype=torch.int64)
    _, _, g3 = gammas()
    if ce_fn is None:
        loss0, loss1, _ = dl._PairCEFn.apply(m_all, g3, cls_all, lab_all)
    else:
        loss0, loss1 = ce_fn(m_all, g3, cls_all, lab_all)
    return loss0, loss1, att_maps


def sharded_sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, group=None, loss_fn=None):
    """sent_loss (DAMSM_losses.py:233-270) over the global batch: all-gather both codes
    (B x 1 KB) and evaluate redundantly; every rank then holds the complete gradient and keeps
    the rows of its own samples."""
    from . import damsm_losses as dl
    world, _ = _world(group)
    if world == 1:
        return (loss_fn or dl.sent_loss)(cnn_code, rnn_code, labels, class_ids, batch_size, eps)
    if labels is None:
        return None, None
    b = batch_size
    cnn_all = _AllGatherRows.apply(cnn_code[:b].contiguous(), group, False)
    rnn_all = _AllGatherRows.apply(rnn_code[:b].contiguous(), group, False)
    cls_all = _gather_ids(class_ids, cnn_all.device, group)
    lab_all = torch.arange(world * b, device=cnn_all.device, dtype=torch.int64)
    return (loss_fn or dl.sent_loss)(cnn_all, rnn_all, lab_all, cls_all, world * b, eps)

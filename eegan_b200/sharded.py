"""Sharded DAMSM losses for one process per GPU.

The reference computes the full-batch B x B grid m[image j][caption i] on GPU 0 after nn.DataParallel gathers the
generator's output (train.py:195, 419-435).  Here rank g owns samples [g*b, (g+1)*b) and evaluates one block of the grid.
Two partitions of the same grid, same kernels, same result (``shard=`` / ``EEGAN_SHARD_BY``):

``"images"`` (default): the rank's IMAGES against all captions — the block m[imgs_g, :] (b x B pairs).
  fwd: all-gather the WORD features and caption lengths (B x D x T: 18 KB per caption); block; all-gather of the row blocks ->
       the full grid everywhere; the O(B^2) cross-entropy tail redundantly on every rank.
  bwd: d_img of the own images is complete locally; the rank holds a partial d_words [B, D, T] -> reduce-scatter(sum).
  Per step and rank this moves 2 x B x 18 KB; the caption partition below moves 2 x B x 296 KB (the region features are 16x
  the word features): 7 MB against 114 MB per collective at B = 384.

``"captions"`` (the partition BASELINE.json's north_star describes): the rank's CAPTIONS against all images, m[:, caps_g]:

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

import os

from .config import gammas

SHARD_BY = os.environ.get("EEGAN_SHARD_BY", "images")  # "images" | "captions": default partition of the pair grid (read once)


def _shard_mode(shard):
    mode = SHARD_BY if shard is None else shard
    if mode not in ("images", "captions"):
        raise ValueError("shard must be 'images' or 'captions', got %r" % (mode,))
    return mode


class _no_auto_shard:
    """Marks the calling thread as being inside a sharded evaluation, so that damsm_losses' AUTO_SHARD dispatch
    (``eegan_b200.install(distributed=True)``) does not shard an already-gathered batch a second time."""

    def __init__(self, dl):
        self._tls = dl._tls

    def __enter__(self):
        self._prev = getattr(self._tls, "inside_sharded", False)
        self._tls.inside_sharded = True

    def __exit__(self, *a):
        self._tls.inside_sharded = self._prev


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
                       grid_fn=None, ce_fn=None, shard=None):
    """words_loss (DAMSM_losses.py:272-342) over the GLOBAL batch, from per-rank shards.

    Arguments are the rank's local shard with the reference's meaning (``batch_size`` is the
    local batch b; ``labels`` only selects loss/no-loss — the global labels are arange(B) as in
    train.py:93).  Returns (loss0, loss1, att_maps of the local captions).
    """
    from . import damsm_losses as dl
    grid_fn = grid_fn or dl.pair_grid
    world, rank = _world(group)
    if world == 1:
        with _no_auto_shard(dl):
            return dl.words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size)
    b = batch_size
    if _shard_mode(shard) == "images":
        # own images x all captions: the word features travel (partial d_words comes back through the reduce-scatter of the
        # gather's backward), d_img is complete locally
        words_all = _AllGatherRows.apply(words_emb[:b].contiguous(), group, True)
        lens_all = _gather_ids(torch.as_tensor(cap_lens).reshape(-1)[:b], words_all.device, group)
        m_block, att_all = grid_fn(img_features[:b], words_all, lens_all, diag_offset=-rank * b)  # [b, B], [B, T, R]
        att = att_all[rank * b:(rank + 1) * b] if att_all is not None else None
        att_maps = dl._att_maps(att, cap_lens, dl._spatial(img_features))
        if labels is None:
            return None, None, att_maps
        m_all = _AllGatherRows.apply(m_block, group, False)  # row blocks; every rank gets the complete gradient, keeps its rows
    else:
        img_all = _AllGatherRows.apply(img_features[:b].contiguous(), group, True)
        m_block, att = grid_fn(img_all, words_emb[:b], cap_lens, diag_offset=rank * b)
        att_maps = dl._att_maps(att, cap_lens, dl._spatial(img_features))
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
        with _no_auto_shard(dl):
            return (loss_fn or dl.sent_loss)(cnn_code, rnn_code, labels, class_ids, batch_size, eps)
    if labels is None:
        return None, None
    b = batch_size
    cnn_all = _AllGatherRows.apply(cnn_code[:b].contiguous(), group, False)
    rnn_all = _AllGatherRows.apply(rnn_code[:b].contiguous(), group, False)
    cls_all = _gather_ids(class_ids, cnn_all.device, group)
    lab_all = torch.arange(world * b, device=cnn_all.device, dtype=torch.int64)
    with _no_auto_shard(dl):  # the gathered codes are the full batch: evaluate it as such, even under install(distributed=True)
        return (loss_fn or dl.sent_loss)(cnn_all, rnn_all, lab_all, cls_all, world * b, eps)


class ShardedWordsLossStep:
    """words_loss forward + backward over the GLOBAL batch from per-rank shards, without autograd: the same
    collectives and kernels as ``sharded_words_loss(...)`` followed by ``(w0 * loss0 + w1 * loss1).backward()``,
    enqueued directly (C ABI + torch.distributed) on static buffers.  At CUB sizes the autograd route spends ~1 ms
    of host time per step on bookkeeping, several times what the GPU needs; this route costs the host four library
    calls, four collectives and two small copies.

        step = ShardedWordsLossStep(b, 256, 17, 17, T_max, device)        # b = local batch
        loss0, loss1, d_img, d_words = step(img, words, cap_lens, class_ids)   # static tensors, valid until the next call

    ``graph=True`` captures the step (collectives included) into one CUDA graph."""

    def __init__(self, local_batch, D, H, W, T_max, device, group=None, use_class_ids=True, words_grad=True, w0=1.0, w1=1.0,
                 graph=False, shard=None):
        from . import _lib
        self._lib = _lib
        L = _lib.lib()
        self.group = group
        self.world, self.rank = _world(group)
        self.shard = _shard_mode(shard)
        if self.shard == "images":
            self._init_images(L, int(local_batch), D, H, W, T_max, torch.device(device), use_class_ids, words_grad, w0, w1, graph)
            return
        b, R = int(local_batch), H * W
        Bt = b * self.world
        dev = torch.device(device)
        self.dims = (b, Bt, D, R, T_max)
        f32 = dict(dtype=torch.float32, device=dev)
        self.img = torch.zeros(b, D, H, W, **f32)
        self.words = torch.zeros(b, D, T_max, **f32)
        self.cap_lens32 = torch.full((b,), T_max, dtype=torch.int32, device=dev)
        self.class_ids = torch.arange(self.rank * b, (self.rank + 1) * b, dtype=torch.int64, device=dev) if use_class_ids else None
        self._cls_all = torch.zeros(Bt, dtype=torch.int64, device=dev) if use_class_ids else None
        self._labels = torch.arange(Bt, dtype=torch.int64, device=dev)
        self._img_all = torch.empty(Bt, D, R, **f32)
        self._ws = torch.empty(L.eegan_damsm_pair_workspace_bytes(Bt, b, D, R, T_max), dtype=torch.uint8, device=dev)
        self._m_block = torch.empty(Bt, b, **f32)
        self._m_parts = torch.empty(self.world, Bt, b, **f32)
        self._m_all = torch.empty(Bt, Bt, **f32)
        self._sim = torch.empty(Bt, Bt, **f32)
        self._lse = torch.empty(2, Bt, **f32)
        self._loss01 = torch.zeros(2, **f32)
        self._gvec = torch.tensor([float(w0), float(w1)], **f32)
        self._dsim = torch.empty(Bt, Bt, **f32)
        self._dm_block = torch.empty(Bt, b, **f32)
        self._d_img_all = torch.empty(Bt, D, R, **f32)
        self.att = torch.empty(b, T_max, R, **f32)
        self.d_img = torch.zeros(b, D, H, W, **f32)
        self.d_words = torch.zeros(b, D, T_max, **f32) if words_grad else None
        self._use_graph, self._graph = bool(graph), None
        self._phased = L.eegan_get_contraction_engine() == 3 and D % 128 == 0

    def _init_images(self, L, b, D, H, W, T_max, dev, use_class_ids, words_grad, w0, w1, graph):
        """Static buffers of the image-row partition: own images x all captions."""
        R = H * W
        Bt = b * self.world
        self.dims = (b, Bt, D, R, T_max)
        f32 = dict(dtype=torch.float32, device=dev)
        self.img = torch.zeros(b, D, H, W, **f32)
        self.words = torch.zeros(b, D, T_max, **f32)
        self.cap_lens32 = torch.full((b,), T_max, dtype=torch.int32, device=dev)
        self.class_ids = torch.arange(self.rank * b, (self.rank + 1) * b, dtype=torch.int64, device=dev) if use_class_ids else None
        self._cls_all = torch.zeros(Bt, dtype=torch.int64, device=dev) if use_class_ids else None
        self._labels = torch.arange(Bt, dtype=torch.int64, device=dev)
        self._words_all = torch.empty(Bt, D, T_max, **f32)
        self._lens_all = torch.empty(Bt, dtype=torch.int32, device=dev)
        self._ws = torch.empty(L.eegan_damsm_pair_workspace_bytes(b, Bt, D, R, T_max), dtype=torch.uint8, device=dev)
        self._m_block = torch.empty(b, Bt, **f32)
        self._m_all = torch.empty(Bt, Bt, **f32)
        self._sim = torch.empty(Bt, Bt, **f32)
        self._lse = torch.empty(2, Bt, **f32)
        self._loss01 = torch.zeros(2, **f32)
        self._gvec = torch.tensor([float(w0), float(w1)], **f32)
        self._dsim = torch.empty(Bt, Bt, **f32)
        self._att_all = torch.empty(Bt, T_max, R, **f32)
        self.att = self._att_all[self.rank * b:(self.rank + 1) * b]
        self.d_img = torch.zeros(b, D, H, W, **f32)
        self._d_words_all = torch.empty(Bt, D, T_max, **f32) if words_grad else None
        self.d_words = torch.zeros(b, D, T_max, **f32) if words_grad else None
        self._use_graph, self._graph = bool(graph), None
        self._phased = False

    def _enqueue_images(self):
        _lib = self._lib
        L, p, st = _lib.lib(), _lib.ptr, _lib.stream_ptr()
        b, Bt, D, R, Tm = self.dims
        g1, g2, g3 = gammas()
        rank, grp = self.rank, self.group
        h_cls = None
        if self.world > 1:
            dist.all_gather_into_tensor(self._words_all, self.words, group=grp)
            dist.all_gather_into_tensor(self._lens_all, self.cap_lens32, group=grp)
            if self._cls_all is not None:  # only the cross-entropy needs the class ids: gathered behind the pair grid
                h_cls = dist.all_gather_into_tensor(self._cls_all, self.class_ids, group=grp, async_op=True)
        else:
            self._words_all.copy_(self.words)
            self._lens_all.copy_(self.cap_lens32)
            if self._cls_all is not None:
                self._cls_all.copy_(self.class_ids)
        dev = self.img.device
        with torch.cuda.device(dev):
            _lib.check(L.eegan_damsm_pair_fwd(p(self.img), p(self._words_all), p(self._lens_all), b, Bt, D, R, Tm, g1, g2,
                                              p(self._m_block), p(self._att_all), -rank * b, p(self._ws), self._ws.numel(), st), "damsm_pair_fwd")
        if self.world > 1:  # row blocks: the gathered tensor IS the grid
            dist.all_gather_into_tensor(self._m_all, self._m_block, group=grp)
        else:
            self._m_all.copy_(self._m_block)
        if h_cls is not None:
            h_cls.wait()
        with torch.cuda.device(dev):
            _lib.check(L.eegan_pair_ce_fwd(p(self._m_all), g3, p(self._cls_all), p(self._labels), Bt, p(self._sim), p(self._loss01),
                                           p(self._lse), st), "pair_ce_fwd")
            _lib.check(L.eegan_pair_ce_bwd(p(self._sim), p(self._lse), p(self._labels), p(self._gvec), g3, Bt, p(self._dsim), st),
                       "pair_ce_bwd")
            dm_block = self._dsim[rank * b:(rank + 1) * b]  # the rank's rows: contiguous, no copy
            _lib.check(L.eegan_damsm_pair_bwd(p(self.img), p(self._words_all), p(self._lens_all), b, Bt, D, R, Tm, g1, g2,
                                              p(dm_block), p(self.d_img), p(self._d_words_all), p(self._ws), self._ws.numel(), st),
                       "damsm_pair_bwd")
        if self.d_words is not None:
            if self.world > 1:
                dist.reduce_scatter_tensor(self.d_words, self._d_words_all, op=dist.ReduceOp.SUM, group=grp)
            else:
                self.d_words.copy_(self._d_words_all)

    def _enqueue(self):
        if self.shard == "images":
            return self._enqueue_images()
        _lib = self._lib
        L, p, st = _lib.lib(), _lib.ptr, _lib.stream_ptr()
        b, Bt, D, R, Tm = self.dims
        g1, g2, g3 = gammas()
        rank, grp = self.rank, self.group
        h_cls = None
        if self.world > 1:
            dist.all_gather_into_tensor(self._img_all, self.img.view(b, D, R), group=grp)
            if self._cls_all is not None:  # only the cross-entropy needs the class ids: gathered behind the pair grid
                h_cls = dist.all_gather_into_tensor(self._cls_all, self.class_ids, group=grp, async_op=True)
        else:
            self._img_all.copy_(self.img.view(b, D, R))
            if self._cls_all is not None:
                self._cls_all.copy_(self.class_ids)
        with torch.cuda.device(self.img.device):
            _lib.check(L.eegan_damsm_pair_fwd(p(self._img_all), p(self.words), p(self.cap_lens32), Bt, b, D, R, Tm, g1, g2,
                                              p(self._m_block), p(self.att), rank * b, p(self._ws), self._ws.numel(), st), "damsm_pair_fwd")
        if self.world > 1:
            dist.all_gather_into_tensor(self._m_parts.view(self.world * Bt, b), self._m_block, group=grp)
            self._m_all.view(Bt, self.world, b).copy_(self._m_parts.permute(1, 0, 2))
        else:
            self._m_all.copy_(self._m_block)
        if h_cls is not None:
            h_cls.wait()
        with torch.cuda.device(self.img.device):
            _lib.check(L.eegan_pair_ce_fwd(p(self._m_all), g3, p(self._cls_all), p(self._labels), Bt, p(self._sim), p(self._loss01),
                                           p(self._lse), st), "pair_ce_fwd")
            _lib.check(L.eegan_pair_ce_bwd(p(self._sim), p(self._lse), p(self._labels), p(self._gvec), g3, Bt, p(self._dsim), st),
                       "pair_ce_bwd")
        self._dm_block.copy_(self._dsim[:, rank * b:(rank + 1) * b])
        split = self.world > 1 and self.d_words is not None and self._phased
        with torch.cuda.device(self.img.device):
            if split:  # image part first: the reduce-scatter of d_img then runs behind the words part (GEMM5 + unpack)
                _lib.check(L.eegan_damsm_pair_bwd_phased(p(self._img_all), p(self.words), p(self.cap_lens32), Bt, b, D, R, Tm, g1, g2,
                                                         p(self._dm_block), p(self._d_img_all), None, 1, p(self._ws), self._ws.numel(), st),
                           "damsm_pair_bwd(image part)")
                h_rs = dist.reduce_scatter_tensor(self.d_img.view(b, D, R), self._d_img_all, op=dist.ReduceOp.SUM, group=grp, async_op=True)
                _lib.check(L.eegan_damsm_pair_bwd_phased(p(self._img_all), p(self.words), p(self.cap_lens32), Bt, b, D, R, Tm, g1, g2,
                                                         p(self._dm_block), None, p(self.d_words), 2, p(self._ws), self._ws.numel(), st),
                           "damsm_pair_bwd(words part)")
                h_rs.wait()
                return
            _lib.check(L.eegan_damsm_pair_bwd(p(self._img_all), p(self.words), p(self.cap_lens32), Bt, b, D, R, Tm, g1, g2,
                                              p(self._dm_block), p(self._d_img_all), p(self.d_words), p(self._ws), self._ws.numel(), st),
                       "damsm_pair_bwd")
        if self.world > 1:
            dist.reduce_scatter_tensor(self.d_img.view(b, D, R), self._d_img_all, op=dist.ReduceOp.SUM, group=grp)
        else:
            self.d_img.view(b, D, R).copy_(self._d_img_all)

    def load(self, img, words, cap_lens, class_ids=None):
        self.img.copy_(img.reshape(self.img.shape), non_blocking=True)
        self.words.copy_(words, non_blocking=True)
        self.cap_lens32.copy_(torch.as_tensor(cap_lens).reshape(-1), non_blocking=True)
        if self.class_ids is not None and class_ids is not None:
            self.class_ids.copy_(torch.as_tensor(class_ids).reshape(-1), non_blocking=True)

    def run(self):
        """Enqueue (or replay) one step on the data already in the static input buffers."""
        if not self._use_graph:
            self._enqueue()
            return
        if self._graph is None:
            side = torch.cuda.Stream(device=self.img.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._enqueue()
            torch.cuda.current_stream().wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._enqueue()
        self._graph.replay()

    def __call__(self, img_features, words_emb, cap_lens, class_ids=None):
        self.load(img_features, words_emb, cap_lens, class_ids)
        self.run()
        return self._loss01[0], self._loss01[1], self.d_img, self.d_words

    def release_graph(self):
        """Drop the captured graph (it holds NCCL work on the communicator): call before
        ``torch.distributed.destroy_process_group()`` — tearing the communicator down under a live graph that
        captured its collectives hangs on torch 2.11 / NCCL 2.28."""
        if self._graph is not None:
            torch.cuda.synchronize(self.img.device)
            self._graph = None
            torch.cuda.synchronize(self.img.device)


class _CudaStepKernels:
    """The library calls of the sharded step (C ABI on the current stream)."""

    def __init__(self):
        from . import _lib
        self._lib = _lib

    def workspace(self, Bi, Bc, D, R, Tm, device):
        n = self._lib.lib().eegan_damsm_pair_workspace_bytes(Bi, Bc, D, R, Tm)
        return torch.empty(n, dtype=torch.uint8, device=device)

    def pair_fwd(self, img, words, lens32, Bi, Bc, D, R, Tm, m, att, ws):
        _lib = self._lib
        g1, g2, _ = gammas()
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().eegan_damsm_pair_fwd(_lib.ptr(img), _lib.ptr(words), _lib.ptr(lens32), Bi, Bc, D, R, Tm, g1, g2,
                                                       _lib.ptr(m), _lib.ptr(att), 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       "damsm_pair_fwd")

    def pair_bwd(self, img, words, lens32, Bi, Bc, D, R, Tm, dm, d_img, d_words, ws):
        _lib = self._lib
        g1, g2, _ = gammas()
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().eegan_damsm_pair_bwd(_lib.ptr(img), _lib.ptr(words), _lib.ptr(lens32), Bi, Bc, D, R, Tm, g1, g2,
                                                       _lib.ptr(dm), _lib.ptr(d_img), _lib.ptr(d_words), _lib.ptr(ws), ws.numel(),
                                                       _lib.stream_ptr()), "damsm_pair_bwd")

    def ce(self, m_all, cls_all, labels, gvec, Bt, sim, lse, loss01, dsim):
        _lib = self._lib
        L, p, st = _lib.lib(), _lib.ptr, _lib.stream_ptr()
        _, _, g3 = gammas()
        with torch.cuda.device(m_all.device):
            _lib.check(L.eegan_pair_ce_fwd(p(m_all), g3, p(cls_all), p(labels), Bt, p(sim), p(loss01), p(lse), st), "pair_ce_fwd")
            _lib.check(L.eegan_pair_ce_bwd(p(sim), p(lse), p(labels), p(gvec), g3, Bt, p(dsim), st), "pair_ce_bwd")


class OverlappedShardedWordsLossStep:
    """ShardedWordsLossStep with the two large collectives hidden behind the rank's LOCAL image block (opt-in; same contract:
    ``loss0, loss1, d_img, d_words = step(img, words, cap_lens, class_ids)`` on static tensors, no autograd).

    The rank's own images need no communication, so the grid is evaluated in two calls per direction:

      fwd   all-gather(img) starts (async)   ||  pair grid of the LOCAL images x own captions (b x b pairs, attention maps)
            wait                              ->  pair grid of the REMOTE images x own captions ((N-1) b x b pairs)
            all-gather of the m blocks, two-way CE forward / backward on every rank (O(B^2), redundant)
      bwd   backward of the REMOTE block: partial d_img for the other ranks' images, d_words part
            reduce-scatter(d_img; own slot = zeros) starts (async)   ||  backward of the LOCAL block
            wait -> d_img = scattered sum + local part;  d_words = remote part + local part

    The gathered features live in a buffer ROTATED by the rank (slot s holds the images of rank (rank + s) % N), so that
    the remote images are one contiguous array for the second call; rows of m / dm are permuted between rank order and
    slot order by two small index_selects.  The collectives are the list forms of torch.distributed (NCCL on the GPU box;
    gloo in tests/test_sharded_gloo.py, where ``kernels`` is an oracle-based stand-in for the C ABI calls).
    The result equals the single-device full-batch result (summation order of d_words differs: two partial sums).
    """

    def __init__(self, local_batch, D, H, W, T_max, device, group=None, use_class_ids=True, words_grad=True, w0=1.0, w1=1.0,
                 kernels=None, dtype=torch.float32):
        self.group = group
        self.world, self.rank = _world(group)
        self.K = kernels if kernels is not None else _CudaStepKernels()
        b, R, N = int(local_batch), H * W, self.world
        Bt = b * N
        dev = torch.device(device)
        self.dims = (b, Bt, D, R, T_max)
        f = dict(dtype=dtype, device=dev)
        self.img = torch.zeros(b, D, H, W, **f)
        self.words = torch.zeros(b, D, T_max, **f)
        self.cap_lens32 = torch.full((b,), T_max, dtype=torch.int32, device=dev)
        self.class_ids = torch.arange(self.rank * b, (self.rank + 1) * b, dtype=torch.int64, device=dev) if use_class_ids else None
        self._cls_all = torch.zeros(Bt, dtype=torch.int64, device=dev) if use_class_ids else None
        self._labels = torch.arange(Bt, dtype=torch.int64, device=dev)
        self._img_rot = torch.zeros(Bt, D, R, **f)            # slot s = images of rank (rank + s) % N; slot 0 is not read
        self._ws_loc = self.K.workspace(b, b, D, R, T_max, dev)
        self._ws_rem = self.K.workspace(Bt - b, b, D, R, T_max, dev) if N > 1 else None
        self._m_rot = torch.empty(Bt, b, **f)                 # rows in slot order
        self._m_block = torch.empty(Bt, b, **f)               # rows in rank order
        self._m_parts = torch.empty(N, Bt, b, **f)
        self._m_all = torch.empty(Bt, Bt, **f)
        self._sim = torch.empty(Bt, Bt, **f)
        self._lse = torch.empty(2, Bt, **f)
        self._loss01 = torch.zeros(2, **f)
        self._gvec = torch.tensor([float(w0), float(w1)], **f)
        self._dsim = torch.empty(Bt, Bt, **f)
        self._dm_block = torch.empty(Bt, b, **f)
        self._dm_rot = torch.empty(Bt, b, **f)
        self._d_img_rot = torch.zeros(Bt, D, R, **f)          # slots 1.. = partial gradients for the other ranks' images
        self._zero = torch.zeros(b, D, R, **f)                # the rank's own slot of the reduce-scatter
        self._rs_out = torch.empty(b, D, R, **f)
        self._d_img_loc = torch.empty(b, D, R, **f)
        self._dw_loc = torch.zeros(b, D, T_max, **f) if words_grad else None
        self._dw_rem = torch.zeros(b, D, T_max, **f) if words_grad else None
        self.att = torch.empty(b, T_max, R, **f)
        self.d_img = torch.zeros(b, D, H, W, **f)
        self.d_words = torch.zeros(b, D, T_max, **f) if words_grad else None
        # rank order <-> slot order of the Bt image rows
        slot_of_rank = [(r - self.rank) % N for r in range(N)]
        rows = torch.arange(b)
        self._to_rank_order = torch.cat([slot_of_rank[r] * b + rows for r in range(N)]).to(dev)      # m_block = m_rot[idx]
        self._to_slot_order = torch.cat([((self.rank + s) % N) * b + rows for s in range(N)]).to(dev)  # dm_rot = dm_block[idx]
        self._slot_of_rank = slot_of_rank

    def _slot(self, buf, s):
        b = self.dims[0]
        return buf[s * b:(s + 1) * b]

    def run(self):
        b, Bt, D, R, Tm = self.dims
        N, rank, grp, K = self.world, self.rank, self.group, self.K
        img_loc = self.img.view(b, D, R)
        h = None
        if N > 1:
            outs = [self._slot(self._img_rot, self._slot_of_rank[r]) for r in range(N)]
            h = dist.all_gather(outs, img_loc, group=grp, async_op=True)
            if self._cls_all is not None:
                dist.all_gather_into_tensor(self._cls_all, self.class_ids, group=grp)
        elif self._cls_all is not None:
            self._cls_all.copy_(self.class_ids)
        K.pair_fwd(img_loc, self.words, self.cap_lens32, b, b, D, R, Tm, self._m_rot[:b], self.att, self._ws_loc)
        if N > 1:
            h.wait()
            K.pair_fwd(self._img_rot[b:], self.words, self.cap_lens32, Bt - b, b, D, R, Tm, self._m_rot[b:], None, self._ws_rem)
            torch.index_select(self._m_rot, 0, self._to_rank_order, out=self._m_block)
            dist.all_gather_into_tensor(self._m_parts.view(N * Bt, b), self._m_block, group=grp)
            self._m_all.view(Bt, N, b).copy_(self._m_parts.permute(1, 0, 2))
        else:
            self._m_all.copy_(self._m_rot)
        K.ce(self._m_all, self._cls_all, self._labels, self._gvec, Bt, self._sim, self._lse, self._loss01, self._dsim)
        self._dm_block.copy_(self._dsim[:, rank * b:(rank + 1) * b])
        torch.index_select(self._dm_block, 0, self._to_slot_order, out=self._dm_rot)
        h2 = None
        if N > 1:
            K.pair_bwd(self._img_rot[b:], self.words, self.cap_lens32, Bt - b, b, D, R, Tm, self._dm_rot[b:], self._d_img_rot[b:],
                       self._dw_rem, self._ws_rem)
            ins = [self._zero if r == rank else self._slot(self._d_img_rot, self._slot_of_rank[r]) for r in range(N)]
            h2 = dist.reduce_scatter(self._rs_out, ins, op=dist.ReduceOp.SUM, group=grp, async_op=True)
        K.pair_bwd(img_loc, self.words, self.cap_lens32, b, b, D, R, Tm, self._dm_rot[:b], self._d_img_loc, self._dw_loc, self._ws_loc)
        if N > 1:
            h2.wait()
            torch.add(self._rs_out, self._d_img_loc, out=self.d_img.view(b, D, R))
            if self.d_words is not None:
                torch.add(self._dw_loc, self._dw_rem, out=self.d_words)
        else:
            self.d_img.view(b, D, R).copy_(self._d_img_loc)
            if self.d_words is not None:
                self.d_words.copy_(self._dw_loc)

    def load(self, img, words, cap_lens, class_ids=None):
        self.img.copy_(img.reshape(self.img.shape), non_blocking=True)
        self.words.copy_(words, non_blocking=True)
        self.cap_lens32.copy_(torch.as_tensor(cap_lens).reshape(-1), non_blocking=True)
        if self.class_ids is not None and class_ids is not None:
            self.class_ids.copy_(torch.as_tensor(class_ids).reshape(-1), non_blocking=True)

    def __call__(self, img_features, words_emb, cap_lens, class_ids=None):
        self.load(img_features, words_emb, cap_lens, class_ids)
        self.run()
        return self._loss01[0], self._loss01[1], self.d_img, self.d_words

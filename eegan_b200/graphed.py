"""CUDA-graph capture of the DAMSM words_loss step.

At CUB sizes the eager path is host-bound: ~20 kernel launches, five TMA descriptor encodes
and the autograd bookkeeping cost more CPU time per step than the GPU needs to run the kernels.
`GraphedWordsLoss` captures forward + backward of `words_loss` (miscc/DAMSM_losses.py:272-342)
once for a fixed shape and replays it; inputs are copied into static buffers, outputs (the two
losses and both input gradients) are static tensors.  The single-device graph holds only the
library's own kernels — pair grid forward, two-way CE forward / backward, pair grid backward,
called through the C ABI in the order autograd would run them — chained by programmatic dependent
launch; the autograd glue of the eager call (fills, stack, add) has no counterpart in it.  The
sharded (N > 1) step still goes through autograd because of its collectives.
"""
from __future__ import annotations

import torch

from . import _lib
from . import damsm_losses as dl
from .config import gammas


class GraphedWordsLoss:
    """words_loss forward+backward as one CUDA graph.

    loss = w0 * loss0 + w1 * loss1 is differentiated w.r.t. img (and words when
    ``words_grad=True``; train.py:172 detaches the words).  ``class_ids`` must be given as
    int64 values per call (or None at construction for "no class mask")."""

    def __init__(self, batch_size, D, H, W, T_max, device, use_class_ids=True, words_grad=True, w0=1.0, w1=1.0,
                 warmup=3, sharded=False, group=None):
        self.B = batch_size
        dev = torch.device(device)
        self.img = torch.zeros(batch_size, D, H, W, device=dev).requires_grad_()
        self.words = torch.zeros(batch_size, D, T_max, device=dev).requires_grad_(words_grad)
        self.cap_lens = torch.full((batch_size,), T_max, dtype=torch.int64, device=dev)
        self.class_ids = torch.arange(batch_size, dtype=torch.int64, device=dev) if use_class_ids else None
        self.labels = torch.arange(batch_size, dtype=torch.int64, device=dev)
        self.w0, self.w1 = float(w0), float(w1)
        self.graph = None
        self._warmup = warmup
        # sharded=True: batch_size is the rank's local batch; the loss is the global-batch loss of
        # eegan_b200.sharded.sharded_words_loss (NCCL collectives are captured into the graph)
        self.sharded, self.group = sharded, group
        self.loss0 = self.loss1 = None
        self.words_grad = bool(words_grad)
        if not sharded:
            L = _lib.lib()
            R = H * W
            self._dims = (batch_size, D, R, T_max)
            self.cap_lens32 = torch.full((batch_size,), T_max, dtype=torch.int32, device=dev)
            need = L.eegan_damsm_pair_workspace_bytes(batch_size, batch_size, D, R, T_max)
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._m = torch.empty(batch_size, batch_size, device=dev)
            self._sim = torch.empty(batch_size, batch_size, device=dev)
            self._lse = torch.empty(2, batch_size, device=dev)
            self._loss01 = torch.zeros(2, device=dev)
            self._gvec = torch.tensor([self.w0, self.w1], dtype=torch.float32, device=dev)
            self._dm = torch.empty(batch_size, batch_size, device=dev)
            self.att = torch.empty(batch_size, T_max, R, device=dev)
            self._d_img = torch.zeros(batch_size, D, H, W, device=dev)
            self._d_words = torch.zeros(batch_size, D, T_max, device=dev) if words_grad else None

    def _direct_step(self):
        """Forward + backward of words_loss through the C ABI, in autograd's order, on static buffers."""
        L = _lib.lib()
        B, D, R, Tm = self._dims
        g1, g2, g3 = gammas()
        p, st = _lib.ptr, _lib.stream_ptr()
        with torch.cuda.device(self.img.device):
            _lib.check(L.eegan_damsm_pair_fwd(p(self.img), p(self.words), p(self.cap_lens32), B, B, D, R, Tm, g1, g2, p(self._m),
                                              p(self.att), 0, p(self._ws), self._ws.numel(), st), "damsm_pair_fwd")
            _lib.check(L.eegan_pair_ce_fwd(p(self._m), g3, p(self.class_ids), p(self.labels), B, p(self._sim), p(self._loss01),
                                           p(self._lse), st), "pair_ce_fwd")
            _lib.check(L.eegan_pair_ce_bwd(p(self._sim), p(self._lse), p(self.labels), p(self._gvec), g3, B, p(self._dm), st),
                       "pair_ce_bwd")
            _lib.check(L.eegan_damsm_pair_bwd(p(self.img), p(self.words), p(self.cap_lens32), B, B, D, R, Tm, g1, g2, p(self._dm),
                                              p(self._d_img), p(self._d_words), p(self._ws), self._ws.numel(), st), "damsm_pair_bwd")
        return self._loss01[0], self._loss01[1]

    def _step(self):
        if not self.sharded:
            return self._direct_step()
        self.img.grad = None
        self.words.grad = None
        if self.sharded:
            from .sharded import sharded_words_loss
            l0, l1, _ = sharded_words_loss(self.img, self.words, self.labels, self.cap_lens, self.class_ids, self.B,
                                           group=self.group)
        else:
            l0, l1, _ = dl.words_loss(self.img, self.words, self.labels, self.cap_lens, self.class_ids, self.B)
        (self.w0 * l0 + self.w1 * l1).backward()
        return l0.detach(), l1.detach()

    def _load(self, img, words, cap_lens, class_ids):
        with torch.no_grad():
            self.img.copy_(img.reshape(self.img.shape), non_blocking=True)
            self.words.copy_(words, non_blocking=True)
            self.cap_lens.copy_(torch.as_tensor(cap_lens).reshape(-1), non_blocking=True)
            if not self.sharded:
                self.cap_lens32.copy_(self.cap_lens, non_blocking=True)
            if self.class_ids is not None and class_ids is not None:
                self.class_ids.copy_(torch.as_tensor(class_ids).reshape(-1), non_blocking=True)

    def capture(self):
        """Warm up on a side stream, then capture.  Called lazily by the first __call__."""
        s = torch.cuda.Stream(device=self.img.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(self._warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        self.img.grad = None
        self.words.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss0, self.loss1 = self._step()
        return self

    def __call__(self, img_features, words_emb, cap_lens, class_ids=None):
        """Returns (loss0, loss1, d_img, d_words): static tensors, valid until the next call."""
        if self.graph is None:
            self._load(img_features, words_emb, cap_lens, class_ids)
            self.capture()
        self._load(img_features, words_emb, cap_lens, class_ids)
        self.graph.replay()
        if not self.sharded:
            return self.loss0, self.loss1, self._d_img, self._d_words
        return self.loss0, self.loss1, self.img.grad, self.words.grad

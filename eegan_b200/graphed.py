"""CUDA-graph capture of the DAMSM words_loss step.

At CUB sizes the eager path is host-bound: ~20 kernel launches, five TMA descriptor encodes
and the autograd bookkeeping cost more CPU time per step than the GPU needs to run the kernels.
`GraphedWordsLoss` captures forward + backward of `words_loss` (miscc/DAMSM_losses.py:272-342)
once for a fixed shape and replays it; inputs are copied into static buffers, outputs (the two
losses and both input gradients) are static tensors.  Everything the graph launches is the
same set of kernels the eager call launches.
"""
from __future__ import annotations

import torch

from . import damsm_losses as dl


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

    def _step(self):
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
        return self.loss0, self.loss1, self.img.grad, self.words.grad

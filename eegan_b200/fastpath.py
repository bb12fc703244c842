"""The reference's own call — ``words_loss(...)`` followed by ``.backward()`` (train.py:428, 500) — at the speed
of the captured step.

At CUB sizes one step is ~0.25 ms of device work spread over ~15 kernels; the plain autograd route spends 4-6x
that on the host (two autograd Functions, a 300 MB workspace allocation, five TMA-descriptor encodes, ~20 launches
through ctypes).  This module keeps, per (device, shape), a *plan*: the caller-owned workspace of the C ABI,
static input / output buffers, and two CUDA graphs — forward (pair grid + two-way CE) and backward (CE backward +
pair-grid backward) — captured on the second call with that shape and replayed from then on.  One
``torch.autograd.Function`` spans pair grid and cross-entropy, so ``loss0`` / ``loss1`` come out of a single node
and the backward is: copy the two upstream scalars into a device vector, replay, clone the two gradients.

What the caller may rely on (same contract as the un-planned route):
  * inputs are not modified, outputs are fresh tensors (gradients and attention maps are cloned out of the plan);
  * several forwards may be in flight before their backwards: the stash belongs to the plan, so a backward whose
    forward is no longer the plan's latest re-runs that forward first (generation counter) — correct, just slower;
  * ``backward`` twice on one forward (``retain_graph=True``) is valid;
  * inside someone else's stream capture the plan launches the same kernels directly instead of nesting a capture.
Plans are evicted least-recently-used beyond ``_MAX_PLAN_BYTES``.  ``EEGAN_WORDS_LOSS_PLAN=0`` (read once) selects
the un-planned route.
"""
from __future__ import annotations

import os
import threading
from collections import OrderedDict

import torch

from . import _lib

_MAX_PLAN_BYTES = 64 << 30
_ENABLED = os.environ.get("EEGAN_WORDS_LOSS_PLAN", "1") != "0"
_plans: "OrderedDict[tuple, _Plan]" = OrderedDict()
_lock = threading.Lock()


def enabled() -> bool:
    return _ENABLED


def clear_plans() -> None:
    with _lock:
        _plans.clear()


class _Plan:
    def __init__(self, dev, Bi, Bc, D, R, Tm, has_cls, gam, diag_offset):
        L = _lib.lib()
        self.dev, self.dims, self.gam, self.diag = dev, (Bi, Bc, D, R, Tm), gam, int(diag_offset)
        f32 = dict(dtype=torch.float32, device=dev)
        self.img = torch.empty(Bi, D, R, **f32)
        self.words = torch.empty(Bc, D, Tm, **f32)
        self.lens32 = torch.empty(Bc, dtype=torch.int32, device=dev)
        self.cls = torch.empty(Bi, dtype=torch.int64, device=dev) if has_cls else None
        self.labels = torch.empty(Bi, dtype=torch.int64, device=dev)
        need = L.eegan_damsm_pair_workspace_bytes(Bi, Bc, D, R, Tm)
        self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self.m = torch.empty(Bi, Bc, **f32)
        self.sim = torch.empty(Bi, Bc, **f32)
        self.lse = torch.empty(2, Bi, **f32)
        self.loss01 = torch.zeros(2, **f32)
        self.gvec = torch.zeros(2, **f32)
        self.dm = torch.empty(Bi, Bc, **f32)
        self.att = torch.empty(Bc, Tm, R, **f32)
        self.d_img = torch.empty(Bi, D, R, **f32)
        self.d_words = torch.empty(Bc, D, Tm, **f32)
        self.bytes = need + 4 * (2 * self.img.numel() + 2 * self.words.numel() + self.att.numel() + 4 * Bi * Bc)
        self.gen = 0          # bumped by every forward: tells a backward whether the stash is still its own
        self.calls = 0
        self.fwd_graph = None
        self.bwd_graphs = {}  # (need_img, need_words) -> CUDAGraph
        self.engine = L.eegan_get_contraction_engine()
        self._labels_tag = self._labels_ref = None

    # ---- the library calls, on whatever stream is current -------------------------------------------------
    def enqueue_fwd(self):
        L, p, st = _lib.lib(), _lib.ptr, _lib.stream_ptr()
        Bi, Bc, D, R, Tm = self.dims
        g1, g2, g3 = self.gam
        with torch.cuda.device(self.dev):
            _lib.check(L.eegan_damsm_pair_fwd(p(self.img), p(self.words), p(self.lens32), Bi, Bc, D, R, Tm, g1, g2, p(self.m),
                                              p(self.att), self.diag, p(self.ws), self.ws.numel(), st), "damsm_pair_fwd")
            _lib.check(L.eegan_pair_ce_fwd(p(self.m), g3, p(self.cls), p(self.labels), Bi, p(self.sim), p(self.loss01),
                                           p(self.lse), st), "pair_ce_fwd")

    def enqueue_bwd(self, need_img, need_words):
        L, p, st = _lib.lib(), _lib.ptr, _lib.stream_ptr()
        Bi, Bc, D, R, Tm = self.dims
        g1, g2, g3 = self.gam
        with torch.cuda.device(self.dev):
            _lib.check(L.eegan_pair_ce_bwd(p(self.sim), p(self.lse), p(self.labels), p(self.gvec), g3, Bi, p(self.dm), st),
                       "pair_ce_bwd")
            _lib.check(L.eegan_damsm_pair_bwd(p(self.img), p(self.words), p(self.lens32), Bi, Bc, D, R, Tm, g1, g2, p(self.dm),
                                              p(self.d_img) if need_img else None, p(self.d_words) if need_words else None,
                                              p(self.ws), self.ws.numel(), st), "damsm_pair_bwd")

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            fn()
        return g

    @staticmethod
    def _flat(t, n):
        return t if (t.dim() == 1 and t.shape[0] == n) else t.reshape(-1)[:n]

    def load(self, img3, words, cap_lens, cls, labels):
        """Inputs -> the plan's static buffers (every host-side op here is on the critical path of a host-bound step)."""
        self.img.copy_(img3, non_blocking=True)
        self.words.copy_(words, non_blocking=True)
        self.lens32.copy_(self._flat(cap_lens, self.lens32.shape[0]), non_blocking=True)
        if self.cls is not None:
            self.cls.copy_(self._flat(cls, self.cls.shape[0]), non_blocking=True)
        # labels are the same device tensor step after step (train.py:93 builds match_labels once): skip the copy when the
        # very same, unmodified tensor comes back (identity + autograd version counter)
        tag = (id(labels), labels._version, labels.data_ptr())
        if tag != self._labels_tag or not labels.is_cuda:
            self.labels.copy_(self._flat(labels, self.labels.shape[0]), non_blocking=True)
            self._labels_tag = tag if labels.is_cuda else None
            self._labels_ref = labels if labels.is_cuda else None  # keeps id() from being recycled

    def run_fwd(self, flags):
        """flags = (need_img, need_words) of the backward that may follow (captured together with the forward)."""
        self.gen += 1
        self.calls += 1
        if torch.cuda.is_current_stream_capturing():
            self.enqueue_fwd()
            return
        if self.fwd_graph is None:
            if self.calls == 1:  # first call with this shape: plain launches (also warms every lazy one-off of the library)
                self.enqueue_fwd()
                return
            self.fwd_graph = self._capture(self.enqueue_fwd)
        if (flags[0] or flags[1]) and flags not in self.bwd_graphs and self.calls > 1:
            self.bwd_graphs[flags] = self._capture(lambda: self.enqueue_bwd(*flags))
        self.fwd_graph.replay()

    def run_bwd(self, flags):
        g = None if torch.cuda.is_current_stream_capturing() else self.bwd_graphs.get(flags)
        if g is not None:
            g.replay()
        else:
            self.enqueue_bwd(*flags)


def _get_plan(key, make):
    with _lock:
        plan = _plans.get(key)
        if plan is not None:
            _plans.move_to_end(key)
            return plan
    plan = make()
    with _lock:
        _plans[key] = plan
        total = sum(p.bytes for p in _plans.values())
        while total > _MAX_PLAN_BYTES and len(_plans) > 1:
            _, old = _plans.popitem(last=False)
            total -= old.bytes
    return plan


class _WordsLossFn(torch.autograd.Function):
    """(loss0, loss1, att [B_cap, T_max, R]) of words_loss, DAMSM_losses.py:272-342, as one autograd node."""

    @staticmethod
    def forward(ctx, img, words, cap_lens, cls, labels, gam, diag_offset):
        Bi, D, R = img.shape
        Bc, _, Tm = words.shape
        dev = img.device
        L = _lib.lib()
        key = (dev.index, Bi, Bc, D, R, Tm, cls is not None, gam, int(diag_offset), L.eegan_get_contraction_engine())
        plan = _get_plan(key, lambda: _Plan(dev, Bi, Bc, D, R, Tm, cls is not None, gam, diag_offset))
        flags = (bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]))
        plan.load(img, words, cap_lens, cls, labels)
        plan.run_fwd(flags)
        ctx.plan, ctx.gen, ctx.flags = plan, plan.gen, flags
        ctx.save_for_backward(img, words)
        ctx.aux = (cap_lens, cls, labels)
        loss0, loss1 = plan.loss01.clone().unbind(0)
        att = plan.att.clone()
        ctx.mark_non_differentiable(att)
        return loss0, loss1, att

    @staticmethod
    def backward(ctx, g0, g1, _gatt):
        plan, flags = ctx.plan, ctx.flags
        img, words = ctx.saved_tensors
        if plan.gen != ctx.gen:  # another forward used the plan since: its stash is not ours any more
            plan.load(img, words, *ctx.aux)
            plan.run_fwd(flags)
            ctx.gen = plan.gen
        if g0 is not None and g1 is not None:
            torch.stack((g0, g1) if g0.dim() == 0 and g1.dim() == 0 else (g0.reshape(()), g1.reshape(())), out=plan.gvec)
        else:
            for k, g in enumerate((g0, g1)):
                if g is None:
                    plan.gvec[k].zero_()
                else:
                    plan.gvec[k].copy_(g.reshape(()), non_blocking=True)
        plan.run_bwd(flags)
        d_img = plan.d_img.clone() if flags[0] else None  # img is the [B_img, D, R] view the wrapper passed: same shape
        d_words = plan.d_words.clone() if flags[1] else None
        return d_img, d_words, None, None, None, None, None


def supported(img, words) -> bool:
    """The planned route takes what the default contraction engine takes."""
    if not (_ENABLED and img.is_cuda):
        return False
    D = img.shape[1]
    return _lib.lib().eegan_get_contraction_engine() >= 2 and D % 128 == 0 and words.shape[2] <= 32 and words.shape[0] <= 4096


def words_loss_planned(img3, words, cap_lens, cls, labels, gam, diag_offset=0):
    """img3 [B_img, D, R] fp32 contiguous CUDA; cap_lens / labels / cls: tensors (any device, integer)."""
    return _WordsLossFn.apply(img3, words, cap_lens, cls, labels, tuple(float(g) for g in gam), int(diag_offset))

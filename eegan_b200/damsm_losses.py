"""Drop-in for the reference's ``miscc/DAMSM_losses.py`` — same names, same positional
signatures, same return structures — backed by the sm_100a kernels of libeegan_b200.so.

    cosine_similarity        miscc/DAMSM_losses.py:17-23
    func_attention           miscc/DAMSM_losses.py:25-63
    GlobalAttentionGeneral   miscc/DAMSM_losses.py:65-132
    sent_similarity          miscc/DAMSM_losses.py:134-166
    words_similarity         miscc/DAMSM_losses.py:168-231
    sent_loss                miscc/DAMSM_losses.py:233-270
    words_loss               miscc/DAMSM_losses.py:272-342

``import eegan_b200; eegan_b200.install()`` makes ``from miscc.DAMSM_losses import
words_loss, sent_loss`` (train.py:24) resolve here, so train.py / models.py / DAMSM.py run
unmodified.  PyTorch is used for device memory, streams and autograd plumbing only; every
arithmetic step of the path runs in the CUDA library.  No CPU path exists.
"""
from __future__ import annotations

import math
import threading
from collections.abc import Sequence

import torch
import torch.nn as nn

from . import _lib
from .config import gammas, get_cfg

__all__ = [
    "cosine_similarity", "func_attention", "GlobalAttentionGeneral", "sent_similarity",
    "words_similarity", "sent_loss", "words_loss",
]


_tls = threading.local()
_GAG_BWD_MAX_IDF = 256  # eegan_gag_bwd's shared-memory forms (gag.cu / gag_bwd2.cu); the forward alone takes idf <= 512


# ---------------------------------------------------------------------------------------
# autograd plumbing
# ---------------------------------------------------------------------------------------
class _PairGridFn(torch.autograd.Function):
    """m[j,i] = log sum_t exp(g2 cos_t) for every (image j, caption i), plus the diagonal
    attention maps.  Kernels: eegan_damsm_pair_fwd / _bwd."""

    @staticmethod
    def forward(ctx, img, words, cap_lens32, g1, g2, diag_offset, want_att):
        L = _lib.lib()
        Bi, D, R = img.shape
        Bc, _, Tm = words.shape
        need = L.eegan_damsm_pair_workspace_bytes(Bi, Bc, D, R, Tm)
        ws = torch.empty(need, dtype=torch.uint8, device=img.device)
        m = torch.empty(Bi, Bc, dtype=torch.float32, device=img.device)
        att = torch.empty(Bc, Tm, R, dtype=torch.float32, device=img.device) if want_att else None
        with torch.cuda.device(img.device):
            _lib.check(L.eegan_damsm_pair_fwd(_lib.ptr(img), _lib.ptr(words), _lib.ptr(cap_lens32), Bi, Bc, D, R, Tm,
                                              g1, g2, _lib.ptr(m), _lib.ptr(att), diag_offset, _lib.ptr(ws), need,
                                              _lib.stream_ptr()), "damsm_pair_fwd")
        ctx.save_for_backward(img, words, cap_lens32)
        ctx.ws, ctx.g = ws, (g1, g2)
        ctx.engine = L.eegan_get_contraction_engine()  # the stash layout belongs to the engine that wrote it
        if att is None:
            att = torch.empty(0, device=img.device)
        ctx.mark_non_differentiable(att)
        return m, att

    @staticmethod
    def backward(ctx, dm, _datt):
        img, words, cap_lens32 = ctx.saved_tensors
        L = _lib.lib()
        Bi, D, R = img.shape
        Bc, _, Tm = words.shape
        need_img, need_words = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_img = torch.empty_like(img) if need_img else None
        d_words = torch.empty_like(words) if need_words else None
        dm = _lib.f32c(dm)
        if ctx.ws is None:
            raise RuntimeError("words_loss backward: the forward stash was consumed by an earlier backward "
                               "(only the fused contraction engines, 2 and 3, keep it for retain_graph)")
        if L.eegan_get_contraction_engine() != ctx.engine:
            raise RuntimeError("words_loss backward: the contraction engine changed since the forward")
        with torch.cuda.device(img.device):
            _lib.check(L.eegan_damsm_pair_bwd(_lib.ptr(img), _lib.ptr(words), _lib.ptr(cap_lens32), Bi, Bc, D, R, Tm,
                                              ctx.g[0], ctx.g[1], _lib.ptr(dm), _lib.ptr(d_img), _lib.ptr(d_words),
                                              _lib.ptr(ctx.ws), ctx.ws.numel(), _lib.stream_ptr()), "damsm_pair_bwd")
        if ctx.engine < 2 or D % 128:
            ctx.ws = None  # engines 0/1 consume the stash (U and dA are overwritten in place); the fused engines keep it
        return d_img, d_words, None, None, None, None, None


class _PairCEFn(torch.autograd.Function):
    """Scale + class mask + CE over rows and columns.  Kernels: eegan_pair_ce_fwd / _bwd."""

    @staticmethod
    def forward(ctx, scores, scale, class_ids, labels):
        L = _lib.lib()
        B = scores.shape[0]
        out = torch.empty_like(scores)
        loss01 = torch.empty(2, dtype=torch.float32, device=scores.device)
        lse = torch.empty(2, B, dtype=torch.float32, device=scores.device)
        with torch.cuda.device(scores.device):
            _lib.check(L.eegan_pair_ce_fwd(_lib.ptr(scores), scale, _lib.ptr(class_ids), _lib.ptr(labels), B,
                                           _lib.ptr(out), _lib.ptr(loss01), _lib.ptr(lse), _lib.stream_ptr()),
                       "pair_ce_fwd")
        ctx.save_for_backward(out, lse, labels)
        ctx.scale = scale
        ctx.mark_non_differentiable(out)
        return loss01[0], loss01[1], out

    @staticmethod
    def backward(ctx, g0, g1, _gout):
        out, lse, labels = ctx.saved_tensors
        L = _lib.lib()
        B = out.shape[0]
        z = torch.zeros((), dtype=torch.float32, device=out.device)
        g = torch.stack([g0.float() if g0 is not None else z, g1.float() if g1 is not None else z]).contiguous()
        ds = torch.empty_like(out)
        with torch.cuda.device(out.device):
            _lib.check(L.eegan_pair_ce_bwd(_lib.ptr(out), _lib.ptr(lse), _lib.ptr(labels), _lib.ptr(g), ctx.scale, B,
                                           _lib.ptr(ds), _lib.stream_ptr()), "pair_ce_bwd")
        return ds, None, None, None


class _ScaleMaskFn(torch.autograd.Function):
    """scores * scale with the class mask set to -inf (the *_similarity API): the forward
    half of eegan_pair_ce_fwd.  The reference applies the mask through ``.data.masked_fill_``
    (DAMSM_losses.py:163, 229), which autograd does not see: the gradient is ``g * scale`` on
    EVERY cell, masked or not."""

    @staticmethod
    def forward(ctx, scores, scale, class_ids):
        L = _lib.lib()
        B = scores.shape[0]
        out = torch.empty_like(scores)
        lse = torch.empty(2, B, dtype=torch.float32, device=scores.device)
        with torch.cuda.device(scores.device):
            _lib.check(L.eegan_pair_ce_fwd(_lib.ptr(scores), scale, _lib.ptr(class_ids), None, B, _lib.ptr(out), None,
                                           _lib.ptr(lse), _lib.stream_ptr()), "pair_ce_fwd(mask)")
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None, None


class _SentScoresFn(torch.autograd.Function):
    """gamma3 * cos(cnn_i, rnn_j) for all i, j.  Kernels: eegan_sent_scores_fwd / _bwd."""

    @staticmethod
    def forward(ctx, cnn, rnn, g3, eps):
        L = _lib.lib()
        B, D = cnn.shape
        scores = torch.empty(B, B, dtype=torch.float32, device=cnn.device)
        norms = torch.empty(2, B, dtype=torch.float32, device=cnn.device)
        with torch.cuda.device(cnn.device):
            _lib.check(L.eegan_sent_scores_fwd(_lib.ptr(cnn), _lib.ptr(rnn), B, D, g3, eps, _lib.ptr(scores),
                                               _lib.ptr(norms), _lib.stream_ptr()), "sent_scores_fwd")
        ctx.save_for_backward(cnn, rnn, norms)
        ctx.c = (g3, eps)
        return scores

    @staticmethod
    def backward(ctx, g):
        cnn, rnn, norms = ctx.saved_tensors
        L = _lib.lib()
        B, D = cnn.shape
        d_cnn, d_rnn = torch.empty_like(cnn), torch.empty_like(rnn)
        g = _lib.f32c(g)
        with torch.cuda.device(cnn.device):
            _lib.check(L.eegan_sent_scores_bwd(_lib.ptr(cnn), _lib.ptr(rnn), _lib.ptr(norms), _lib.ptr(g), B, D,
                                               ctx.c[0], ctx.c[1], _lib.ptr(d_cnn), _lib.ptr(d_rnn),
                                               _lib.stream_ptr()), "sent_scores_bwd")
        return d_cnn, d_rnn, None, None


class _SentLossFn(torch.autograd.Function):
    """sent_loss (DAMSM_losses.py:233-270) as ONE autograd node: cosine scores + class mask + two-way CE forward,
    CE backward + score backward.  Four launches each way and a single round of autograd bookkeeping (the function
    is called twice per training step, train.py:425, 432, and is pure launch latency)."""

    @staticmethod
    def forward(ctx, cnn, rnn, g3, eps, cls, labels):
        L = _lib.lib()
        B, D = cnn.shape
        dev = cnn.device
        scores = torch.empty(B, B, dtype=torch.float32, device=dev)
        out = torch.empty_like(scores)
        small = torch.empty(4 * B + 2, dtype=torch.float32, device=dev)  # norms [2,B] | lse [2,B] | loss01 [2]
        norms, lse, loss01 = small[:2 * B], small[2 * B:4 * B], small[4 * B:]
        p, st = _lib.ptr, _lib.stream_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.eegan_sent_scores_fwd(p(cnn), p(rnn), B, D, g3, eps, p(scores), p(norms), st), "sent_scores_fwd")
            _lib.check(L.eegan_pair_ce_fwd(p(scores), 1.0, p(cls), p(labels), B, p(out), p(loss01), p(lse), st), "pair_ce_fwd")
        ctx.save_for_backward(cnn, rnn, out, small, labels)
        ctx.c = (g3, eps)
        return loss01[0], loss01[1]

    @staticmethod
    def backward(ctx, g0, g1):
        cnn, rnn, out, small, labels = ctx.saved_tensors
        L = _lib.lib()
        B, D = cnn.shape
        norms, lse = small[:2 * B], small[2 * B:4 * B]
        z = None
        if g0 is None or g1 is None:
            z = torch.zeros((), dtype=torch.float32, device=out.device)
        g = torch.stack([(g0 if g0 is not None else z).float().reshape(()), (g1 if g1 is not None else z).float().reshape(())])
        ds = torch.empty_like(out)
        d_cnn = torch.empty_like(cnn) if ctx.needs_input_grad[0] else None
        d_rnn = torch.empty_like(rnn) if ctx.needs_input_grad[1] else None
        p, st = _lib.ptr, _lib.stream_ptr()
        with torch.cuda.device(out.device):
            _lib.check(L.eegan_pair_ce_bwd(p(out), p(lse), p(labels), p(g), 1.0, B, p(ds), st), "pair_ce_bwd")
            tmp_c = d_cnn if d_cnn is not None else torch.empty_like(cnn)
            tmp_r = d_rnn if d_rnn is not None else torch.empty_like(rnn)
            _lib.check(L.eegan_sent_scores_bwd(p(cnn), p(rnn), p(norms), p(ds), B, D, ctx.c[0], ctx.c[1], p(tmp_c), p(tmp_r), st),
                       "sent_scores_bwd")
        return d_cnn, d_rnn, None, None, None, None


class _GagFn(torch.autograd.Function):
    """GlobalAttentionGeneral.forward.  Kernels: eegan_gag_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, key, value, mask_u8, mask_mode):
        L = _lib.lib()
        B, idf, Q = x.shape
        T = key.shape[2]
        out = torch.empty_like(x)
        attn = torch.empty(B, T, Q, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.eegan_gag_fwd(_lib.ptr(x), _lib.ptr(key), _lib.ptr(value), _lib.ptr(mask_u8), mask_mode,
                                       B, idf, Q, T, _lib.ptr(out), _lib.ptr(attn), _lib.stream_ptr()), "gag_fwd")
        ctx.save_for_backward(x, key, value, attn)
        ctx.set_materialize_grads(False)  # an unused output arrives as None (the kernels take NULL), not as a zero tensor to read
        return out, attn

    @staticmethod
    def backward(ctx, d_out, d_attn):
        x, key, value, attn = ctx.saved_tensors
        L = _lib.lib()
        B, idf, Q = x.shape
        T = key.shape[2]
        d_out = _lib.f32c(d_out) if d_out is not None else None
        d_attn = _lib.f32c(d_attn) if d_attn is not None else None
        d_x, d_key, d_val = torch.empty_like(x), torch.empty_like(key), torch.empty_like(value)
        need = L.eegan_gag_bwd_workspace_bytes(B, idf, Q, T)  # 0: shape outside the two-kernel form -> one-kernel fallback
        ws = torch.empty(need, dtype=torch.uint8, device=x.device) if need else None
        with torch.cuda.device(x.device):
            _lib.check(L.eegan_gag_bwd_ws(_lib.ptr(x), _lib.ptr(key), _lib.ptr(value), _lib.ptr(attn), _lib.ptr(d_out),
                                          _lib.ptr(d_attn), B, idf, Q, T, _lib.ptr(d_x), _lib.ptr(d_key), _lib.ptr(d_val),
                                          _lib.ptr(ws), need, _lib.stream_ptr()), "gag_bwd")
        return d_x, d_key, d_val, None, None


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
def _device_i64(v, device, n=None):
    """class_ids / labels may arrive as a CPU LongTensor, numpy array or list
    (train.py:423 builds class_ids with torch.LongTensor on the host)."""
    if v is None:
        return None
    t = torch.as_tensor(v)
    if n is not None:
        t = t.reshape(-1)[:n]
    return t.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


def _slice_att_maps(att, lens, hw):
    H, W = hw
    return [att[i:i + 1, :lens[i]].reshape(1, lens[i], H, W).contiguous() for i in range(att.shape[0])]


class _LazyAttMaps(Sequence):
    """``att_maps`` as the reference returns it — [1, T_i, H, W] tensors, one per caption
    (DAMSM_losses.py:301) — sliced out of the kernel's [B, T_max, R] buffer only when first
    touched: slicing needs cap_lens on the host, and the reference's ``cap_lens.data.tolist()``
    sync (:280) is the one thing train.py never needs (it drops att_maps, train.py:428).

    Deliberately NOT a ``list`` subclass: C-level consumers (``torch.cat``, ``list.copy``, pickling)
    read a list's storage directly and would see an empty one before the first Python-level
    access.  As a plain Sequence they either go through ``__getitem__`` / ``__iter__`` (and get
    the tensors) or raise a TypeError; ``list(att_maps)`` or ``att_maps.tolist()`` gives the
    reference's exact type.  When ``cap_lens`` is already on the host, ``_att_maps()`` returns a
    real list straight away and this class is not used."""

    def __init__(self, att, cap_lens, hw):
        self._src = (att, cap_lens, hw)
        self._items = None

    def _fill(self):
        if self._items is None:
            att, cap_lens, hw = self._src
            self._items = _slice_att_maps(att, [int(v) for v in cap_lens.reshape(-1).tolist()], hw)
            self._src = None
        return self._items

    def tolist(self):
        return list(self._fill())

    def __len__(self):
        return self._src[0].shape[0] if self._items is None else len(self._items)

    def __getitem__(self, k):
        return self._fill()[k]

    def __iter__(self):
        return iter(self._fill())

    def __add__(self, other):
        return self.tolist() + list(other)

    def __radd__(self, other):
        return list(other) + self.tolist()

    def __reduce__(self):
        return (list, (self.tolist(),))

    def __repr__(self):
        return repr(self._fill())


def _att_maps(att, cap_lens, hw):
    """A real list when the caption lengths are host data (no sync needed), the lazy Sequence otherwise."""
    if att is None or att.numel() == 0:
        return []
    if not (torch.is_tensor(cap_lens) and cap_lens.is_cuda):
        lens = [int(v) for v in torch.as_tensor(cap_lens).reshape(-1).tolist()]
        return _slice_att_maps(att, lens, hw)
    return _LazyAttMaps(att, cap_lens, hw)


def pair_grid(img_features, words_emb, cap_lens, diag_offset=0, want_att=True):
    """Un-scaled similarity grid m [B_img, B_cap] and diagonal attention [B_cap, T_max, R].
    Building block shared by words_similarity / words_loss and the sharded multi-GPU path."""
    _lib.require_cuda(img_features, words_emb)
    g1, g2, _ = gammas()
    Bi, D = img_features.shape[0], img_features.shape[1]
    img = _lib.f32c(img_features).reshape(Bi, D, -1)
    words = _lib.f32c(words_emb)
    lens = torch.as_tensor(cap_lens).reshape(-1)[: words.shape[0]]
    lens32 = lens.to(device=img.device, dtype=torch.int32, non_blocking=True).contiguous()
    m, att = _PairGridFn.apply(img, words, lens32, g1, g2, int(diag_offset), bool(want_att))
    return m, att


def _spatial(img_features):
    if img_features.dim() == 4:
        return img_features.shape[2], img_features.shape[3]
    r = img_features.shape[2]
    h = int(round(math.sqrt(r)))
    return (h, r // h)


# ---------------------------------------------------------------------------------------
# public API (reference signatures)
# ---------------------------------------------------------------------------------------
def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """DAMSM_losses.py:17-23 — cosine similarity along ``dim``."""
    from .attention import cosine_rows
    return cosine_rows(x1, x2, dim, eps)


def func_attention(query, context, gamma1):
    """DAMSM_losses.py:25-63 — query [B,D,T], context [B,D,H,W] ->
    (weightedContext [B,D,T], attn [B,T,H,W])."""
    from .attention import func_attention as _fa
    return _fa(query, context, gamma1)


class GlobalAttentionGeneral(nn.Module):
    """DAMSM_losses.py:65-132.  ``idf``/``cdf`` are accepted and unused exactly as in the
    reference (its 1x1 ``conv_context`` is commented out, :68).  ``mask_mode="reference"``
    reproduces the reference's row/mask pairing (row (b,q) uses mask[(b*Q+q) % B],
    :114-118); ``"intended"`` pairs row (b,q) with mask[b]."""

    def __init__(self, idf, cdf, mask_mode="reference"):
        super().__init__()
        self.sm = nn.Softmax(dim=1)  # kept for attribute parity; the kernel does the softmax
        self.mask = None
        if mask_mode not in ("reference", "intended"):
            raise ValueError("mask_mode must be 'reference' or 'intended'")
        self.mask_mode = mask_mode

    def applyMask(self, mask):
        self.mask = mask  # batch x sourceL

    def forward(self, input, context_key, content_value):
        _lib.require_cuda(input, context_key, content_value)
        B, idf = input.shape[0], input.shape[1]
        ih, iw = input.size(2), input.size(3)
        x = _lib.f32c(input).reshape(B, idf, ih * iw)
        key, val = _lib.f32c(context_key), _lib.f32c(content_value)
        mask = None
        if self.mask is not None:
            mask = self.mask.to(device=x.device).contiguous()
            # a bool mask IS one byte of 0 / 1 per word: reinterpret it (no conversion kernel per call)
            mask = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
        if idf > _GAG_BWD_MAX_IDF and torch.is_grad_enabled() and (x.requires_grad or key.requires_grad or val.requires_grad):
            raise RuntimeError("GlobalAttentionGeneral: idf=%d > %d has a forward kernel but no backward; "
                               "run it under torch.no_grad()" % (idf, _GAG_BWD_MAX_IDF))
        out, attn = _GagFn.apply(x, key, val, mask, 0 if self.mask_mode == "reference" else 1)
        return out.view(B, -1, ih, iw), attn.view(B, -1, ih, iw)


def sent_similarity(cnn_code, rnn_code, class_ids, batch_size, eps=1e-8):
    """DAMSM_losses.py:134-166 -> scores [B,B] (gamma3-scaled, same-class cells -inf)."""
    _lib.require_cuda(cnn_code, rnn_code)
    if cnn_code.dim() == 3:  # seq_len x B x nef with seq_len == 1 (:149-151, :162)
        cnn_code, rnn_code = cnn_code.squeeze(0), rnn_code.squeeze(0)
    _, _, g3 = gammas()
    scores = _SentScoresFn.apply(_lib.f32c(cnn_code), _lib.f32c(rnn_code), g3, float(eps))
    cls = _device_i64(class_ids, scores.device, batch_size)
    if cls is None:
        return scores
    return _ScaleMaskFn.apply(scores, 1.0, cls)


# Set by ``eegan_b200.install(distributed=...)``: with one process per GPU and an initialised process group the
# reference's call sites (train.py:425-432) hand these functions the RANK's shard, while the reference evaluates the
# loss on the full batch gathered on GPU 0 (train.py:195).  When True, words_loss / sent_loss called under a group of
# more than one rank evaluate the GLOBAL-batch loss from the shards (eegan_b200.sharded).
AUTO_SHARD = False


def _auto_shard():
    if not AUTO_SHARD:
        return False
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8):
    """DAMSM_losses.py:233-270 -> (loss0, loss1)."""
    _lib.require_cuda(cnn_code, rnn_code)
    if labels is not None and _auto_shard() and not getattr(_tls, "inside_sharded", False):
        from . import sharded
        _tls.inside_sharded = True
        try:
            return sharded.sharded_sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps)
        finally:
            _tls.inside_sharded = False
    if cnn_code.dim() == 3:
        cnn_code, rnn_code = cnn_code.squeeze(0), rnn_code.squeeze(0)
    if labels is None:
        return None, None
    _, _, g3 = gammas()
    cls = _device_i64(class_ids, cnn_code.device, batch_size)
    lab = _device_i64(labels, cnn_code.device)
    return _SentLossFn.apply(_lib.f32c(cnn_code), _lib.f32c(rnn_code), g3, float(eps), cls, lab)


def words_similarity(img_features, words_emb, cap_lens, class_ids, batch_size):
    """DAMSM_losses.py:168-231 -> (similarities [B,B] gamma3-scaled and class-masked, att_maps)."""
    _, _, g3 = gammas()
    m, att = pair_grid(img_features[:batch_size], words_emb[:batch_size], cap_lens)
    cls = _device_i64(class_ids, m.device, batch_size)
    sim = _ScaleMaskFn.apply(m, g3, cls)
    return sim, _att_maps(att, cap_lens, _spatial(img_features))


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size):
    """DAMSM_losses.py:272-342 -> (loss0, loss1, att_maps).

    Default route: one autograd node over a per-shape plan with captured forward / backward graphs
    (eegan_b200/fastpath.py); the two-node route below it serves labels=None, the validation engines and
    EEGAN_WORDS_LOSS_PLAN=0."""
    from . import fastpath
    if _auto_shard() and not getattr(_tls, "inside_sharded", False):
        from . import sharded
        _tls.inside_sharded = True
        try:
            return sharded.sharded_words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size)
        finally:
            _tls.inside_sharded = False
    if labels is not None and torch.is_tensor(img_features) and torch.is_tensor(words_emb):
        _lib.require_cuda(img_features, words_emb)
        img_b = img_features if img_features.shape[0] == batch_size else img_features[:batch_size]
        words_b = words_emb if words_emb.shape[0] == batch_size else words_emb[:batch_size]
        if fastpath.supported(img_b, words_b):
            Bi, D = img_b.shape[0], img_b.shape[1]
            img3 = _lib.f32c(img_b).view(Bi, D, -1)
            cls = class_ids if (class_ids is None or torch.is_tensor(class_ids)) else torch.as_tensor(class_ids)
            lens = cap_lens if torch.is_tensor(cap_lens) else torch.as_tensor(cap_lens)
            lab = labels if torch.is_tensor(labels) else torch.as_tensor(labels)
            loss0, loss1, att = fastpath.words_loss_planned(img3, _lib.f32c(words_b), lens, cls, lab, gammas())
            return loss0, loss1, _att_maps(att, cap_lens, _spatial(img_features))
    _, _, g3 = gammas()
    m, att = pair_grid(img_features[:batch_size], words_emb[:batch_size], cap_lens)
    att_maps = _att_maps(att, cap_lens, _spatial(img_features))
    if labels is None:
        return None, None, att_maps
    cls = _device_i64(class_ids, m.device, batch_size)
    lab = _device_i64(labels, m.device)
    loss0, loss1, _ = _PairCEFn.apply(m, g3, cls, lab)
    return loss0, loss1, att_maps

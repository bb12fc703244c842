"""TEST INFRASTRUCTURE ONLY — CPU restatement of EE-GAN's DAMSM / word-region attention math.

Follows (file:line relative to the reference repo qikizh/EE-GAN):
  miscc/DAMSM_losses.py:17-23   cosine_similarity
  miscc/DAMSM_losses.py:25-63   func_attention            (two-stage softmax attention)
  miscc/DAMSM_losses.py:65-132  GlobalAttentionGeneral    (masked word softmax per pixel)
  miscc/DAMSM_losses.py:134-166 sent_similarity
  miscc/DAMSM_losses.py:168-231 words_similarity
  miscc/DAMSM_losses.py:233-270 sent_loss
  miscc/DAMSM_losses.py:272-342 words_loss
  miscc/config.py:47-51         gamma1=5, gamma2=5, gamma3=10

Two independent restatements are kept so that they check each other:
  * ``port_*``  — per-caption loop over batched attentions, fp32 torch ops in the same
    order of operations as the reference.  This is also the CPU arm ``bench.py`` times
    (``cpu_baseline.kind == "port"``): same op mix (B iterations of bmm / softmax /
    transposed copies) and therefore the same cost profile as the reference's own code.
  * ``dense_*`` — all B x B pairs at once with a padding mask, dtype-generic (run it in
    float64 for a high-precision oracle), plus ``dense_words_backward`` which is the
    hand-derived backward the CUDA kernels implement, checked against
    autograd in tests/test_oracle.py.

Parity pin: checked against the live reference and against tests/golden/*.npz (generated
from the live reference by oracle/make_golden.py).  Never imported by the product.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

GAMMA1 = 5.0  # miscc/config.py:48
GAMMA2 = 5.0  # miscc/config.py:50
GAMMA3 = 10.0  # miscc/config.py:49


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def class_mask(class_ids, batch_size: int) -> Optional[torch.Tensor]:
    """mask[a,b] = class[a]==class[b] and a!=b   (DAMSM_losses.py:282-285, 325-327)."""
    if class_ids is None:
        return None
    c = torch.as_tensor(class_ids).reshape(-1)[:batch_size].cpu()
    m = c.view(-1, 1) == c.view(1, -1)
    m.fill_diagonal_(False)
    return m


def _two_way_ce(scores: torch.Tensor, labels: Optional[torch.Tensor]):
    """CE over rows and over columns (DAMSM_losses.py:264-267, 335-338)."""
    if labels is None:
        return None, None
    return F.cross_entropy(scores, labels), F.cross_entropy(scores.t(), labels)


# --------------------------------------------------------------------------------------
# port_*: loop-structured fp32 restatement (also the timed CPU arm)
# --------------------------------------------------------------------------------------
def port_cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """DAMSM_losses.py:17-23."""
    num = (x1 * x2).sum(dim)
    den = (x1.norm(2, dim) * x2.norm(2, dim)).clamp(min=eps)
    return (num / den).squeeze()


def port_func_attention(query, context, gamma1):
    """DAMSM_losses.py:25-63.  query [B,D,T], context [B,D,H,W] -> ([B,D,T], [B,T,H,W])."""
    B, D, T = query.shape
    H, W = context.shape[2], context.shape[3]
    R = H * W
    ctx = context.reshape(B, D, R)
    s = torch.bmm(ctx.transpose(1, 2).contiguous(), query)  # [B,R,T]            :42
    p = torch.softmax(s.reshape(B * R, T), dim=1).reshape(B, R, T)  # over words   :44-45
    z = p.transpose(1, 2).contiguous().reshape(B * T, R) * gamma1  #               :50-53
    a = torch.softmax(z, dim=1).reshape(B, T, R)  # over regions                   :54-55
    u = torch.bmm(ctx, a.transpose(1, 2).contiguous())  # [B,D,T]                  :61
    return u, a.reshape(B, T, H, W)


def _port_similarity_grid(img_features, words_emb, cap_lens, batch_size, g1, g2):
    """The caption loop of DAMSM_losses.py:281-321 -> sim [B_img, B_cap] (un-scaled), att_maps."""
    lens = [int(v) for v in torch.as_tensor(cap_lens).reshape(-1).tolist()]
    cols, att_maps = [], []
    for i in range(batch_size):
        T = lens[i]
        w = words_emb[i : i + 1, :, :T].expand(batch_size, -1, -1).contiguous()  # :289-291
        u, attn = port_func_attention(w, img_features, g1)  # :300
        att_maps.append(attn[i : i + 1].contiguous())  # :301
        wt = w.transpose(1, 2).reshape(batch_size * T, -1)
        ut = u.transpose(1, 2).reshape(batch_size * T, -1)
        cos = port_cosine_similarity(wt, ut).reshape(batch_size, T)  # :310-312
        cols.append(torch.log(torch.exp(cos * g2).sum(dim=1, keepdim=True)))  # :315-317
    return torch.cat(cols, dim=1), att_maps  # :324


def port_words_similarity(img_features, words_emb, cap_lens, class_ids, batch_size,
                          g1=GAMMA1, g2=GAMMA2, g3=GAMMA3):
    """DAMSM_losses.py:168-231."""
    sim, att_maps = _port_similarity_grid(img_features, words_emb, cap_lens, batch_size, g1, g2)
    sim = sim * g3
    m = class_mask(class_ids, batch_size)
    if m is not None:
        sim = sim.masked_fill(m.to(sim.device), -math.inf)
    return sim, att_maps


def port_words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size,
                    g1=GAMMA1, g2=GAMMA2, g3=GAMMA3):
    """DAMSM_losses.py:272-342 -> (loss0, loss1, att_maps)."""
    sim, att_maps = port_words_similarity(img_features, words_emb, cap_lens, class_ids,
                                          batch_size, g1, g2, g3)
    l0, l1 = _two_way_ce(sim, labels)
    return l0, l1, att_maps


def port_sent_similarity(cnn_code, rnn_code, class_ids, batch_size, eps=1e-8, g3=GAMMA3):
    """DAMSM_losses.py:134-166 (2-D inputs)."""
    nc = cnn_code.norm(2, dim=1, keepdim=True)
    nr = rnn_code.norm(2, dim=1, keepdim=True)
    scores = cnn_code @ rnn_code.t() / (nc @ nr.t()).clamp(min=eps) * g3
    m = class_mask(class_ids, batch_size)
    if m is not None:
        scores = scores.masked_fill(m.to(scores.device), -math.inf)
    return scores


def port_sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8, g3=GAMMA3):
    """DAMSM_losses.py:233-270 -> (loss0, loss1)."""
    return _two_way_ce(port_sent_similarity(cnn_code, rnn_code, class_ids, batch_size, eps, g3), labels)


def port_global_attention(x, key, value, mask=None, mask_mode="reference"):
    """GlobalAttentionGeneral.forward, DAMSM_losses.py:75-132.

    x [B,idf,H,W], key/value [B,idf,T], mask [B,T] bool (True = padding) or None.
    mask_mode "reference": row (b,q) is masked with mask[(b*Q+q) % B]  (the
    ``mask.repeat(queryL, 1)`` at :117 against batch-major rows at :114 — SURVEY.md D8).
    mask_mode "intended": row (b,q) is masked with mask[b].
    Returns (weightedContext [B,idf,H,W], attn [B,T,H,W]).
    """
    B, idf, H, W = x.shape
    Q = H * W
    T = key.shape[2]
    s = torch.bmm(x.reshape(B, idf, Q).transpose(1, 2), key)  # [B,Q,T]   :96
    if mask is not None:
        if mask_mode == "reference":
            rows = (torch.arange(B * Q, device=x.device) % B)
            mrow = mask.bool()[rows].reshape(B, Q, T)
        else:
            mrow = mask.bool()[:, None, :].expand(B, Q, T)
        s = s.masked_fill(mrow, -math.inf)  # :118 (autograd-transparent in the reference)
    p = torch.softmax(s, dim=2)  # :119
    pt = p.transpose(1, 2)  # [B,T,Q]   :123
    out = torch.bmm(value, pt)  # [B,idf,Q] :127
    return out.reshape(B, idf, H, W), pt.reshape(B, T, H, W)


# --------------------------------------------------------------------------------------
# dense_*: all pairs at once, dtype-generic (float64 for the high-precision oracle)
# --------------------------------------------------------------------------------------
def _len_mask(cap_lens, T, device):
    lens = torch.as_tensor(cap_lens, device=device).reshape(-1)
    return torch.arange(T, device=device)[None, :] < lens[:, None]  # [B,T] True = valid


def dense_pair_terms(img, words, cap_lens, g1=GAMMA1, g2=GAMMA2):
    """All intermediates of SURVEY.md App. A for the full grid.

    img [Bi,D,R] or [Bi,D,H,W]; words [Bc,D,T]; returns dict with
      s,p [Bi,Bc,R,T]; a [Bi,Bc,T,R]; u [Bi,Bc,D,T]; cos [Bi,Bc,T]; m [Bi,Bc] (un-scaled sim).
    Padded words (t >= cap_lens[i]) carry p=0, a=0, cos excluded from the LSE.
    """
    Bi, D = img.shape[0], img.shape[1]
    c = img.reshape(Bi, D, -1)
    Bc, _, T = words.shape
    valid = _len_mask(cap_lens, T, img.device)  # [Bc,T]
    s = torch.einsum("jdr,idt->jirt", c, words)
    s_m = s.masked_fill(~valid[None, :, None, :], -math.inf)
    p = torch.softmax(s_m, dim=3)  # over words
    a = torch.softmax(g1 * p.transpose(2, 3), dim=3)  # [Bi,Bc,T,R] over regions
    a = a * valid[None, :, :, None]
    u = torch.einsum("jdr,jitr->jidt", c, a)
    wn = words.norm(2, dim=1)  # [Bc,T]
    un = u.norm(2, dim=2)  # [Bi,Bc,T]
    dot = torch.einsum("idt,jidt->jit", words, u)
    cos = dot / (wn[None] * un).clamp(min=1e-8)
    e = torch.exp(g2 * cos) * valid[None]
    m = torch.log(e.sum(dim=2))
    return dict(s=s, p=p, a=a, u=u, cos=cos, m=m, wn=wn, un=un, valid=valid, c=c)


def dense_words_similarity(img, words, cap_lens, class_ids, g1=GAMMA1, g2=GAMMA2, g3=GAMMA3):
    t = dense_pair_terms(img, words, cap_lens, g1, g2)
    sim = t["m"] * g3
    m = class_mask(class_ids, sim.shape[0])
    if m is not None:
        sim = sim.masked_fill(m.to(sim.device), -math.inf)
    return sim, t


def dense_words_loss(img, words, labels, cap_lens, class_ids, g1=GAMMA1, g2=GAMMA2, g3=GAMMA3):
    sim, t = dense_words_similarity(img, words, cap_lens, class_ids, g1, g2, g3)
    l0, l1 = _two_way_ce(sim, labels)
    H = int(round(math.sqrt(t["c"].shape[2])))
    lens = torch.as_tensor(cap_lens).reshape(-1).tolist()
    att = [t["a"][i, i, : int(lens[i])].reshape(1, int(lens[i]), H, -1) for i in range(sim.shape[0])]
    return l0, l1, att, sim


def ce_pair_grad(sim, labels, g_loss0=1.0, g_loss1=1.0):
    """d(loss0*g0 + loss1*g1)/dsim for the two-way CE; -inf cells get 0."""
    B = sim.shape[0]
    onehot = F.one_hot(labels, B).to(sim.dtype)
    row = (torch.softmax(sim, dim=1) - onehot) / B  # loss0: rows, softmax over columns
    col = (torch.softmax(sim.t(), dim=1) - onehot).t() / B  # loss1: CE(sim^T)
    g = g_loss0 * row + g_loss1 * col
    return torch.where(torch.isinf(sim), torch.zeros_like(g), g)


def dense_words_backward(img, words, cap_lens, dm, g1=GAMMA1, g2=GAMMA2):
    """Hand-derived backward of m[j,i]; dm = dL/dm [Bi,Bc] (already
    includes gamma3 and zeros at masked cells).  Returns (d_img like img, d_words)."""
    t = dense_pair_terms(img, words, cap_lens, g1, g2)
    c, p, a, u, cos, wn, un, valid = (t[k] for k in ("c", "p", "a", "u", "cos", "wn", "un", "valid"))
    e = torch.exp(g2 * cos) * valid[None]
    dcos = dm[:, :, None] * g2 * e / e.sum(dim=2, keepdim=True)  # [Bi,Bc,T]
    nn_ = wn[None] * un  # [Bi,Bc,T]
    live = (nn_ > 1e-8).to(c.dtype) * valid[None]
    inv = live / nn_.clamp(min=1e-30)
    w_b = words[None]  # [1,Bc,D,T]
    du = dcos[:, :, None, :] * (w_b * inv[:, :, None, :]
                                - (cos * live / un.clamp(min=1e-30) ** 2)[:, :, None, :] * u)
    dw = dcos[:, :, None, :] * (u * inv[:, :, None, :]
                                - (cos * live / wn[None].clamp(min=1e-30) ** 2)[:, :, None, :] * w_b)
    da = torch.einsum("jidt,jdr->jitr", du, c)
    dc = torch.einsum("jidt,jitr->jdr", du, a)
    dz = a * (da - (a * da).sum(dim=3, keepdim=True))
    dp = g1 * dz.transpose(2, 3)  # [Bi,Bc,R,T]
    ds = p * (dp - (p * dp).sum(dim=3, keepdim=True))
    dc = dc + torch.einsum("jirt,idt->jdr", ds, words)
    dw = dw.sum(dim=0) + torch.einsum("jirt,jdr->idt", ds, c)
    return dc.reshape(img.shape), dw


def dense_sent_scores(cnn_code, rnn_code, class_ids, eps=1e-8, g3=GAMMA3):
    return port_sent_similarity(cnn_code, rnn_code, class_ids, cnn_code.shape[0], eps, g3)


# --------------------------------------------------------------------------------------
# SyncBatchNorm statistics (sync_batchnorm/batchnorm.py:48-78, 113-125)
# --------------------------------------------------------------------------------------
def syncbn_forward(x_shards: Sequence[torch.Tensor], weight, bias, eps=1e-5):
    """N-replica SyncBN forward: per-replica sum / square-sum, reduced, then
    mean = S/N; sumvar = SS - S*mean; inv_std = clamp(sumvar/N, eps)^-0.5   (:113-125).
    Returns (list of outputs, mean, inv_std, unbiased_var)."""
    C = x_shards[0].shape[1]
    tot = sum(x.numel() // C for x in x_shards)
    s = sum(x.transpose(0, 1).reshape(C, -1).sum(1) for x in x_shards)
    ss = sum((x.transpose(0, 1).reshape(C, -1) ** 2).sum(1) for x in x_shards)
    mean = s / tot
    sumvar = ss - s * mean
    inv_std = (sumvar / tot).clamp(min=eps) ** -0.5
    shape = [1, C] + [1] * (x_shards[0].dim() - 2)
    outs = []
    for x in x_shards:
        if weight is not None:  # fused scale, batchnorm.py:73
            y = (x - mean.view(shape)) * (inv_std * weight).view(shape) + bias.view(shape)
        else:
            y = (x - mean.view(shape)) * inv_std.view(shape)
        outs.append(y)
    return outs, mean, inv_std, sumvar / (tot - 1)


def port_rprecision(cnn_code, rnn_codes, eps=1e-8):
    """test.py:306-336 (Tester.cal_sim_one_by_one), the arithmetic of its per-sample loop body
    (:323-330) in the reference's own op order: one image against its R_val candidate sentence
    codes, candidate 0 = the matching caption.  Returns (hits [B] bool, argmax [B], scores0 [B,R_val])."""
    B = cnn_code.shape[0]
    hits, best, rows = [], [], []
    for ix in range(B):
        rnn_code = rnn_codes[ix]
        scores = torch.mm(cnn_code[ix].unsqueeze(0), rnn_code.transpose(0, 1))  # 1 x R_val   (:323)
        cnn_code_norm = torch.norm(cnn_code[ix].unsqueeze(0), 2, dim=1, keepdim=True)  # (:324)
        rnn_code_norm = torch.norm(rnn_code, 2, dim=1, keepdim=True)  # (:325)
        norm = torch.mm(cnn_code_norm, rnn_code_norm.transpose(0, 1))  # (:326)
        scores0 = scores / norm.clamp(min=eps)  # (:327)
        am = int(torch.argmax(scores0))  # (:329)
        hits.append(am == 0)
        best.append(am)
        rows.append(scores0[0])
    return torch.tensor(hits), torch.tensor(best), torch.stack(rows)


def port_affine_ssa(feat, weight, bias, semi_mask, eps=1e-5, n_replica_formula=False):
    """models.py:68-86 (affine_ssa.forward) after its two MLPs: ``weight`` / ``bias`` are the outputs
    of fc_gamma / fc_beta [N,C].  norm2d is SyncBN(affine=False) in training mode: batch statistics
    over (N,H,W); a single replica goes through F.batch_norm (sync_batchnorm/batchnorm.py:50-53,
    1/sqrt(var+eps)), several replicas through the clamp formula (:113-125)."""
    mean = feat.mean(dim=(0, 2, 3), keepdim=True)
    var = ((feat - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    inv_std = var.clamp(min=eps) ** -0.5 if n_replica_formula else (var + eps) ** -0.5
    xhat = (feat - mean) * inv_std  # :69
    size = xhat.size()
    w = weight.unsqueeze(-1).unsqueeze(-1).expand(size)  # :79-81
    b = bias.unsqueeze(-1).unsqueeze(-1).expand(size)
    w = w * semi_mask + 1  # :84
    b = b * semi_mask  # :85
    return w * xhat + b  # :86


def port_attr_enhance(sent, attrs, Wq, bq, Wk, bk, Wv, bv, norm_fact):
    """models.py:154-169 (ATTR_Enhance.forward), op for op: cat, three Linear, softmax(q k^T, -1) * norm_fact
    (the scale comes AFTER the softmax, :166), bmm with v.  Returns (attn_sent, attn_attrs).  dtype-generic."""
    import torch.nn.functional as F
    combine = torch.cat([sent.unsqueeze(1), attrs], dim=1)
    q = F.linear(combine, Wq, bq)
    k = F.linear(combine, Wk, bk)
    v = F.linear(combine, Wv, bv)
    attn = torch.softmax(torch.bmm(q, k.permute(0, 2, 1)), dim=-1) * norm_fact
    out = torch.bmm(attn, v)
    return out[:, 0, :], out


def port_emb_features(x, weight):
    """DAMSM.py:23-26, 229: conv1x1(768, nef) without bias = a per-pixel matrix product over the channels."""
    import torch.nn.functional as F
    return F.conv2d(x, weight.reshape(weight.shape[0], -1, 1, 1))

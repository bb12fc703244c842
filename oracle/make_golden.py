"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

The reference (qikizh/EE-GAN) has no golden vectors of its own, so these
fixtures — outputs of miscc/DAMSM_losses.py and sync_batchnorm/batchnorm.py on seeded
inputs from oracle/cases.py — are what pins parity.  Inputs are stored in full for tiny
cases, and as (recipe, checksum) for CUB-shaped ones (regenerated from the seed on use;
the checksum catches RNG drift).  Gradients of large tensors are stored on a strided
subset plus a checksum to keep the fixtures small.
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import torch

from . import cases
from .ref_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# loss = W0*loss0 + W1*loss1 so that the two CE directions are distinguishable in the grads
W0, W1 = 1.0, 2.0
IMG_GRAD_STRIDE = (4, 5)  # d_img[:, ::4, ::5] (flattened regions)

WORDS_CASES = {
    # name: kwargs for cases.words_case
    "words_tiny": dict(B=4, T=6, D=32, H=3, kind="stress", class_mode="cub", min_len=2),
    "words_cub6_real": dict(B=6, T=18, kind="realistic", class_mode="cub"),
    "words_cub6_stress": dict(B=6, T=18, kind="stress", class_mode="unique"),
    "words_cub5_t20_nocls": dict(B=5, T=20, kind="realistic", class_mode="none"),
    "words_c1_b16": dict(B=16, T=18, kind="realistic", class_mode="cub"),
}
SENT_CASES = {
    "sent_b16_cub": dict(B=16, class_mode="cub"),
    "sent_b5_nocls": dict(B=5, class_mode="none"),
}
GAG_CASES = {
    "gag_b3_mask_quirk": dict(B=3, idf=32, H=8, T=18, masked=True),
    "gag_b4_nomask": dict(B=4, idf=64, H=16, T=12, masked=False),
    "gag_b5_mask": dict(B=5, idf=128, H=4, T=20, masked=True),
}
BN_CASES = {
    "bn_2x3_c8": dict(N=6, C=8, H=5, shards=2),
    "bn_4x2_c32": dict(N=8, C=32, H=4, shards=4),
}


ATTR_CASES = {
    "attr_b6_d128": dict(B=6, D=128, A=3, seed=11),
    "attr_b3_d64_a5": dict(B=3, D=64, A=5, seed=12),
}
EMB_SUB = (4, 3, 7)  # strides of the stored subsets: output channels, regions, input channels
EMB_CASES = {
    "emb_b2_768_256": dict(B=2, Cin=768, Cout=256, H=17, seed=21),
    "emb_b3_64_32_h5": dict(B=3, Cin=64, Cout=32, H=5, seed=22),
}


def _np(t):
    return t.detach().cpu().numpy()


def _ref_module(name):
    import importlib
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(name)


def gen_attr(ref, name, kw):
    """models.py:146-169 (ATTR_Enhance) executed live with seeded parameters; gradients of
    loss = sum(attn_sent * gs) + sum(attn_attrs * ga) w.r.t. both inputs and all six parameters."""
    models = _ref_module("models")
    g = cases._gen(kw["seed"])
    B, D, A = kw["B"], kw["D"], kw["A"]
    mod = models.ATTR_Enhance(ntf=D)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (1.0 / D ** 0.5 if p.dim() == 2 else 0.1))
    sent = (torch.randn(B, D, generator=g) * 0.7).requires_grad_()
    attrs = (torch.randn(B, A, D, generator=g) * 0.7).requires_grad_()
    gs, ga = torch.randn(B, D, generator=g), torch.randn(B, A + 1, D, generator=g)
    a_sent, a_attrs = mod(sent, attrs)
    ((a_sent * gs).sum() + (a_attrs * ga).sum()).backward()
    d = dict(kind="attr", recipe=json.dumps(kw), sent=_np(sent), attrs=_np(attrs), gs=_np(gs), ga=_np(ga),
             attn_sent=_np(a_sent), attn_attrs=_np(a_attrs), d_sent=_np(sent.grad), d_attrs=_np(attrs.grad),
             merged=_np(models.ATTR_Enhance.attr_merge(a_attrs.detach())))
    for k, v in mod.state_dict().items():
        d["param." + k] = _np(v)
    for k, v in mod.named_parameters():
        d["grad." + k] = _np(v.grad)
    return d


def gen_emb(ref, name, kw):
    """DAMSM.py:23-26 conv1x1(Cin, Cout) executed live (the layer CNN_ENCODER.emb_features, :162, 229)."""
    damsm = _ref_module("DAMSM")
    c = cases.emb_case(kw["B"], kw["Cin"], kw["Cout"], kw["H"], kw["seed"])
    conv = damsm.conv1x1(kw["Cin"], kw["Cout"])
    with torch.no_grad():
        conv.weight.copy_(c["weight"])
    x = c["x"].clone().requires_grad_()
    go = c["go"]
    y = conv(x)
    (y * go).sum().backward()
    dx = x.grad
    dw = conv.weight.grad.reshape(kw["Cout"], kw["Cin"])
    y3 = y.reshape(y.shape[0], y.shape[1], -1)
    return dict(kind="emb", recipe=json.dumps(kw),
                in_checksum=np.array(cases.checksum(c["x"]) + cases.checksum(c["weight"]) + cases.checksum(go)),
                y_sub=_np(y3[:, ::EMB_SUB[0], ::EMB_SUB[1]]), y_checksum=np.array(cases.checksum(y)), y_absmax=np.array(float(y.abs().max())),
                d_weight_sub=_np(dw[::EMB_SUB[0], ::EMB_SUB[2]]), d_weight_checksum=np.array(cases.checksum(dw)),
                d_weight_absmax=np.array(float(dw.abs().max())),
                d_x_sub=_np(dx.reshape(dx.shape[0], dx.shape[1], -1)[:, ::EMB_SUB[2], ::EMB_SUB[1]]),
                d_x_checksum=np.array(cases.checksum(dx)), d_x_absmax=np.array(float(dx.abs().max())))


def gen_words(ref, name, kw):
    c = cases.words_case(**kw)
    img = c["img"].clone().requires_grad_()
    words = c["words"].clone().requires_grad_()
    B = kw["B"]
    l0, l1, att = ref.losses.words_loss(img, words, c["labels"], c["cap_lens"], c["class_ids"], B)
    (W0 * l0 + W1 * l1).backward()
    sim, att2 = ref.losses.words_similarity(c["img"], c["words"], c["cap_lens"], c["class_ids"], B)
    d_img = img.grad.reshape(B, img.shape[1], -1)
    out = dict(
        kind="words", recipe=json.dumps(kw),
        in_checksum=np.array(cases.checksum(c["img"]) + cases.checksum(c["words"])),
        cap_lens=_np(c["cap_lens"]),
        class_ids=_np(c["class_ids"]) if c["class_ids"] is not None else np.array([-1]),
        loss0=_np(l0), loss1=_np(l1), sim=_np(sim),
        att=np.concatenate([_np(a).reshape(-1) for a in att]),
        d_words=_np(words.grad),
        d_img_sub=_np(d_img[:, :: IMG_GRAD_STRIDE[0], :: IMG_GRAD_STRIDE[1]]),
        d_img_checksum=np.array(cases.checksum(d_img)),
        d_img_absmax=np.array(float(d_img.abs().max())),
    )
    if c["img"].numel() < 50_000:
        out["img"] = _np(c["img"])
        out["words"] = _np(c["words"])
        out["d_img"] = _np(img.grad)
    # func_attention on caption 0 against every image (DAMSM_losses.py:25-63)
    T0 = int(c["cap_lens"][0])
    q = c["words"][0:1, :, :T0].repeat(B, 1, 1)
    u, attn = ref.losses.func_attention(q, c["img"], ref.cfg.TRAIN.SMOOTH.GAMMA1)
    out["fa_u"] = _np(u)
    out["fa_attn_argmax_words"] = _np(attn.reshape(B, T0, -1).argmax(dim=1)).astype(np.int32)
    out["fa_attn_sub"] = _np(attn.reshape(B, T0, -1)[:, :, ::5])
    return out


def gen_sent(ref, name, kw):
    c = cases.sent_case(**kw)
    cnn = c["cnn"].clone().requires_grad_()
    rnn = c["rnn"].clone().requires_grad_()
    B = kw["B"]
    l0, l1 = ref.losses.sent_loss(cnn, rnn, c["labels"], c["class_ids"], B)
    (W0 * l0 + W1 * l1).backward()
    scores = ref.losses.sent_similarity(c["cnn"], c["rnn"], c["class_ids"], B)
    return dict(kind="sent", recipe=json.dumps(kw), cnn=_np(c["cnn"]), rnn=_np(c["rnn"]),
                class_ids=_np(c["class_ids"]) if c["class_ids"] is not None else np.array([-1]),
                loss0=_np(l0), loss1=_np(l1), scores=_np(scores),
                d_cnn=_np(cnn.grad), d_rnn=_np(rnn.grad))


def gen_gag(ref, name, kw):
    c = cases.gag_case(**kw)
    x = c["x"].clone().requires_grad_()
    k = c["key"].clone().requires_grad_()
    v = c["value"].clone().requires_grad_()
    mod = ref.losses.GlobalAttentionGeneral(kw["idf"], 256)
    if c["mask"] is not None:
        mod.applyMask(c["mask"])
    out, attn = mod(x, k, v)
    g = cases._gen(99)
    go = torch.randn(out.shape, generator=g)
    ga = torch.randn(attn.shape, generator=g)
    ((out * go).sum() + (attn * ga).sum()).backward()
    d = dict(kind="gag", recipe=json.dumps(kw),
             in_checksum=np.array(cases.checksum(c["x"]) + cases.checksum(c["key"])),
             mask=_np(c["mask"]).astype(np.uint8) if c["mask"] is not None else np.array([255], np.uint8),
             out=_np(out), attn=_np(attn),
             attn_argmax=_np(attn.reshape(attn.shape[0], attn.shape[1], -1).argmax(dim=1)).astype(np.int32),
             d_x=_np(x.grad), d_key=_np(k.grad), d_value=_np(v.grad))
    return d


def gen_bn(ref, name, kw):
    """The N-replica statistics math of sync_batchnorm/batchnorm.py:56-78,113-125, driven
    through the reference module's own _compute_mean_std (the thread rendezvous around it
    needs >1 CUDA device and cannot run here)."""
    c = cases.bn_case(kw["N"], kw["C"], kw["H"])
    mod = ref.sync_batchnorm.SynchronizedBatchNorm2d(kw["C"])
    with torch.no_grad():
        mod.weight.copy_(c["weight"])
        mod.bias.copy_(c["bias"])
    shards = c["x"].chunk(kw["shards"], dim=0)
    C = kw["C"]
    sums = [s.reshape(s.shape[0], C, -1).sum(0).sum(-1) for s in shards]  # batchnorm.py:60-62
    ssums = [(s.reshape(s.shape[0], C, -1) ** 2).sum(0).sum(-1) for s in shards]
    size = sum(s.shape[0] * s.shape[2] * s.shape[3] for s in shards)
    mean, inv_std = mod._compute_mean_std(sum(sums), sum(ssums), size)  # :113-125
    outs = [((s.reshape(s.shape[0], C, -1) - mean.view(1, C, 1)) * (inv_std * mod.weight).view(1, C, 1)
             + mod.bias.view(1, C, 1)).reshape(s.shape) for s in shards]  # :71-75
    return dict(kind="bn", recipe=json.dumps(kw), x=_np(c["x"]), weight=_np(c["weight"]), bias=_np(c["bias"]),
                mean=_np(mean), inv_std=_np(inv_std), out=_np(torch.cat(outs, 0)),
                running_mean=_np(mod.running_mean), running_var=_np(mod.running_var))


def main():
    warnings.simplefilter("ignore")
    torch.set_num_threads(1)  # fixed reduction order
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    tot = 0
    tables = ((WORDS_CASES, gen_words), (SENT_CASES, gen_sent), (GAG_CASES, gen_gag), (BN_CASES, gen_bn), (ATTR_CASES, gen_attr),
              (EMB_CASES, gen_emb))
    only = os.environ.get("EEGAN_GOLDEN_ONLY")  # e.g. "attr_,emb_": regenerate only these fixture families
    for table, fn in tables:
        for name, kw in table.items():
            if only and not any(name.startswith(p) for p in only.split(",")):
                continue
            d = fn(ref, name, kw)
            path = os.path.join(OUT, name + ".npz")
            np.savez_compressed(path, **d)
            tot += os.path.getsize(path)
            print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))
    print("total %.1f KB" % (tot / 1024))


if __name__ == "__main__":
    main()

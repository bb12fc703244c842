"""GPU (B200): the CUDA path, called through the C-ABI via the drop-in Python boundary,
against (a) the committed fixtures produced by the live reference and (b) the CPU oracle on
the same seeded inputs.  Stated tolerances: sim |d| <= 1e-4 abs, losses
<= 1e-5 rel, grads <= 1e-4 relative-to-max, attention maps <= 1e-6 abs, argmax bit-exact."""
import json

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from helpers import finite_close, gag_inputs, relmax, words_inputs
from oracle import cases
from oracle import damsm_oracle as O
from oracle.make_golden import IMG_GRAD_STRIDE, W0, W1

pytestmark = pytest.mark.gpu

TOL_SIM, TOL_LOSS, TOL_GRAD, TOL_ATT = 1e-4, 1e-5, 1e-4, 1e-6
# The "stress" inputs (randn x randn, score std ~16, near-one-hot softmaxes) amplify the 2^-21
# relative error of the 3xTF32 contraction: attention maps are held to 2.5e-6 there (the fp32
# reference itself sits ~1e-6 from float64 on that set).
TOL_ATT_STRESS = 2.5e-6


def dev(t):
    return None if t is None else t.cuda()


@pytest.mark.parametrize("name", golden_names("words_"))
def test_words_loss_vs_reference_fixture(cuda_lib, name):
    import eegan_b200 as E
    g = load_golden(name)
    kw, c = words_inputs(g)
    B = kw["B"]
    img = c["img"].cuda().requires_grad_()
    words = c["words"].cuda().requires_grad_()
    # class_ids stays a CPU tensor as in train.py:423; cap_lens/labels are device tensors
    l0, l1, att = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
    (W0 * l0 + W1 * l1).backward()
    assert abs(l0.item() - float(g["loss0"])) <= TOL_LOSS * max(1.0, abs(float(g["loss0"])))
    assert abs(l1.item() - float(g["loss1"])) <= TOL_LOSS * max(1.0, abs(float(g["loss1"])))
    sim, att2 = E.words_similarity(img.detach(), words.detach(), c["cap_lens"].cuda(), c["class_ids"], B)
    finite_close(sim.cpu(), g["sim"], TOL_SIM)
    assert len(att) == B and not isinstance(att, torch.Tensor)  # device cap_lens: a lazily sliced Sequence of B maps
    got = np.concatenate([a.detach().cpu().numpy().reshape(-1) for a in att])
    np.testing.assert_allclose(got, g["att"], atol=TOL_ATT_STRESS if kw["kind"] == "stress" else TOL_ATT)
    off = 0
    for i, T in enumerate(g["cap_lens"]):  # argmax word per region: bit-exact
        R = att[i].shape[2] * att[i].shape[3]
        assert tuple(att[i].shape[:2]) == (1, int(T))
        assert np.array_equal(g["att"][off:off + T * R].reshape(T, R).argmax(0), got[off:off + T * R].reshape(T, R).argmax(0))
        off += T * R
    assert relmax(words.grad.cpu(), g["d_words"]) <= TOL_GRAD
    d_img = img.grad.reshape(B, img.shape[1], -1).cpu()
    assert float((d_img[:, ::IMG_GRAD_STRIDE[0], ::IMG_GRAD_STRIDE[1]] - torch.from_numpy(g["d_img_sub"])).abs().max()) \
        <= TOL_GRAD * float(g["d_img_absmax"])
    np.testing.assert_allclose(cases.checksum(d_img), g["d_img_checksum"], rtol=2e-3)
    # padded words receive exactly zero gradient
    for i, T in enumerate(g["cap_lens"]):
        assert float(words.grad[i, :, int(T):].abs().max() if T < words.shape[2] else 0.0) == 0.0


@pytest.mark.parametrize("kind,cls,B,T", [("realistic", "cub", 48, 18), ("stress", "unique", 16, 18),
                                         ("realistic", "none", 7, 20), ("stress", "cub", 33, 5),
                                         ("realistic", "unique", 64, 20)])  # the last: BASELINE config 3 (COCO shape)
def test_words_loss_vs_oracle(cuda_lib, kind, cls, B, T):
    import eegan_b200 as E
    c = cases.words_case(B, T, kind=kind, class_mode=cls, seed=77, min_len=min(5, T))
    img_o = c["img"].double().requires_grad_()
    words_o = c["words"].double().requires_grad_()
    o0, o1, oatt, osim = O.dense_words_loss(img_o, words_o, c["labels"], c["cap_lens"], c["class_ids"])
    (o0 + o1).backward()
    img = c["img"].cuda().requires_grad_()
    words = c["words"].cuda().requires_grad_()
    l0, l1, att = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
    (l0 + l1).backward()
    assert abs(l0.item() - o0.item()) <= 5 * TOL_LOSS * max(1.0, abs(o0.item()))
    assert abs(l1.item() - o1.item()) <= 5 * TOL_LOSS * max(1.0, abs(o1.item()))
    assert relmax(img.grad.cpu(), img_o.grad) <= TOL_GRAD
    assert relmax(words.grad.cpu(), words_o.grad) <= TOL_GRAD
    flips = 0
    for a, b in zip(att, oatt):
        assert float((a.cpu().double() - b.detach()).abs().max()) <= (TOL_ATT_STRESS if kind == "stress" else TOL_ATT)
        flips += int((a.cpu().reshape(a.shape[1], -1).argmax(0) != b.detach().reshape(b.shape[1], -1).argmax(0)).sum())
    if kind == "realistic":
        assert flips == 0


def test_words_loss_only_img_grad_and_labels_none(cuda_lib):
    """train.py:172 detaches the words; labels=None returns (None, None, att_maps) (:336-340)."""
    import eegan_b200 as E
    c = cases.words_case(6, 12, seed=3)
    img = c["img"].cuda().requires_grad_()
    l0, l1, _ = E.words_loss(img, c["words"].cuda(), c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], 6)
    (l0 + l1).backward()
    img2 = c["img"].cuda().requires_grad_()
    w2 = c["words"].cuda().requires_grad_()
    m0, m1, _ = E.words_loss(img2, w2, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], 6)
    (m0 + m1).backward()
    assert torch.allclose(img.grad, img2.grad, rtol=0, atol=1e-7)
    n0, n1, att = E.words_loss(c["img"].cuda(), c["words"].cuda(), None, c["cap_lens"].cuda(), c["class_ids"], 6)
    assert n0 is None and n1 is None and len(att) == 6


@pytest.mark.parametrize("name", golden_names("sent_"))
def test_sent_loss_vs_reference_fixture(cuda_lib, name):
    import eegan_b200 as E
    g = load_golden(name)
    B = json.loads(str(g["recipe"]))["B"]
    cls = None if g["class_ids"][0] < 0 else torch.from_numpy(g["class_ids"])
    cnn = torch.from_numpy(g["cnn"]).cuda().requires_grad_()
    rnn = torch.from_numpy(g["rnn"]).cuda().requires_grad_()
    l0, l1 = E.sent_loss(cnn, rnn, torch.arange(B).cuda(), cls, B)
    (W0 * l0 + W1 * l1).backward()
    assert abs(l0.item() - float(g["loss0"])) <= TOL_LOSS * max(1.0, abs(float(g["loss0"])))
    assert abs(l1.item() - float(g["loss1"])) <= TOL_LOSS * max(1.0, abs(float(g["loss1"])))
    finite_close(E.sent_similarity(cnn.detach(), rnn.detach(), cls, B).cpu(), g["scores"], TOL_SIM)
    assert relmax(cnn.grad.cpu(), g["d_cnn"]) <= TOL_GRAD and relmax(rnn.grad.cpu(), g["d_rnn"]) <= TOL_GRAD


@pytest.fixture(params=[0, 1], ids=["gag-cuda-cores", "gag-tcgen05-fwd"])
def gag_engine(request, cuda_lib):
    """Both forward engines of GlobalAttentionGeneral: the CUDA-core kernels (default) and the tcgen05 kernel."""
    assert cuda_lib.eegan_set_gag_engine(request.param) == 0
    yield request.param
    cuda_lib.eegan_set_gag_engine(1)  # the default


@pytest.mark.parametrize("name", golden_names("gag_"))
def test_gag_vs_reference_fixture(cuda_lib, gag_engine, name):
    import eegan_b200 as E
    g = load_golden(name)
    kw, c = gag_inputs(g)
    x, k, v = (c[n].cuda().requires_grad_() for n in ("x", "key", "value"))
    mod = E.GlobalAttentionGeneral(kw["idf"], 256)
    if c["mask"] is not None:
        mod.applyMask(c["mask"].cuda())
    out, attn = mod(x, k, v)
    assert tuple(out.shape) == g["out"].shape and tuple(attn.shape) == g["attn"].shape
    assert relmax(out.detach().cpu(), g["out"]) <= 1e-5
    np.testing.assert_allclose(attn.detach().cpu().numpy(), g["attn"], atol=TOL_ATT)
    assert np.array_equal(attn.detach().cpu().reshape(attn.shape[0], attn.shape[1], -1).argmax(1).numpy().astype(np.int32),
                          g["attn_argmax"])
    gen = cases._gen(99)
    go = torch.randn(out.shape, generator=gen).cuda()
    ga = torch.randn(attn.shape, generator=gen).cuda()
    ((out * go).sum() + (attn * ga).sum()).backward()
    assert relmax(x.grad.cpu(), g["d_x"]) <= TOL_GRAD
    assert relmax(k.grad.cpu(), g["d_key"]) <= TOL_GRAD
    assert relmax(v.grad.cpu(), g["d_value"]) <= TOL_GRAD


def test_gag_intended_mask_mode_and_big_shape(cuda_lib, gag_engine):
    import eegan_b200 as E
    c = cases.gag_case(4, 64, 64, 18, seed=9)  # Q = 4096 pixels
    mod = E.GlobalAttentionGeneral(64, 256, mask_mode="intended")
    mod.applyMask(c["mask"].cuda())
    out, attn = mod(c["x"].cuda(), c["key"].cuda(), c["value"].cuda())
    oo, oa = O.port_global_attention(c["x"], c["key"], c["value"], c["mask"], "intended")
    assert relmax(out.cpu(), oo) <= 1e-5 and float((attn.cpu() - oa).abs().max()) <= TOL_ATT


@pytest.mark.parametrize("B,idf,H,T", [(3, 256, 16, 20), (2, 48, 12, 7), (5, 32, 23, 18)])
def test_gag_tensor_core_forward_shapes(cuda_lib, B, idf, H, T):
    """The tcgen05 forward at its shape edges: idf = 256 (two TMEM x stages, 256 accumulator columns), idf = 48
    (K tail zero-filled by TMA, 16-column tail of the accumulator), Q = 529 (not a multiple of the 128-pixel
    tile; Q % 4 != 0 falls back to the CUDA-core kernel, which must agree too), quirk mask mode."""
    import eegan_b200 as E
    c = cases.gag_case(B, idf, H, T, seed=B + idf)
    res = {}
    try:
        for eng in (0, 1):
            cuda_lib.eegan_set_gag_engine(eng)
            mod = E.GlobalAttentionGeneral(idf, 256)
            mod.applyMask(c["mask"].cuda())
            res[eng] = mod(c["x"].cuda(), c["key"].cuda(), c["value"].cuda())
    finally:
        cuda_lib.eegan_set_gag_engine(1)
    # float64 oracle: with K = idf up to 256 an fp32 reference carries ~1e-6 of its own rounding in a probability near 1
    oo, oa = O.port_global_attention(c["x"].double(), c["key"].double(), c["value"].double(), c["mask"], "reference")
    for eng in (0, 1):
        out, attn = res[eng]
        ok = torch.isfinite(oa)  # fully masked rows are NaN in the reference (softmax of all -inf) and here
        assert torch.equal(torch.isnan(attn.cpu()), torch.isnan(oa))
        assert float((attn.cpu().double()[ok] - oa[ok]).abs().max()) <= 2.5e-6, eng
        okr = torch.isfinite(oo)
        assert float((out.cpu().double()[okr] - oo[okr]).abs().max()) <= 1e-5 * float(oo[okr].abs().max()), eng


@pytest.fixture(params=[1, 0], ids=["gag-bwd-tcgen05", "gag-bwd-cuda-cores"])
def gag_bwd_engine(request, cuda_lib):
    """Both backward engines of GlobalAttentionGeneral: the one-pass tcgen05 kernel (gag_tc_bwd.cu; default where the shape
    allows: idf = 32 / 64 / 128 with a gradient on `out`) and the CUDA-core kernels (gag_bwd2.cu / gag.cu)."""
    assert cuda_lib.eegan_set_gag_bwd_engine(request.param) == 0
    yield request.param
    cuda_lib.eegan_set_gag_bwd_engine(1)


@pytest.mark.parametrize("B,idf,H,T,grads", [(3, 128, 16, 18, "both"), (2, 64, 20, 18, "both"), (4, 32, 24, 5, "both"),
                                              (2, 96, 12, 20, "both"), (2, 256, 8, 32, "both"), (3, 64, 16, 18, "out"),
                                              (3, 64, 16, 18, "attn"), (2, 48, 12, 7, "both"), (2, 32, 23, 18, "both"),
                                              (2, 64, 12, 32, "both"), (5, 128, 10, 21, "out")])
def test_gag_backward_shapes(cuda_lib, gag_bwd_engine, B, idf, H, T, grads):
    """The two-kernel backward (gag_bwd2.cu: per-pixel-quad pass + key/value row sums with the lanes along the pixels) at its
    shape edges — idf = 32 / 64 / 96 / 128 / 256 (1 … 8 channel sets of 8 channels per warp; 4 per warp once T > 20),
    T = 5 … 32 (every TP instantiation), Q = 144 … 576 (partial 128-pixel rounds, several pixel chunks), a gradient on only
    one of the two outputs — and the one-kernel fallback it hands idf = 48 and Q % 4 != 0 to, against the float64
    oracle's autograd.  With the tcgen05 engine selected the same cases run through gag_tc_bwd.cu where it takes the shape
    (idf = 32 / 64 / 128, a gradient on `out`; partial last tile at Q = 144 / 400 / 576, T <= 20 and T > 20 instantiations, one
    and two 64-channel units) and through the CUDA-core kernels elsewhere."""
    import eegan_b200 as E
    c = cases.gag_case(B, idf, H, T, seed=B * 7 + idf + T, masked=False)
    gen = cases._gen(5)
    go = torch.randn(B, idf, H, H, generator=gen)
    ga = torch.randn(B, T, H, H, generator=gen)
    xo, ko, vo = (c[n].double().requires_grad_() for n in ("x", "key", "value"))
    oo, oa = O.port_global_attention(xo, ko, vo, None)
    loss_o = (oo * go.double()).sum() * (grads != "attn") + (oa * ga.double()).sum() * (grads != "out")
    loss_o.backward()
    x, k, v = (c[n].cuda().requires_grad_() for n in ("x", "key", "value"))
    out, attn = E.GlobalAttentionGeneral(idf, 256)(x, k, v)
    loss = 0.0
    if grads != "attn":
        loss = loss + (out * go.cuda()).sum()
    if grads != "out":
        loss = loss + (attn * ga.cuda()).sum()
    loss.backward()
    assert relmax(x.grad.cpu(), xo.grad) <= TOL_GRAD
    assert relmax(k.grad.cpu(), ko.grad) <= TOL_GRAD
    if grads != "attn":
        assert relmax(v.grad.cpu(), vo.grad) <= TOL_GRAD
    else:  # value only feeds `out`
        assert float(v.grad.abs().max()) == 0.0


def _gag_ref64_gpu(x, key, val, mask, go, ga):
    """float64 autograd of miscc/DAMSM_losses.py:96-132 (intended mask mode) on the GPU: the checker for shapes where the CPU
    oracle would take minutes."""
    x, key, val = (t.detach().double().requires_grad_() for t in (x, key, val))
    B, idf, H, W = x.shape
    s = torch.bmm(x.view(B, idf, H * W).transpose(1, 2), key)
    if mask is not None:
        s = s.masked_fill(mask[:, None, :], float("-inf"))
    p = torch.softmax(s, dim=2)
    out = torch.bmm(val, p.transpose(1, 2)).view(B, idf, H, W)
    attn = p.transpose(1, 2).reshape(B, -1, H, W)
    ((out * go.double()).sum() + ((attn * ga.double()).sum() if ga is not None else 0.0)).backward()
    return out.detach(), attn.detach(), x.grad, key.grad, val.grad


@pytest.mark.parametrize("B,idf,H,T,dattn", [(40, 64, 128, 18, True), (48, 128, 64, 18, True), (37, 32, 100, 12, False),
                                              (150, 32, 72, 18, True)])
def test_gag_tensor_core_backward_long_chains(cuda_lib, B, idf, H, T, dattn):
    """gag_tc_bwd.cu where a CTA walks many tiles: several tiles per CTA (the unit ring wraps, every barrier changes phase many
    times), more than one accumulation group (43 tiles per CTA at B = 40, 128^2: two groups; B = 150: one CTA per sample, 41
    tiles), a benchmarked shape (B = 48, 64^2 x 128), a partial last tile (Q = 10000) and no gradient on attn — against float64
    autograd of the same math on the GPU.  The d_key / d_value bound also holds the accumulation-drift fix (one group per 32
    tiles) in place: without it the error grows with the number of tiles a CTA accumulates."""
    import eegan_b200 as E
    gen = torch.Generator(device="cpu").manual_seed(B + idf + H)
    lens = torch.randint(3, T + 1, (B,), generator=gen)
    mask = (torch.arange(T)[None, :] >= lens[:, None]).cuda()
    x = torch.randn(B, idf, H, H, generator=gen).cuda().requires_grad_()
    key = (torch.randn(B, idf, T, generator=gen) * idf ** -0.5).cuda().requires_grad_()
    val = torch.randn(B, idf, T, generator=gen).cuda().requires_grad_()
    go = torch.randn(B, idf, H, H, generator=gen).cuda()
    ga = torch.randn(B, T, H, H, generator=gen).cuda() if dattn else None
    mod = E.GlobalAttentionGeneral(idf, 256, mask_mode="intended")
    mod.applyMask(mask)
    assert cuda_lib.eegan_get_gag_bwd_engine() == 1
    out, attn = mod(x, key, val)
    torch.autograd.backward([out, attn] if dattn else [out], [go, ga] if dattn else [go])
    ro, ra, rx, rk, rv = _gag_ref64_gpu(x, key, val, mask, go, ga)
    assert float((attn.double() - ra).abs().max()) <= TOL_ATT_STRESS
    assert relmax(out.detach().cpu(), ro.cpu()) <= 1e-5
    assert relmax(x.grad.cpu(), rx.cpu()) <= TOL_GRAD
    assert relmax(key.grad.cpu(), rk.cpu()) <= TOL_GRAD
    assert relmax(val.grad.cpu(), rv.cpu()) <= TOL_GRAD
    # the two engines agree far inside the tolerance (same math, bf16-pair operands against fp32 FMA)
    gx, gk, gv = x.grad.clone(), key.grad.clone(), val.grad.clone()
    x.grad = key.grad = val.grad = None
    try:
        cuda_lib.eegan_set_gag_bwd_engine(0)
        out, attn = mod(x, key, val)
        torch.autograd.backward([out, attn] if dattn else [out], [go, ga] if dattn else [go])
    finally:
        cuda_lib.eegan_set_gag_bwd_engine(1)
    assert relmax(gx.cpu(), x.grad.cpu()) <= 5e-5 and relmax(gk.cpu(), key.grad.cpu()) <= 5e-5 and relmax(gv.cpu(), val.grad.cpu()) <= 5e-5


@pytest.mark.parametrize("name", golden_names("words_")[:3])
def test_func_attention_vs_reference_fixture(cuda_lib, name):
    import eegan_b200 as E
    g = load_golden(name)
    kw, c = words_inputs(g)
    B, T0 = kw["B"], int(c["cap_lens"][0])
    q = c["words"][0:1, :, :T0].repeat(B, 1, 1).cuda().requires_grad_()
    ctx = c["img"].cuda().requires_grad_()
    u, attn = E.func_attention(q, ctx, 5.0)
    assert relmax(u.detach().cpu(), g["fa_u"]) <= 1e-5
    a = attn.detach().cpu().reshape(B, T0, -1)
    np.testing.assert_allclose(a[:, :, ::5].numpy(), g["fa_attn_sub"], atol=TOL_ATT)
    assert np.array_equal(a.argmax(dim=1).numpy().astype(np.int32), g["fa_attn_argmax_words"])
    # backward on both outputs against the oracle's autograd
    qo = q.detach().cpu().double().requires_grad_()
    co = ctx.detach().cpu().double().requires_grad_()
    uo, ao = O.port_func_attention(qo, co, 5.0)
    gen = cases._gen(5)
    gu, ga = torch.randn(u.shape, generator=gen), torch.randn(attn.shape, generator=gen)
    ((uo * gu.double()).sum() + (ao * ga.double()).sum()).backward()
    ((u * gu.cuda()).sum() + (attn * ga.cuda()).sum()).backward()
    assert relmax(q.grad.cpu(), qo.grad) <= TOL_GRAD and relmax(ctx.grad.cpu(), co.grad) <= TOL_GRAD


def test_cosine_similarity(cuda_lib):
    import eegan_b200 as E
    gen = cases._gen(1)
    a = torch.randn(37, 256, generator=gen)
    b = torch.randn(37, 256, generator=gen)
    a[3] = 0.0  # clamp path: |a||b| < eps
    x, y = a.cuda().requires_grad_(), b.cuda().requires_grad_()
    out = E.cosine_similarity(x, y)
    ao, bo = a.double().requires_grad_(), b.double().requires_grad_()
    ref = O.port_cosine_similarity(ao, bo)
    assert float((out.cpu().double() - ref).abs().max()) <= 1e-6
    w = torch.randn(37, generator=gen)
    (out * w.cuda()).sum().backward()
    (ref * w.double()).sum().backward()
    assert relmax(x.grad.cpu(), ao.grad) <= TOL_GRAD and relmax(y.grad.cpu(), bo.grad) <= TOL_GRAD


def test_full_size_properties(cuda_lib):
    """BASELINE config sizes (B=48 CUB, B=64 COCO): size-independent properties.
    (1) permuting the batch permutes the grid; (2) padding words beyond cap_lens never matter;
    (3) every attention row sums to one; (4) loss gradients sum to zero over a softmax row of
    the grid: sum_i dL0/dsim[j,i] = 0."""
    import eegan_b200 as E
    for B, T in ((48, 18), (64, 20)):
        c = cases.words_case(B, T, seed=123, class_mode="none")
        img, words, lens = c["img"].cuda(), c["words"].cuda(), c["cap_lens"].cuda()
        sim, att = E.words_similarity(img, words, lens, None, B)
        perm = torch.randperm(B, generator=cases._gen(4)).cuda()
        sim_p, _ = E.words_similarity(img[perm], words[perm], lens[perm], None, B)
        assert float((sim_p - sim[perm][:, perm]).abs().max()) <= 2e-5
        junk = words.clone()
        for i in range(B):
            junk[i, :, int(c["cap_lens"][i]):] = 1e3
        sim_j, _ = E.words_similarity(img, junk, lens, None, B)
        assert torch.equal(sim_j, sim)
        for a in att:
            assert float((a.sum(dim=(2, 3)) - 1).abs().max()) <= 1e-5


@pytest.mark.parametrize("B,T", [(256, 20), (512, 20)])
def test_large_batch_blocks_vs_oracle(cuda_lib, B, T):
    """BASELINE config 5 sizes (flower sweep up to B = 512; the 31.5 GB stash of B = 512 is allocated here).  The float64
    oracle of the whole grid would need tens of GB on the host, so VALUES are checked block-wise, which loses nothing: the
    similarity m[j, i] depends on image j and caption i only, and with an upstream gradient G that is non-zero on a block
    (J x I) only, d_words[i in I] and d_img[j in J] are exactly the oracle's gradients of that block."""
    import eegan_b200 as E
    from eegan_b200 import damsm_losses as dl
    c = cases.words_case(B, T, seed=B, class_mode="cub")
    gen = cases._gen(B + 1)
    J = torch.randperm(B, generator=gen)[:6].sort().values
    I = torch.randperm(B, generator=gen)[:5].sort().values
    J[0], I[0] = 0, 0                 # first tile / first bin
    J[-1], I[-1] = B - 1, B - 1       # last tile / last (partial) bin
    G = torch.zeros(B, B)
    Gb = torch.randn(len(J), len(I), generator=gen)
    G[J[:, None], I[None, :]] = Gb
    img, words = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
    m, att = dl.pair_grid(img, words, c["cap_lens"].cuda())
    (m * G.cuda()).sum().backward()
    io, wo = c["img"][J].double().requires_grad_(), c["words"][I].double().requires_grad_()
    t = O.dense_pair_terms(io, wo, c["cap_lens"][I])
    (t["m"] * Gb.double()).sum().backward()
    assert float((m.detach().cpu().double()[J[:, None], I[None, :]] - t["m"].detach()).abs().max()) <= TOL_SIM / 10.0  # m = sim / gamma3
    assert relmax(words.grad[I.cuda()].cpu(), wo.grad) <= TOL_GRAD
    assert relmax(img.grad[J.cuda()].cpu(), io.grad) <= TOL_GRAD
    # rows / columns outside the block receive exactly nothing
    rest_i = torch.ones(B, dtype=torch.bool); rest_i[I] = False
    rest_j = torch.ones(B, dtype=torch.bool); rest_j[J] = False
    assert float(words.grad[rest_i.cuda()].abs().max()) == 0.0 and float(img.grad[rest_j.cuda()].abs().max()) == 0.0
    # the losses of the full grid: finite, and the two-way CE of the reference on the GPU's own sim (float64 on the host)
    l0, l1, _ = E.words_loss(c["img"].cuda(), c["words"].cuda(), c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
    sim, _ = E.words_similarity(c["img"].cuda(), c["words"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
    r0, r1 = O._two_way_ce(sim.cpu().double(), c["labels"])
    assert abs(l0.item() - float(r0)) <= 5 * TOL_LOSS * max(1.0, abs(float(r0))) and abs(l1.item() - float(r1)) <= 5 * TOL_LOSS * max(1.0, abs(float(r1)))
    # diagonal attention maps of the sampled captions that are also sampled images
    for i in [int(v) for v in I if int(v) in set(int(u) for u in J)]:
        ji, ii = int((J == i).nonzero()[0]), int((I == i).nonzero()[0])
        Ti = int(c["cap_lens"][i])
        assert float((att[i, :Ti].cpu().double() - t["a"][ji, ii, :Ti].detach()).abs().max()) <= TOL_ATT
    from eegan_b200 import fastpath
    fastpath.clear_plans()
    torch.cuda.empty_cache()


def test_ffma_engine_still_matches(cuda_lib):
    """The exact-fp32 CUDA-core engine (0), the unfused tensor-core pipeline (1), the fused 3xTF32 engine (2) and the
    fused half-pair engine (3) stay selectable for A/B validation: all agree to fp32 noise."""
    import eegan_b200 as E
    c = cases.words_case(12, 18, seed=5)
    res = {}
    default = cuda_lib.eegan_get_contraction_engine()
    try:
        for eng in (0, 1, 2, 3):
            assert cuda_lib.eegan_set_contraction_engine(eng) == 0
            img = c["img"].cuda().requires_grad_()
            words = c["words"].cuda().requires_grad_()
            l0, l1, _ = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], 12)
            (l0 + l1).backward()
            res[eng] = (l0.item(), l1.item(), img.grad.clone(), words.grad.clone())
    finally:
        cuda_lib.eegan_set_contraction_engine(default)
    for eng in (1, 2, 3):
        assert abs(res[0][0] - res[eng][0]) <= 2e-5 and abs(res[0][1] - res[eng][1]) <= 2e-5, eng
        assert relmax(res[eng][2], res[0][2]) <= TOL_GRAD and relmax(res[eng][3], res[0][3]) <= TOL_GRAD, eng


def test_graphed_words_loss_matches_eager(cuda_lib):
    """The CUDA-graph step API replays exactly the kernels of the eager call."""
    import eegan_b200 as E
    from eegan_b200.graphed import GraphedWordsLoss
    B, T = 10, 18
    gw = GraphedWordsLoss(B, 256, 17, 17, T, "cuda", use_class_ids=True, words_grad=True, w0=1.0, w1=2.0)
    for seed in (31, 32):  # second call exercises replay with new data and new ragged lengths
        c = cases.words_case(B, T, seed=seed)
        l0, l1, d_img, d_words = gw(c["img"].cuda(), c["words"].cuda(), c["cap_lens"].cuda(), c["class_ids"].cuda())
        img = c["img"].cuda().requires_grad_()
        words = c["words"].cuda().requires_grad_()
        e0, e1, _ = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
        (e0 + 2.0 * e1).backward()
        assert abs(l0.item() - e0.item()) <= 1e-6 and abs(l1.item() - e1.item()) <= 1e-6
        assert relmax(d_img, img.grad) <= 1e-6 and relmax(d_words, words.grad) <= 1e-6

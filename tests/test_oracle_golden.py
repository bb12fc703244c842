"""CPU: both oracle restatements against the fixtures generated from the live reference
(oracle/make_golden.py).  Tolerances: the fp32 port runs the reference's op sequence, so it
is held to 2e-6; the float64 dense oracle differs from the fp32 reference by fp32 rounding."""
import json

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from helpers import finite_close, gag_inputs, relmax, words_inputs
from oracle import cases
from oracle import damsm_oracle as O
from oracle.make_golden import IMG_GRAD_STRIDE, W0, W1

torch.set_num_threads(2)


@pytest.mark.parametrize("name", golden_names("words_"))
def test_words_port_matches_reference_fixture(name):
    g = load_golden(name)
    kw, c = words_inputs(g)
    B = kw["B"]
    img = c["img"].clone().requires_grad_()
    words = c["words"].clone().requires_grad_()
    l0, l1, att = O.port_words_loss(img, words, c["labels"], c["cap_lens"], c["class_ids"], B)
    (W0 * l0 + W1 * l1).backward()
    assert abs(float(l0) - float(g["loss0"])) < 2e-6 and abs(float(l1) - float(g["loss1"])) < 2e-6
    sim, _ = O.port_words_similarity(c["img"], c["words"], c["cap_lens"], c["class_ids"], B)
    finite_close(sim, g["sim"], 2e-5)
    np.testing.assert_allclose(np.concatenate([a.detach().numpy().reshape(-1) for a in att]), g["att"], atol=1e-7)
    assert relmax(words.grad, g["d_words"]) < 1e-5
    d_img = img.grad.reshape(B, img.shape[1], -1)
    assert relmax(d_img[:, ::IMG_GRAD_STRIDE[0], ::IMG_GRAD_STRIDE[1]], g["d_img_sub"]) < 1e-5
    np.testing.assert_allclose(cases.checksum(d_img), g["d_img_checksum"], rtol=1e-4)
    if "img" in g:  # tiny case: the stored inputs are authoritative
        np.testing.assert_array_equal(g["img"], c["img"].numpy())


@pytest.mark.parametrize("name", golden_names("words_"))
def test_words_dense64_and_manual_backward(name):
    g = load_golden(name)
    kw, c = words_inputs(g)
    img = c["img"].double().requires_grad_()
    words = c["words"].double().requires_grad_()
    l0, l1, att, sim = O.dense_words_loss(img, words, c["labels"], c["cap_lens"], c["class_ids"])
    assert abs(float(l0) - float(g["loss0"])) < 2e-5 and abs(float(l1) - float(g["loss1"])) < 2e-5
    finite_close(sim.detach(), torch.from_numpy(g["sim"]).double(), 1e-4)
    (W0 * l0 + W1 * l1).backward()
    assert relmax(words.grad, g["d_words"]) < 5e-5
    # the hand-derived backward (what the CUDA kernels implement) against autograd, float64
    dm = O.ce_pair_grad(sim.detach(), c["labels"], W0, W1) * O.GAMMA3
    mi, mw = O.dense_words_backward(img.detach(), words.detach(), c["cap_lens"], dm)
    assert relmax(mi, img.grad) < 1e-11 and relmax(mw, words.grad) < 1e-11
    # argmax word per region of the diagonal attention maps is bit-exact vs the fixture
    got = np.concatenate([a.detach().numpy().reshape(-1) for a in att])
    lens = g["cap_lens"]
    off = 0
    for i, T in enumerate(lens):
        R = att[i].shape[2] * att[i].shape[3]
        a_ref = g["att"][off:off + T * R].reshape(T, R)
        a_got = got[off:off + T * R].reshape(T, R)
        assert np.array_equal(a_ref.argmax(0), a_got.argmax(0))
        off += T * R


@pytest.mark.parametrize("name", golden_names("words_"))
def test_func_attention_fixture(name):
    g = load_golden(name)
    kw, c = words_inputs(g)
    B, T0 = kw["B"], int(c["cap_lens"][0])
    q = c["words"][0:1, :, :T0].repeat(B, 1, 1)
    u, attn = O.port_func_attention(q, c["img"], O.GAMMA1)
    assert relmax(u, g["fa_u"]) < 1e-6
    a = attn.reshape(B, T0, -1)
    np.testing.assert_allclose(a[:, :, ::5].numpy(), g["fa_attn_sub"], atol=1e-7)
    assert np.array_equal(a.argmax(dim=1).numpy().astype(np.int32), g["fa_attn_argmax_words"])


@pytest.mark.parametrize("name", golden_names("sent_"))
def test_sent_fixture(name):
    g = load_golden(name)
    kw = json.loads(str(g["recipe"]))
    B = kw["B"]
    cls = None if g["class_ids"][0] < 0 else torch.from_numpy(g["class_ids"])
    cnn = torch.from_numpy(g["cnn"]).requires_grad_()
    rnn = torch.from_numpy(g["rnn"]).requires_grad_()
    l0, l1 = O.port_sent_loss(cnn, rnn, torch.arange(B), cls, B)
    (W0 * l0 + W1 * l1).backward()
    assert abs(float(l0) - float(g["loss0"])) < 1e-6 and abs(float(l1) - float(g["loss1"])) < 1e-6
    finite_close(O.port_sent_similarity(cnn.detach(), rnn.detach(), cls, B), g["scores"], 1e-5)
    assert relmax(cnn.grad, g["d_cnn"]) < 1e-5 and relmax(rnn.grad, g["d_rnn"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("gag_"))
def test_gag_fixture(name):
    g = load_golden(name)
    kw, c = gag_inputs(g)
    out, attn = O.port_global_attention(c["x"], c["key"], c["value"], c["mask"], "reference")
    assert relmax(out, g["out"]) < 1e-6
    np.testing.assert_allclose(attn.numpy(), g["attn"], atol=1e-7)
    assert np.array_equal(attn.reshape(attn.shape[0], attn.shape[1], -1).argmax(1).numpy().astype(np.int32), g["attn_argmax"])
    if c["mask"] is not None and kw["B"] > 1:
        out_i, _ = O.port_global_attention(c["x"], c["key"], c["value"], c["mask"], "intended")
        assert relmax(out_i, g["out"]) > 1e-3, "the reference's mask pairing is NOT mask[b] (SURVEY D8)"


@pytest.mark.parametrize("name", golden_names("bn_"))
def test_syncbn_fixture(name):
    g = load_golden(name)
    kw = json.loads(str(g["recipe"]))
    x = torch.from_numpy(g["x"])
    outs, mean, inv_std, unbiased = O.syncbn_forward(list(x.chunk(kw["shards"], 0)), torch.from_numpy(g["weight"]),
                                                     torch.from_numpy(g["bias"]))
    np.testing.assert_allclose(mean.numpy(), g["mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(inv_std.numpy(), g["inv_std"], rtol=1e-5)
    np.testing.assert_allclose(torch.cat(outs, 0).numpy(), g["out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(0.1 * mean.numpy(), g["running_mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(0.9 + 0.1 * unbiased.numpy(), g["running_var"], rtol=1e-5)


def test_port_rprecision_known_answers():
    """The R-precision restatement (test.py:323-330) on hand-checkable inputs."""
    cnn = torch.tensor([[1.0, 0.0], [0.0, 2.0], [0.0, 0.0]])
    rnn = torch.tensor([[[3.0, 0.0], [0.0, 1.0], [-1.0, 0.0]],       # candidate 0 is parallel: hit
                        [[1.0, 0.0], [0.0, 5.0], [0.0, -1.0]],       # candidate 1 wins: miss
                        [[1.0, 1.0], [2.0, 0.0], [0.0, 3.0]]])       # zero image code: all scores 0 -> argmax 0
    hits, best, scores = O.port_rprecision(cnn, rnn)
    assert hits.tolist() == [True, False, True] and best.tolist() == [0, 1, 0]
    assert torch.allclose(scores[0], torch.tensor([1.0, 0.0, -1.0])) and torch.allclose(scores[1], torch.tensor([0.0, 1.0, -1.0]))
    assert float(scores[2].abs().max()) == 0.0

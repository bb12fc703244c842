"""CPU, build container only: the oracle against the LIVE, UNMODIFIED reference imported from
/root/reference (skipped on the GPU box, where the mount does not exist — the committed
fixtures carry the pin there)."""
import pytest
import torch

from oracle import cases
from oracle import damsm_oracle as O
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("kind,cls", [("realistic", "cub"), ("stress", "unique"), ("realistic", "none")])
def test_words_loss_bit_exact(kind, cls):
    ref = load_reference()
    c = cases.words_case(5, 9, kind=kind, class_mode=cls, seed=11)
    a = [c["img"].clone().requires_grad_(), c["words"].clone().requires_grad_()]
    b = [c["img"].clone().requires_grad_(), c["words"].clone().requires_grad_()]
    r0, r1, ratt = ref.losses.words_loss(a[0], a[1], c["labels"], c["cap_lens"], c["class_ids"], 5)
    p0, p1, patt = O.port_words_loss(b[0], b[1], c["labels"], c["cap_lens"], c["class_ids"], 5)
    (r0 + r1).backward()
    (p0 + p1).backward()
    assert torch.equal(r0, p0) and torch.equal(r1, p1)
    assert all(torch.equal(x, y) for x, y in zip(ratt, patt))
    assert torch.equal(a[0].grad, b[0].grad) and torch.equal(a[1].grad, b[1].grad)


def test_sent_and_gag_bit_exact():
    ref = load_reference()
    s = cases.sent_case(7, seed=5)
    r = ref.losses.sent_loss(s["cnn"], s["rnn"], s["labels"], s["class_ids"], 7)
    p = O.port_sent_loss(s["cnn"], s["rnn"], s["labels"], s["class_ids"], 7)
    assert torch.equal(r[0], p[0]) and torch.equal(r[1], p[1])
    g = cases.gag_case(3, 16, 6, 9, seed=5)
    mod = ref.losses.GlobalAttentionGeneral(16, 256)
    mod.applyMask(g["mask"])
    ro, ra = mod(g["x"], g["key"], g["value"])
    po, pa = O.port_global_attention(g["x"], g["key"], g["value"], g["mask"])
    assert torch.equal(ro, po) and torch.equal(ra, pa)


def test_affine_ssa_port_and_state_dict_match_reference():
    """models.py:43-86 run live (single replica, training mode) against the restatement, and the drop-in
    module's parameter / buffer names against the reference's state_dict."""
    import importlib
    load_reference()
    with __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore")
        models = importlib.import_module("models")
    torch.manual_seed(3)
    ref = models.affine_ssa(8, ntf=16)
    for p in ref.parameters():
        torch.nn.init.normal_(p, std=0.3)
    ref.train()
    feat = torch.randn(4, 8, 5, 5) * 1.5 + 0.3
    cond = torch.randn(4, 16)
    mask = torch.sigmoid(torch.randn(4, 1, 5, 5))
    out = ref(feat, cond, mask)
    port = O.port_affine_ssa(feat, ref.fc_gamma(cond), ref.fc_beta(cond), mask, eps=ref.norm2d.eps)
    assert float((out - port).abs().max()) <= 2e-6
    import eegan_b200.ssa as ssa
    ours = ssa.affine_ssa(8, ntf=16)
    assert set(ours.state_dict().keys()) == set(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict())

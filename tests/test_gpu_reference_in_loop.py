"""GPU (B200): the UNMODIFIED reference in the loop.

`oracle/stage_reference.py` puts a byte-for-byte copy of the reference's files for this path under the
git-ignored `baseline/_ref/` (it travels with the gpurun snapshot); `oracle/ref_loader.py` imports it.
Here the reference's own code runs on the same B200 (`cfg.CUDA = True`) next to the CUDA path:

  * `miscc.DAMSM_losses.words_loss` / `sent_loss` / `GlobalAttentionGeneral` at benchmark sizes, value for value;
  * `train.Trainer.DAMSM_loss` (train.py:419-435) imported once as the reference has it and once with
    `eegan_b200.install()` active (the zero-edit route of INTEGRATION.md);
  * `models.Gen` built with the reference's SyncBN and with eegan_b200's (24 layers), same weights, same inputs.
Skipped (not failed) where the staged copy is absent.
"""
import pytest
import torch

from helpers import relmax
from oracle import cases
from oracle import ref_loader as RL

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not RL.reference_available(), reason="reference neither mounted nor staged under baseline/_ref")
needs_models = pytest.mark.skipif(not RL.reference_models_available(), reason="reference models.py / train.py not staged")


class _cuda_cfg:
    """cfg.CUDA = True for the duration (the reference moves its masks to the GPU on that flag, DAMSM_losses.py:14)."""

    def __init__(self, cfg):
        self.cfg = cfg

    def __enter__(self):
        self.prev = self.cfg.CUDA
        self.cfg.CUDA = True

    def __exit__(self, *a):
        self.cfg.CUDA = self.prev


@needs_ref
@pytest.mark.parametrize("B,T,cls", [(48, 18, "cub"), (64, 20, "unique"), (16, 18, "none")])
def test_reference_words_loss_same_gpu(cuda_lib, B, T, cls):
    import warnings
    import eegan_b200 as E
    ref = RL.load_reference()
    c = cases.words_case(B, T, class_mode=cls, seed=11)
    ri, rw = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
    with _cuda_cfg(ref.cfg), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r0, r1, ratt = ref.losses.words_loss(ri, rw, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
        (r0 + 0.5 * r1).backward()
    for rep in range(3):  # 1st call plain launches, 2nd captures the plan's graphs, 3rd replays them
        img, words = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
        l0, l1, att = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
        (l0 + 0.5 * l1).backward()
        assert abs(l0.item() - r0.item()) <= 2e-5 * max(1.0, abs(r0.item())), rep
        assert abs(l1.item() - r1.item()) <= 2e-5 * max(1.0, abs(r1.item())), rep
        assert relmax(img.grad, ri.grad) <= 1e-4 and relmax(words.grad, rw.grad) <= 1e-4, rep
        assert len(att) == len(ratt) == B
        for a, b in zip(att, ratt):
            assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-6
            assert torch.equal(a.reshape(a.shape[1], -1).argmax(0), b.reshape(b.shape[1], -1).argmax(0))


@needs_ref
def test_reference_gag_same_gpu_at_a_benchmarked_shape(cuda_lib):
    """B=48, 64x64, idf=128 (bench `extra.global_attention_general[0]`) against the reference module itself."""
    import warnings
    import eegan_b200 as E
    ref = RL.load_reference()
    g = cases.gag_case(48, 128, 64, 18, seed=5)
    mask = g["mask"].cuda()
    x, k, v = (g[n].cuda().requires_grad_() for n in ("x", "key", "value"))
    rx, rk, rv = (g[n].cuda().requires_grad_() for n in ("x", "key", "value"))
    go = torch.randn(48, 128, 64, 64, generator=cases._gen(9)).cuda()
    ga = torch.randn(48, 18, 64, 64, generator=cases._gen(10)).cuda()
    rm = ref.losses.GlobalAttentionGeneral(128, 256)
    rm.applyMask(mask)
    with _cuda_cfg(ref.cfg), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ro, ra = rm(rx, rk, rv)
    m = E.GlobalAttentionGeneral(128, 256)
    m.applyMask(mask)
    o, a = m(x, k, v)
    ok = torch.isfinite(ra)  # rows whose mask row (quirk D8) covers every word are NaN in both
    assert torch.equal(torch.isfinite(a), ok)
    assert float((a[ok] - ra[ok]).abs().max()) <= 2e-6
    oko = torch.isfinite(ro)
    assert relmax(o[oko], ro[oko]) <= 2e-5
    # gradients on an unmasked copy (a fully masked row poisons the whole backward with NaN in the reference too)
    rm.applyMask(torch.zeros_like(mask))
    m.applyMask(torch.zeros_like(mask))
    with _cuda_cfg(ref.cfg), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ro, ra = rm(rx, rk, rv)
    torch.autograd.backward([ro, ra], [go, ga])
    o, a = m(x, k, v)
    torch.autograd.backward([o, a], [go, ga])
    assert relmax(x.grad, rx.grad) <= 1e-4 and relmax(k.grad, rk.grad) <= 1e-4 and relmax(v.grad, rv.grad) <= 1e-4


class _StandInEncoder(torch.nn.Module):
    """image_encoder stand-in with CNN_ENCODER's output contract (DAMSM.py:229-230): region features
    [B, 256, 17, 17] and a global code [B, 256], differentiable w.r.t. the images."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.f = torch.nn.Conv2d(3, 256, 3, padding=1)
        self.c = torch.nn.Linear(256, 256)

    def forward(self, x):
        h = torch.nn.functional.adaptive_avg_pool2d(torch.relu(self.f(x)) * 0.3 - 0.05, 17)
        return h, self.c(h.mean(dim=(2, 3)))


@needs_models
def test_trainer_damsm_loss_through_install(cuda_lib):
    """train.py:419-435 unmodified: once on the reference's losses, once with eegan_b200.install()."""
    import warnings
    ref, inst = RL.load_reference_models(), RL.load_reference_installed()
    assert inst.train.words_loss.__module__ == "eegan_b200.damsm_losses" and inst.train.sent_loss.__module__ == "eegan_b200.damsm_losses"
    B, T = 32, 18
    g = cases._gen(21)
    enc = _StandInEncoder().cuda()
    fake = torch.tanh(torch.randn(B, 3, 64, 64, generator=g)).cuda()
    sent = torch.randn(B, 256, generator=g).cuda()
    words = (torch.tanh(torch.randn(B, 256, T, generator=g)) * 0.5).cuda()
    attrs = torch.randn(B, 256, generator=g).cuda()
    lens = torch.randint(5, T + 1, (B,), generator=g).cuda()
    cls = torch.randint(1, 18, (B,), generator=g).numpy()  # train.py:56 hands the class ids over as a numpy array
    labels = torch.arange(B).cuda()
    outs = []
    for ns in (ref, inst, inst, inst):  # the installed route three times: plain launches, capture, replay
        f = fake.clone().requires_grad_()
        a = attrs.clone().requires_grad_()
        with _cuda_cfg(ref.cfg), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            w_loss, s_loss, a_loss = ns.train.Trainer.DAMSM_loss(f, sent, words, a, cls, B, labels, lens, enc)
            (w_loss + s_loss + a_loss).backward()
        outs.append((w_loss.item(), s_loss.item(), a_loss.item(), f.grad.clone(), a.grad.clone()))
    r = outs[0]
    for o in outs[1:]:
        for k in range(3):
            assert abs(o[k] - r[k]) <= 2e-5 * max(1.0, abs(r[k])), (k, o[k], r[k])
        assert relmax(o[3], r[3]) <= 2e-4 and relmax(o[4], r[4]) <= 1e-4


def _randomise(module, seed):
    g = cases._gen(seed)
    with torch.no_grad():
        for p in module.parameters():  # Gen's residual gates and affine heads are zero-initialised: make every path count
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() <= 1 else 0.5 / max(1.0, p[0].numel() ** 0.5)))


@needs_models
def test_gen_with_install_matches_reference_gen(cuda_lib):
    """models.Gen (models.py:183-256): the reference's own class built on its SyncBN (one replica -> F.batch_norm,
    batchnorm.py:50-53) and on eegan_b200's 24 SynchronizedBatchNorm2d layers; same weights and inputs."""
    ref, inst = RL.load_reference_models(), RL.load_reference_installed()
    ngf, B = 16, 4
    Gr = ref.models.Gen(ngf, 100)
    _randomise(Gr, 1)
    Gi = inst.models.Gen(ngf, 100)
    Gi.load_state_dict(Gr.state_dict())  # state_dict keys are identical
    n_bn = sum(1 for m in Gi.modules() if type(m).__module__.startswith("eegan_b200") and "BatchNorm" in type(m).__name__)
    assert n_bn == 24
    Gr, Gi = Gr.cuda().train(), Gi.cuda().train()
    g = cases._gen(2)
    z, sent, attrs = torch.randn(B, 100, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda(), torch.randn(B, 256, generator=g).cuda()
    go = [torch.randn(B, 3, s, s, generator=g).cuda() for s in (64, 128, 256)]
    # The arbiter is the reference Gen in FLOAT64: a 7-stage generator with random weights and 4-sample batch statistics
    # amplifies fp32 rounding, so "equal" means: eegan_b200's fp32 result is as close to float64 as the reference's own fp32 is.
    import copy
    G64 = copy.deepcopy(Gr).double()
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        res = []
        for G, dt in ((G64, torch.float64), (Gr, torch.float32), (Gi, torch.float32)):
            s = sent.to(dt).clone().requires_grad_()
            imgs = G(z.to(dt), s, attrs.to(dt))
            torch.autograd.backward(imgs, [g_.to(dt) for g_ in go])
            res.append(([im.detach().double() for im in imgs], s.grad.double(), {n: p.grad.double() for n, p in G.named_parameters()},
                        {n: b.double() for n, b in G.named_buffers()}))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    (di, ds, dp, db), (ri, rs, rp, rb), (ii, is_, ip, ib) = res

    bad = []

    def held(ours, ref32, f64, what, floor):
        e_ref, e_ours = float((ref32 - f64).abs().max()), float((ours - f64).abs().max())
        if not e_ours <= 4.0 * e_ref + floor * max(1e-30, float(f64.abs().max())):
            bad.append((what, e_ours, e_ref, float(f64.abs().max())))

    for k in range(3):
        held(ii[k], ri[k], di[k], "image %d" % k, 1e-5)
    held(is_, rs, ds, "d sent", 1e-5)
    for n in dp:
        # the scalar residual gates (`gamma`, models.py:103, 135) collect sum(residual * grad) over up to 8.4e6 activations with heavy
        # cancellation, downstream of batch statistics that eegan_b200 forms as the reference's own N-replica code does
        # (sum / square-sum, batchnorm.py:113-125) while the one-replica reference calls cuDNN's batch norm: a wider floor there
        held(ip[n], rp[n], dp[n], n, 1e-3 if n.endswith("gamma") else 5e-5)
    for n in db:
        if "running" in n:
            held(ib[n], rb[n], db[n], n, 1e-5)
    assert not bad, bad


# ---------------------------------------------------------------------------------------
# the planned route of words_loss (eegan_b200/fastpath.py)
# ---------------------------------------------------------------------------------------
def test_planned_route_matches_plain_route_and_survives_interleaving(cuda_lib):
    import eegan_b200 as E
    from eegan_b200 import damsm_losses as dl
    from eegan_b200 import fastpath
    B, T = 24, 12
    c1, c2 = cases.words_case(B, T, seed=1), cases.words_case(B, T, seed=2)

    def plain(c):
        img, words = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
        m, _ = dl.pair_grid(img, words, c["cap_lens"].cuda())
        l0, l1, _ = dl._PairCEFn.apply(m, 10.0, c["class_ids"].cuda(), c["labels"].cuda())
        (l0 + 3 * l1).backward()
        return l0.item(), l1.item(), img.grad, words.grad

    p1, p2 = plain(c1), plain(c2)
    fastpath.clear_plans()
    for rep in range(4):
        # two forwards in flight on ONE plan, backwards in the opposite order: the generation counter re-runs forward 1
        ins = []
        for c in (c1, c2):
            img, words = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
            l0, l1, att = E.words_loss(img, words, c["labels"].cuda(), c["cap_lens"].cuda(), c["class_ids"], B)
            ins.append((img, words, l0, l1))
        for (img, words, l0, l1), p in ((ins[1], p2), (ins[0], p1)):
            (l0 + 3 * l1).backward(retain_graph=True)
            assert abs(l0.item() - p[0]) <= 1e-6 and abs(l1.item() - p[1]) <= 1e-6
            assert relmax(img.grad, p[2]) <= 1e-6 and relmax(words.grad, p[3]) <= 1e-6, rep
        # a second backward on a retained graph accumulates the same gradient again
        img, words, l0, l1 = ins[0]
        (l0 + 3 * l1).backward()
        assert relmax(img.grad, 2 * p1[2]) <= 1e-6
    plan = next(iter(fastpath._plans.values()))
    assert plan.fwd_graph is not None and (True, True) in plan.bwd_graphs
    # only the image gradient (train.py:172 detaches the words): its own backward graph on the same plan
    for rep in range(3):
        img = c1["img"].cuda().requires_grad_()
        l0, l1, _ = E.words_loss(img, c1["words"].cuda(), c1["labels"].cuda(), c1["cap_lens"].cuda(), c1["class_ids"], B)
        (l0 + 3 * l1).backward()
        assert relmax(img.grad, p1[2]) <= 1e-6
    assert (True, False) in plan.bwd_graphs and len(fastpath._plans) == 1
    # outputs are fresh tensors: a later call must not change an earlier call's gradient or attention maps
    img = c1["img"].cuda().requires_grad_()
    l0, l1, att = E.words_loss(img, c1["words"].cuda(), c1["labels"].cuda(), c1["cap_lens"], c1["class_ids"], B)
    (l0 + 3 * l1).backward()
    keep_g, keep_a = img.grad.clone(), [a.clone() for a in att]
    img2 = c2["img"].cuda().requires_grad_()
    m0, m1, _ = E.words_loss(img2, c2["words"].cuda(), c2["labels"].cuda(), c2["cap_lens"], c2["class_ids"], B)
    (m0 + m1).backward()
    assert torch.equal(img.grad, keep_g) and all(torch.equal(a, b) for a, b in zip(att, keep_a))

"""TEST INFRASTRUCTURE — seeded synthetic inputs shared by make_golden.py, tests/ and bench.py.

Shapes follow SURVEY.md §8(d): D=256 (miscc/config.py:65), 17x17 regions (DAMSM.py:197-210),
T<=20 words (miscc/config.py:66), seed 3407 (the reference's own default, train.py:50).
All tensors are generated on the CPU generator so the CPU oracle and the GPU path see
identical bits.
"""
from __future__ import annotations

import torch

SEED = 3407


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def words_case(B, T, D=256, H=17, kind="realistic", class_mode="cub", seed=SEED, min_len=5):
    """Returns dict(img [B,D,H,H], words [B,D,T], cap_lens [B] int64, labels, class_ids|None)."""
    g = _gen(seed)
    if kind == "realistic":
        img = torch.relu(torch.randn(B, D, H, H, generator=g)) * 0.3 - 0.05
        words = torch.tanh(torch.randn(B, D, T, generator=g)) * 0.5
    elif kind == "stress":
        img = torch.randn(B, D, H, H, generator=g)
        words = torch.randn(B, D, T, generator=g)
    else:
        raise ValueError(kind)
    lo = min(min_len, T)
    cap = torch.randint(lo, T + 1, (B,), generator=g)
    cap[0] = T
    if B > 1:
        cap[1] = lo
    if class_mode == "cub":  # 200 classes -> collisions -> -inf cells (datasets.py class ids)
        cls = torch.randint(1, max(2, min(201, B // 2 + 2)), (B,), generator=g)
    elif class_mode == "unique":
        cls = torch.arange(B)
    elif class_mode == "none":
        cls = None
    else:
        raise ValueError(class_mode)
    return dict(img=img, words=words, cap_lens=cap, labels=torch.arange(B), class_ids=cls)


def sent_case(B, D=256, class_mode="cub", seed=SEED):
    g = _gen(seed + 1)
    cnn = torch.randn(B, D, generator=g)
    rnn = torch.randn(B, D, generator=g)
    if class_mode == "cub":
        cls = torch.randint(1, max(2, min(201, B // 2 + 2)), (B,), generator=g)
    elif class_mode == "unique":
        cls = torch.arange(B)
    else:
        cls = None
    return dict(cnn=cnn, rnn=rnn, labels=torch.arange(B), class_ids=cls)


def gag_case(B, idf, H, T, seed=SEED, masked=True, min_len=5):
    """GlobalAttentionGeneral inputs: x [B,idf,H,H], key/value [B,idf,T], mask [B,T] (True=pad)."""
    g = _gen(seed + 2)
    x = torch.randn(B, idf, H, H, generator=g)
    key = torch.randn(B, idf, T, generator=g) * (idf ** -0.5)
    val = torch.randn(B, idf, T, generator=g)
    cap = torch.randint(min(min_len, T), T + 1, (B,), generator=g)
    cap[0] = T
    mask = (torch.arange(T)[None, :] >= cap[:, None]) if masked else None
    return dict(x=x, key=key, value=val, mask=mask, cap_lens=cap)


def bn_case(N, C, H, seed=SEED):
    g = _gen(seed + 3)
    x = torch.randn(N, C, H, H, generator=g) * 1.7 + 0.4
    w = torch.rand(C, generator=g) + 0.5
    b = torch.randn(C, generator=g) * 0.1
    return dict(x=x, weight=w, bias=b)


def emb_case(B, Cin=768, Cout=256, H=17, seed=SEED):
    """CNN_ENCODER.emb_features inputs: weight ~ uniform(-0.1, 0.1) (DAMSM.py:166-167), x = relu(randn) like the
    Inception Mixed_6e map, an upstream gradient go."""
    g = _gen(seed)
    weight = (torch.rand(Cout, Cin, 1, 1, generator=g) - 0.5) * 0.2
    x = torch.relu(torch.randn(B, Cin, H, H, generator=g))
    go = torch.randn(B, Cout, H, H, generator=g)
    return dict(weight=weight, x=x, go=go)


def checksum(t: torch.Tensor):
    t = t.double()
    return [float(t.sum()), float(t.abs().sum())]

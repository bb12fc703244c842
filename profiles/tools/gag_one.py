"""GlobalAttentionGeneral at one shape: forward engines 0 / 1, backward engines 0 (CUDA cores) / 1 (tcgen05 one-pass kernel).

    python profiles/tools/gag_one.py RES IDF [B] [--check] [--iters N] [--only-bwd-engine E] [--no-dattn]

--check compares d_x / d_key / d_value of both backward engines with a float64 autograd run of the same math on the GPU
(relative-to-max errors), which is what tells the four contractions of gag_tc_bwd.cu apart: d_value depends only on (2),
d_key on (1) -> ds -> (4), d_x on (1) -> ds -> (3).  Timings are CUDA-graph replays with an L2 flush between them.
Also the ncu target for the GAG kernels (-k regex:gag)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def ref64(x, key, val, mask, go, ga):
    """float64 restatement of miscc/DAMSM_losses.py:96-132 (mask_mode = the reference quirk is irrelevant here: mask=None or
    the intended per-sample mask is applied through the module under test as well)."""
    x, key, val = (t.detach().double().requires_grad_() for t in (x, key, val))
    B, idf, H, W = x.shape
    Q = H * W
    s = torch.bmm(x.view(B, idf, Q).transpose(1, 2), key)  # [B, Q, T]
    if mask is not None:
        s = s.masked_fill(mask[:, None, :], float("-inf"))
    p = torch.softmax(s, dim=2)
    out = torch.bmm(val, p.transpose(1, 2)).view(B, idf, H, W)
    attn = p.transpose(1, 2).reshape(B, -1, H, W)
    loss = (out * go.double()).sum()
    if ga is not None:
        loss = loss + (attn * ga.double()).sum()
    loss.backward()
    return x.grad, key.grad, val.grad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("res", type=int)
    ap.add_argument("idf", type=int)
    ap.add_argument("B", type=int, nargs="?", default=48)
    ap.add_argument("--T", type=int, default=18)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only-bwd-engine", type=int, default=-1)
    ap.add_argument("--fwd-engine", type=int, default=-1)
    ap.add_argument("--flush", default="write", choices=["write", "read", "both"])
    ap.add_argument("--no-dattn", action="store_true")
    ap.add_argument("--no-mask", action="store_true")
    a = ap.parse_args()
    import eegan_b200 as E
    from eegan_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    B, idf, res, T = a.B, a.idf, a.res, a.T
    lens = torch.randint(5, T + 1, (B,)) if T > 5 else torch.full((B,), T)
    mask = None if a.no_mask else (torch.arange(T)[None, :] >= lens[:, None]).to(dev)
    x = torch.randn(B, idf, res, res, device=dev).requires_grad_()
    key = (torch.randn(B, idf, T, device=dev) * idf ** -0.5).requires_grad_()
    val = torch.randn(B, idf, T, device=dev).requires_grad_()
    go = torch.randn(B, idf, res, res, device=dev)
    ga = None if a.no_dattn else torch.randn(B, T, res, res, device=dev)
    mod = E.GlobalAttentionGeneral(idf, 256, mask_mode="intended")
    if mask is not None:
        mod.applyMask(mask)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def fwd_bwd():
        x.grad = key.grad = val.grad = None
        o, at = mod(x, key, val)
        if ga is None:
            torch.autograd.backward([o], [go])
        else:
            torch.autograd.backward([o, at], [go, ga])

    def graph(fn):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay

    sweep = torch.empty(256 * 1024 * 1024 // 4, device=dev) if a.flush != "write" else None

    def time(fn, iters):
        ms = 0.0
        for _ in range(iters):
            if a.flush != "read":
                flush.fill_(1.0)
            if sweep is not None:
                sweep.sum()  # a read sweep larger than L2: the cache is cold AND holds no dirty lines to write back in the timed span
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms / iters

    rows = B * res * res
    by = (5 * idf + 2 * T) * 4 * rows
    by_f = (2 * idf + T) * 4 * rows
    grads = {}
    engines = (0, 1) if a.only_bwd_engine < 0 else (a.only_bwd_engine,)
    for eng in engines:
        assert L.eegan_set_gag_bwd_engine(eng) == 0
        fwd_bwd()
        torch.cuda.synchronize()
        grads[eng] = (x.grad.clone(), key.grad.clone(), val.grad.clone())
        ms = time(graph(fwd_bwd), a.iters)
        print("bwd engine %d: fwd+bwd %.1f us  %.1f%% of 6549 GB/s (algorithmic %.0f MB)" % (eng, ms * 1e3, by / (ms / 1e3) / 1e9 / 6549 * 100, by / 1e6),
              flush=True)
    fwd_engines = (0, 1) if a.fwd_engine < 0 else (a.fwd_engine,)
    for fe in fwd_engines:
        assert L.eegan_set_gag_engine(fe) == 0
        with torch.no_grad():
            ms_f = time(graph(lambda: mod(x, key, val)), a.iters)
            o, at = mod(x, key, val)
        msg = ""
        if a.check:
            xd, kd, vd = x.detach().double().view(B, idf, -1), key.detach().double(), val.detach().double()
            sc = torch.bmm(xd.transpose(1, 2), kd)
            if mask is not None:
                sc = sc.masked_fill(mask[:, None, :], float("-inf"))
            pr = torch.softmax(sc, dim=2)
            ro = torch.bmm(vd, pr.transpose(1, 2))
            msg = "  attn abs err %.2e  out rel err %.2e" % (float((at.double().view(B, T, -1) - pr.transpose(1, 2)).abs().max()),
                                                          float((o.double().view(B, idf, -1) - ro).abs().max() / ro.abs().max()))
        print("fwd engine %d: %.1f us  %.1f%%%s" % (fe, ms_f * 1e3, by_f / (ms_f / 1e3) / 1e9 / 6549 * 100, msg), flush=True)
    L.eegan_set_gag_engine(1)
    if a.check:
        rx, rk, rv = ref64(x, key, val, mask, go, ga)
        rel = lambda u, v: float((u.double() - v).abs().max() / v.abs().max())
        for eng, (gx, gk, gv) in grads.items():
            print("bwd engine %d vs float64: d_x %.2e  d_key %.2e  d_value %.2e  (finite: %s)" % (
                eng, rel(gx, rx), rel(gk, rk), rel(gv, rv), bool(torch.isfinite(gx).all() and torch.isfinite(gk).all() and torch.isfinite(gv).all())))
    L.eegan_set_gag_bwd_engine(1)


if __name__ == "__main__":
    main()

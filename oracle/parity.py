"""TEST INFRASTRUCTURE — the parity block `bench.py` prints before it times anything.

Every rank evaluates, on a small seeded GLOBAL batch, what the timed step evaluates on the big one and holds it to
the float64 oracle (oracle/damsm_oracle.py: the dense restatement of miscc/DAMSM_losses.py:272-342 pinned to the
unmodified reference by tests/test_oracle_vs_reference.py):

  * the caption-row-sharded words_loss step (eegan_b200.sharded.ShardedWordsLossStep — the class bench.py times at
    N > 1; at N = 1 the drop-in `words_loss` + backward): loss0, loss1, d_img and d_words of the rank's shard against
    the full-batch float64 result, i.e. against what the reference computes on GPU 0 after DataParallel gathers
    (train.py:195, 419-435);
  * sharded sent_loss (both gradients);
  * one SynchronizedBatchNorm2d layer (forward output, dx, running statistics) against the N-replica formula of
    sync_batchnorm/batchnorm.py:113-125 on the full batch — over NCCL when N > 1.
Errors are max-reduced over the ranks; the block carries its own tolerances and an `ok` flag.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import cases
from . import damsm_oracle as O

TOL = {"loss_rel": 2e-5, "grad_rel_to_max": 1e-4, "att_abs": 1e-6, "syncbn_y_abs": 2e-5, "syncbn_dx_rel": 1e-4,
       "syncbn_running_rel": 1e-5}


def _relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def run(world: int, rank: int, dev, b: int = None, T: int = 12, mode: str = "serial", shard: str = None):
    """Returns the parity dict (identical on every rank).  `mode` selects the sharded step class under test."""
    import eegan_b200 as E
    from eegan_b200 import sharded
    if b is None:
        b = 16 if world == 1 else max(4, 32 // world)
    B = b * world
    sl = slice(rank * b, (rank + 1) * b)
    c = cases.words_case(B, T, seed=4242, class_mode="cub")
    w0, w1 = 1.0, 0.5
    io, wo = c["img"].double().requires_grad_(), c["words"].double().requires_grad_()
    o0, o1, oatt, _ = O.dense_words_loss(io, wo, c["labels"], c["cap_lens"], c["class_ids"])
    (w0 * o0 + w1 * o1).backward()
    err = {}
    if world == 1:
        img, words = c["img"].to(dev).requires_grad_(), c["words"].to(dev).requires_grad_()
        for _ in range(3):  # plain launches, capture, replay: the third call is the planned route's steady state
            img.grad = words.grad = None
            l0, l1, att = E.words_loss(img, words, c["labels"].to(dev), c["cap_lens"].to(dev), c["class_ids"], B)
            (w0 * l0 + w1 * l1).backward()
        d_img, d_words = img.grad, words.grad
        att_err = max(float((a.cpu().double() - r.detach()).abs().max()) for a, r in zip(att, oatt))
        what = "eegan_b200.words_loss + backward (planned route, 3rd call)"
    else:
        cls_map = {"serial": sharded.ShardedWordsLossStep, "graph": sharded.ShardedWordsLossStep,
                   "overlap": sharded.OverlappedShardedWordsLossStep}
        kw = dict(w0=w0, w1=w1)
        if mode == "graph":
            kw["graph"] = True
        if mode != "overlap":
            kw["shard"] = shard  # None = the package default (eegan_b200.sharded.SHARD_BY)
        step = cls_map[mode](b, c["img"].shape[1], c["img"].shape[2], c["img"].shape[3], T, dev, **kw)
        for _ in range(2):
            l0, l1, d_img, d_words = step(c["img"][sl].to(dev), c["words"][sl].to(dev), c["cap_lens"][sl].to(dev),
                                          c["class_ids"][sl].to(dev))
        att_err = 0.0
        for i in range(b):
            Ti = int(c["cap_lens"][sl][i])
            att_err = max(att_err, float((step.att[i, :Ti].cpu().double().reshape(-1) - oatt[rank * b + i].detach().reshape(-1)).abs().max()))
        what = "eegan_b200.sharded.%s (%s, grid partitioned by %s)" % (type(step).__name__, mode, getattr(step, "shard", "captions"))
        if mode == "graph":
            step.release_graph()
    err["loss0_rel"] = abs(float(l0) - float(o0)) / max(1.0, abs(float(o0)))
    err["loss1_rel"] = abs(float(l1) - float(o1)) / max(1.0, abs(float(o1)))
    # gradients: relative to the largest element of the FULL-batch gradient (SURVEY.md §8d "rel-to-max")
    err["d_img_rel"] = float((d_img.detach().cpu().double().reshape(b, -1) - io.grad[sl].reshape(b, -1)).abs().max() / io.grad.abs().max())
    err["d_words_rel"] = float((d_words.detach().cpu().double() - wo.grad[sl]).abs().max() / wo.grad.abs().max())
    err["att_abs"] = att_err
    # sentence loss, sharded
    s = cases.sent_case(B, seed=77)
    cn, rn = s["cnn"].double().requires_grad_(), s["rnn"].double().requires_grad_()
    q0, q1 = O.port_sent_loss(cn, rn, s["labels"], s["class_ids"], B)
    (q0 + q1).backward()
    gc, gr = s["cnn"][sl].to(dev).requires_grad_(), s["rnn"][sl].to(dev).requires_grad_()
    s0, s1 = sharded.sharded_sent_loss(gc, gr, torch.arange(b, device=dev), s["class_ids"][sl], b)
    (s0 + s1).backward()
    err["sent_loss_rel"] = max(abs(float(s0) - float(q0)) / max(1.0, abs(float(q0))), abs(float(s1) - float(q1)) / max(1.0, abs(float(q1))))
    err["sent_grad_rel"] = max(float((gc.grad.cpu().double() - cn.grad[sl]).abs().max() / cn.grad.abs().max()),
                               float((gr.grad.cpu().double() - rn.grad[sl]).abs().max() / rn.grad.abs().max()))
    # one SyncBN layer: 4 samples per rank, full batch = 4 * world, N-replica formula (clamp) when world > 1
    from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
    bc = cases.bn_case(4 * world, 32, 16, seed=9)
    go = torch.randn(4 * world, 32, 16, 16, generator=cases._gen(10))
    bn = SynchronizedBatchNorm2d(32).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(bc["weight"]); bn.bias.copy_(bc["bias"])
    s4 = slice(4 * rank, 4 * rank + 4)
    xs = bc["x"][s4].to(dev).requires_grad_()
    y = bn(xs)
    y.backward(go[s4].to(dev))
    xf = bc["x"].double().requires_grad_()
    if world > 1:
        outs, mean, inv_std, unb = O.syncbn_forward([xf], bc["weight"].double(), bc["bias"].double())
        yf = outs[0]
    else:  # one replica: F.batch_norm (sync_batchnorm/batchnorm.py:50-53)
        yf = torch.nn.functional.batch_norm(xf, None, None, bc["weight"].double(), bc["bias"].double(), True, 0.1, 1e-5)
        mean = xf.detach().mean(dim=(0, 2, 3))
        unb = xf.detach().transpose(0, 1).reshape(32, -1).var(dim=1, unbiased=True)
    yf.backward(go.double())
    err["syncbn_y_abs"] = float((y.detach().cpu().double() - yf.detach()[s4]).abs().max())
    err["syncbn_dx_rel"] = float((xs.grad.cpu().double() - xf.grad[s4]).abs().max() / xf.grad.abs().max())
    err["syncbn_running_rel"] = max(_relmax(bn.running_mean, 0.1 * mean), _relmax(bn.running_var, 0.9 + 0.1 * unb))
    keys = sorted(err)
    t = torch.tensor([err[k] for k in keys], dtype=torch.float64, device=dev)
    t = torch.nan_to_num(t, nan=1e30, posinf=1e30)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = {k: float(v) for k, v in zip(keys, t.tolist())}
    ok = (err["loss0_rel"] <= TOL["loss_rel"] and err["loss1_rel"] <= TOL["loss_rel"] and err["sent_loss_rel"] <= TOL["loss_rel"]
          and err["d_img_rel"] <= TOL["grad_rel_to_max"] and err["d_words_rel"] <= TOL["grad_rel_to_max"]
          and err["sent_grad_rel"] <= TOL["grad_rel_to_max"] and err["att_abs"] <= TOL["att_abs"]
          and err["syncbn_y_abs"] <= TOL["syncbn_y_abs"] and err["syncbn_dx_rel"] <= TOL["syncbn_dx_rel"]
          and err["syncbn_running_rel"] <= TOL["syncbn_running_rel"])
    return {"ok": bool(ok), "against": "float64 oracle of the full global batch (oracle/damsm_oracle.py), max over ranks",
            "global_batch": B, "per_rank": b, "T_max": T, "path": what,
            "collectives": "nccl" if world > 1 else "none", "errors": err, "tolerances": TOL}

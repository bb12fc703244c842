"""CPU, world_size = 2 over gloo: the host-side sharding / collective logic of the N > 1 path
(eegan_b200/sharded.py and the SyncBN module) with the ORACLE standing in for the CUDA kernels.
The sharded result must equal the single-process full-batch result (the reference computes
the full-batch grid on GPU 0 after DataParallel gathers, train.py:195,419-435)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_grid(img_all, words, cap_lens, diag_offset=0, want_att=True):
    """(m [B_img, B_cap], diagonal attention [B_cap, T, R]) as the library returns them: caption i's own image is
    j = i + diag_offset; captions whose image is not in the block get zeros (eegan_damsm_pair_fwd)."""
    from oracle import damsm_oracle as O
    t = O.dense_pair_terms(img_all, words, cap_lens)
    Bi, Bc = img_all.shape[0], words.shape[0]
    zero = torch.zeros_like(t["a"][0, 0])
    att = torch.stack([t["a"][diag_offset + i, i] if 0 <= diag_offset + i < Bi else zero for i in range(Bc)])  # [Bc, T, R]
    return t["m"], att.detach()


def _oracle_ce(m_all, g3, cls_all, lab_all):
    from oracle import damsm_oracle as O
    sim = m_all * g3
    mask = O.class_mask(cls_all, sim.shape[0])
    if mask is not None:
        sim = sim.masked_fill(mask, float("-inf"))
    return O._two_way_ce(sim, lab_all)


class OracleBNOps:
    """CPU stand-in for the SyncBN kernels (same five steps, torch ops)."""

    @staticmethod
    def stats(x3, out):
        C = x3.shape[1]
        out[:C] = x3.sum(dim=(0, 2))
        out[C:2 * C] = (x3 * x3).sum(dim=(0, 2))

    @staticmethod
    def finalize(stats, C, count, count_dev, eps, momentum, clamp_mode, mean, inv_std, rm, rv):
        n = float(count_dev[0] * 4096 + count_dev[1]) if count_dev is not None else float(count)
        s, ss = stats[:C], stats[C:2 * C]
        mu = s / n
        sumvar = ss - s * mu
        mean.copy_(mu)
        inv_std.copy_((sumvar / n).clamp(min=eps) ** -0.5 if clamp_mode else 1.0 / torch.sqrt(sumvar / n + eps))
        if rm is not None:
            rm.mul_(1 - momentum).add_(momentum * mu)
            rv.mul_(1 - momentum).add_(momentum * sumvar / (n - 1))

    @staticmethod
    def apply(x3, mean, inv_std, w, b, y):
        scale = inv_std * (w if w is not None else 1.0)
        y.copy_((x3 - mean.view(1, -1, 1)) * scale.view(1, -1, 1) + (b.view(1, -1, 1) if b is not None else 0.0))

    @staticmethod
    def bwd_reduce(x3, dy3, mean, inv_std, red):
        C = x3.shape[1]
        xh = (x3 - mean.view(1, -1, 1)) * inv_std.view(1, -1, 1)
        red[:C] = dy3.sum(dim=(0, 2))
        red[C:] = (dy3 * xh).sum(dim=(0, 2))

    @staticmethod
    def bwd_apply(x3, dy3, mean, inv_std, w, red, count, count_dev, eps, clamp_mode, dx):
        n = float(count_dev[0] * 4096 + count_dev[1]) if count_dev is not None else float(count)
        C = x3.shape[1]
        xh = (x3 - mean.view(1, -1, 1)) * inv_std.view(1, -1, 1)
        scale = (inv_std * (w if w is not None else 1.0)).view(1, -1, 1)
        dx.copy_(scale * (dy3 - red[:C].view(1, -1, 1) / n - xh * red[C:].view(1, -1, 1) / n))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        from eegan_b200 import sharded
        from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
        from oracle import cases
        from oracle import damsm_oracle as O
        B, T, b = 8, 7, 4
        c = cases.words_case(B, T, D=16, H=3, seed=21, min_len=2)
        sl = slice(rank * b, (rank + 1) * b)
        img = c["img"][sl].double().requires_grad_()
        words = c["words"][sl].double().requires_grad_()
        # the caption partition first (own captions x all images), checked on its own; then the default image partition
        # (own images x all captions), which the result dict below carries
        l0, l1, att = sharded.sharded_words_loss(img, words, torch.arange(b), c["cap_lens"][sl], c["class_ids"][sl], b,
                                                 grid_fn=_oracle_grid, ce_fn=_oracle_ce, shard="captions")
        (l0 + 2 * l1).backward()
        cap_part = (l0.detach().clone(), l1.detach().clone(), img.grad.clone(), words.grad.clone(), [a.clone() for a in att])
        img.grad = words.grad = None
        l0, l1, att = sharded.sharded_words_loss(img, words, torch.arange(b), c["cap_lens"][sl], c["class_ids"][sl], b,
                                                 grid_fn=_oracle_grid, ce_fn=_oracle_ce, shard="images")
        (l0 + 2 * l1).backward()
        assert sharded.SHARD_BY == "images"
        assert float((cap_part[0] - l0.detach()).abs() + (cap_part[1] - l1.detach()).abs()) < 1e-12
        assert float((cap_part[2] - img.grad).abs().max()) < 1e-12 and float((cap_part[3] - words.grad).abs().max()) < 1e-12
        assert len(att) == len(cap_part[4]) == b and all(float((a - q).abs().max()) < 1e-12 for a, q in zip(att, cap_part[4]))
        # single-process full batch
        fi = c["img"].double().requires_grad_()
        fw = c["words"].double().requires_grad_()
        f0, f1, fatt, _ = O.dense_words_loss(fi, fw, c["labels"], c["cap_lens"], c["class_ids"])
        (f0 + 2 * f1).backward()
        res = dict(
            loss=abs(float(l0.detach() - f0.detach())) + abs(float(l1.detach() - f1.detach())),
            dimg=float((img.grad - fi.grad[sl]).abs().max()), dwords=float((words.grad - fw.grad[sl]).abs().max()),
            att=max(float((a - fa.detach()).abs().max()) for a, fa in zip(att, fatt[sl])), natt=len(att))
        # sentence loss
        s = cases.sent_case(B, D=16, seed=4)
        cn = s["cnn"][sl].double().requires_grad_()
        rn = s["rnn"][sl].double().requires_grad_()
        s0, s1 = sharded.sharded_sent_loss(cn, rn, torch.arange(b), s["class_ids"][sl], b, loss_fn=O.port_sent_loss)
        (s0 + s1).backward()
        fc = s["cnn"].double().requires_grad_()
        fr = s["rnn"].double().requires_grad_()
        g0, g1 = O.port_sent_loss(fc, fr, s["labels"], s["class_ids"], B)
        (g0 + g1).backward()
        res.update(sent=abs(float(s0.detach() - g0.detach())) + abs(float(s1.detach() - g1.detach())),
                   dcnn=float((cn.grad - fc.grad[sl]).abs().max()), drnn=float((rn.grad - fr.grad[sl]).abs().max()))
        # SyncBN: 2 ranks x 3 samples == the N-replica formula on the 6-sample batch
        bc = cases.bn_case(6, 8, 5)
        bn = SynchronizedBatchNorm2d(8)
        bn._ops = OracleBNOps
        with torch.no_grad():
            bn.weight.copy_(bc["weight"]); bn.bias.copy_(bc["bias"])
        xs = bc["x"][rank * 3:(rank + 1) * 3].clone().requires_grad_()
        y = bn(xs)
        gy = torch.randn(6, 8, 5, 5, generator=cases._gen(1))
        (y * gy[rank * 3:(rank + 1) * 3]).sum().backward()
        xf = bc["x"].clone().requires_grad_()
        outs, mean, inv_std, unb = O.syncbn_forward([xf], bc["weight"], bc["bias"])
        (outs[0] * gy).sum().backward()
        res.update(bn_y=float((y.detach() - outs[0].detach()[rank * 3:(rank + 1) * 3]).abs().max()),
                   bn_dx=float((xs.grad - xf.grad[rank * 3:(rank + 1) * 3]).abs().max()),
                   bn_rm=float((bn.running_mean - 0.1 * mean.detach()).abs().max()),
                   bn_rv=float((bn.running_var - (0.9 + 0.1 * unb.detach())).abs().max()))
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_world2_sharded_equals_full_batch():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        r = out[rank]
        assert r["natt"] == 4
        assert r["loss"] < 1e-12 and r["dimg"] < 1e-12 and r["dwords"] < 1e-12 and r["att"] < 1e-14, r
        assert r["sent"] < 1e-12 and r["dcnn"] < 1e-12 and r["drnn"] < 1e-12, r
        assert r["bn_y"] < 1e-5 and r["bn_dx"] < 1e-4 and r["bn_rm"] < 1e-6 and r["bn_rv"] < 1e-5, r


class OracleStepKernels:
    """CPU stand-in for the C ABI calls of eegan_b200.sharded.OverlappedShardedWordsLossStep (float64 oracle)."""

    def workspace(self, Bi, Bc, D, R, Tm, device):
        return None

    def pair_fwd(self, img, words, lens32, Bi, Bc, D, R, Tm, m, att, ws):
        from oracle import damsm_oracle as O
        t = O.dense_pair_terms(img, words, lens32.long())
        m.copy_(t["m"])
        if att is not None:
            for i in range(Bc):
                att[i] = t["a"][i, i]

    def pair_bwd(self, img, words, lens32, Bi, Bc, D, R, Tm, dm, d_img, d_words, ws):
        from oracle import damsm_oracle as O
        di, dw = O.dense_words_backward(img, words, lens32.long(), dm)
        d_img.copy_(di)
        if d_words is not None:
            d_words.copy_(dw)

    def ce(self, m_all, cls_all, labels, gvec, Bt, sim, lse, loss01, dsim):
        from eegan_b200.config import gammas
        from oracle import damsm_oracle as O
        g3 = gammas()[2]
        s = m_all * g3
        mask = O.class_mask(cls_all, Bt) if cls_all is not None else None
        if mask is not None:
            s = s.masked_fill(mask, float("-inf"))
        l0, l1 = O._two_way_ce(s, labels)
        loss01[0], loss01[1] = l0, l1
        dsim.copy_(O.ce_pair_grad(s, labels, float(gvec[0]), float(gvec[1])) * g3)


def _overlap_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        from eegan_b200.sharded import OverlappedShardedWordsLossStep
        from oracle import cases
        from oracle import damsm_oracle as O
        b, T, D, H = 3, 6, 16, 3
        B = b * world
        c = cases.words_case(B, T, D=D, H=H, seed=33, min_len=2)
        sl = slice(rank * b, (rank + 1) * b)
        step = OverlappedShardedWordsLossStep(b, D, H, H, T, "cpu", w0=1.0, w1=2.0, kernels=OracleStepKernels(), dtype=torch.float64)
        res = {}
        for it in range(2):  # twice: the static buffers (zero slot, rotated gradients) must survive a step
            l0, l1, d_img, d_words = step(c["img"][sl].double(), c["words"][sl].double(), c["cap_lens"][sl], c["class_ids"][sl])
            fi = c["img"].double().requires_grad_()
            fw = c["words"].double().requires_grad_()
            f0, f1, fatt, _ = O.dense_words_loss(fi, fw, c["labels"], c["cap_lens"], c["class_ids"])
            (f0 + 2 * f1).backward()
            lens = c["cap_lens"][sl].tolist()
            res["it%d" % it] = dict(
                loss=abs(float(l0 - f0.detach())) + abs(float(l1 - f1.detach())),
                dimg=float((d_img - fi.grad[sl]).abs().max()), dwords=float((d_words - fw.grad[sl]).abs().max()),
                att=max(float((step.att[i, :lens[i]].reshape(1, lens[i], H, H) - fatt[rank * b + i].detach()).abs().max()) for i in range(b)))
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_overlapped_sharded_step_equals_full_batch(world):
    """Host logic of the opt-in overlapped step: rank-rotated gather buffer, local / remote image blocks, row permutations
    of m / dm, reduce-scatter with a zero own slot, partial d_words — world 3 puts a rank in the middle of the rotation."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_overlap_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        for it in ("it0", "it1"):
            r = out[rank][it]
            assert r["loss"] < 1e-12 and r["dimg"] < 1e-12 and r["dwords"] < 1e-12 and r["att"] < 1e-14, (rank, it, r)


# ---------------------------------------------------------------------------------------
# DataParallelWithCallback under a process group: parameter gradients are summed over ranks
# (what nn.DataParallel does over its replicas, sync_batchnorm/replicate.py:50-67 + train.py:220)
# ---------------------------------------------------------------------------------------
def _tiny_net():
    from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 1), SynchronizedBatchNorm2d(8), torch.nn.ReLU(), torch.nn.Conv2d(8, 2, 1))
    net[1]._ops = OracleBNOps
    return net


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        from eegan_b200.sync_batchnorm import DataParallelWithCallback
        g = torch.Generator().manual_seed(11)
        x = torch.randn(6, 3, 4, 4, generator=g)
        gy = torch.randn(6, 2, 4, 4, generator=g)
        sl = slice(rank * 3, (rank + 1) * 3)
        net = DataParallelWithCallback(_tiny_net())
        assert hasattr(net, "module")
        (net(x[sl]) * gy[sl]).sum().backward()  # the rank's share of the global-batch loss
        # single process, full batch, the N-replica statistics formula (batchnorm.py:113-125) in plain torch
        ref = _tiny_net()
        h = ref[0](x)
        mu = h.mean(dim=(0, 2, 3), keepdim=True)
        var = ((h - mu) ** 2).mean(dim=(0, 2, 3), keepdim=True)
        hn = (h - mu) * var.clamp(min=ref[1].eps) ** -0.5
        hn = hn * ref[1].weight.view(1, -1, 1, 1) + ref[1].bias.view(1, -1, 1, 1)
        (ref[3](torch.relu(hn)) * gy).sum().backward()
        # (the bias of the conv in front of the batch norm has an identically-zero gradient: compare on the scale of
        # the largest parameter gradient, not per tensor)
        scale = max(float(q.grad.abs().max()) for q in ref.parameters())
        err = max(float((p.grad - q.grad).abs().max()) for p, q in zip(net.module.parameters(), ref.parameters())) / scale
        # grad_sync=None leaves the local (per-rank) gradients alone: they differ from the full-batch ones
        net2 = DataParallelWithCallback(_tiny_net(), grad_sync=None)
        (net2(x[sl]) * gy[sl]).sum().backward()
        local = float((net2.module[0].weight.grad - ref[0].weight.grad).abs().max())
        out[rank] = dict(err=err, local=local)
    finally:
        dist.destroy_process_group()


def test_world2_data_parallel_wrapper_sums_parameter_gradients():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        assert out[rank]["err"] < 2e-5, dict(out[rank])
        assert out[rank]["local"] > 1e-3, dict(out[rank])

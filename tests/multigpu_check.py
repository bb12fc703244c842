"""Run under torchrun on N GPUs of one box (not collected by pytest: needs N devices):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py

Checks, over NCCL, that the caption-row-sharded words_loss / sent_loss and the NCCL
SynchronizedBatchNorm2d equal the single-device full-batch result (computed redundantly on every
rank by the same CUDA library) — losses, both gradients, attention maps, BN output and grads."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import eegan_b200 as E  # noqa: E402
from eegan_b200.sharded import sharded_sent_loss, sharded_words_loss  # noqa: E402
from oracle import cases  # noqa: E402  (seeded input generator only)


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b, T = 12, 18
    B = b * world
    c = cases.words_case(B, T, seed=11, class_mode="cub")
    sl = slice(rank * b, (rank + 1) * b)
    # full batch on this device
    img = c["img"].to(dev).requires_grad_()
    words = c["words"].to(dev).requires_grad_()
    f0, f1, fatt = E.words_loss(img, words, c["labels"].to(dev), c["cap_lens"].to(dev), c["class_ids"], B)
    (f0 + 2.0 * f1).backward()
    # sharded, both partitions of the grid (own images x all captions: the default; own captions x all images)
    from eegan_b200.sharded import ShardedWordsLossStep
    for shard in ("images", "captions"):
        img_s = c["img"][sl].to(dev).requires_grad_()
        words_s = c["words"][sl].to(dev).requires_grad_()
        s0, s1, satt = sharded_words_loss(img_s, words_s, torch.arange(b, device=dev), c["cap_lens"][sl].to(dev),
                                          c["class_ids"][sl], b, shard=shard)
        (s0 + 2.0 * s1).backward()
        assert abs(s0.item() - f0.item()) <= 2e-6 * max(1.0, abs(f0.item())), (shard, s0.item(), f0.item())
        assert abs(s1.item() - f1.item()) <= 2e-6 * max(1.0, abs(f1.item())), (shard, s1.item(), f1.item())
        assert relmax(img_s.grad, img.grad[sl]) <= 2e-5, (shard, relmax(img_s.grad, img.grad[sl]))
        assert relmax(words_s.grad, words.grad[sl]) <= 2e-5, (shard, relmax(words_s.grad, words.grad[sl]))
        assert len(satt) == b
        for a, r in zip(satt, fatt[sl]):
            assert float((a - r).abs().max()) <= 1e-7, shard
        # the autograd-free step API (static buffers, direct launches; EEGAN_CHECK_SHARDED_GRAPH=1: also as one CUDA graph)
        for use_graph in ([False, True] if os.environ.get("EEGAN_CHECK_SHARDED_GRAPH") == "1" else [False]):
            sstep = ShardedWordsLossStep(b, 256, 17, 17, T, dev, w0=1.0, w1=2.0, graph=use_graph, shard=shard)
            for _ in range(2):  # second call: replay on the same buffers
                q0, q1, qdi, qdw = sstep(c["img"][sl].to(dev), c["words"][sl].to(dev), c["cap_lens"][sl].to(dev), c["class_ids"][sl].to(dev))
            assert abs(q0.item() - f0.item()) <= 2e-6 * max(1.0, abs(f0.item())) and abs(q1.item() - f1.item()) <= 2e-6 * max(1.0, abs(f1.item())), shard
            assert relmax(qdi, img.grad[sl]) <= 2e-5 and relmax(qdw, words.grad[sl]) <= 2e-5, (shard, use_graph, relmax(qdi, img.grad[sl]), relmax(qdw, words.grad[sl]))
            for i in range(b):
                Ti = int(c["cap_lens"][sl][i])
                assert float((sstep.att[i, :Ti].reshape(Ti, -1) - fatt[rank * b + i].reshape(Ti, -1)).abs().max()) <= 1e-7, (shard, "att")
            sstep.release_graph()  # a live graph holding NCCL work must go before destroy_process_group (teardown hang otherwise)
    if os.environ.get("EEGAN_CHECK_SHARDED_OVERLAP") == "1":  # the opt-in overlapped step (sharded.py)
        from eegan_b200.sharded import OverlappedShardedWordsLossStep
        ostep = OverlappedShardedWordsLossStep(b, 256, 17, 17, T, dev, w0=1.0, w1=2.0)
        for _ in range(2):
            q0, q1, qdi, qdw = ostep(c["img"][sl].to(dev), c["words"][sl].to(dev), c["cap_lens"][sl].to(dev), c["class_ids"][sl].to(dev))
        assert abs(q0.item() - f0.item()) <= 2e-6 * max(1.0, abs(f0.item())) and abs(q1.item() - f1.item()) <= 2e-6 * max(1.0, abs(f1.item()))
        assert relmax(qdi, img.grad[sl]) <= 2e-5 and relmax(qdw, words.grad[sl]) <= 2e-5, ("overlap", relmax(qdi, img.grad[sl]))
        for i in range(b):
            Ti = int(c["cap_lens"][sl][i])
            assert float((ostep.att[i, :Ti].reshape(fatt[sl][i].shape) - fatt[sl][i]).abs().max()) <= 1e-7
    # sentence loss
    sc = cases.sent_case(B, seed=12)
    cnn = sc["cnn"].to(dev).requires_grad_()
    rnn = sc["rnn"].to(dev).requires_grad_()
    g0, g1 = E.sent_loss(cnn, rnn, sc["labels"].to(dev), sc["class_ids"], B)
    (g0 + g1).backward()
    cnn_s = sc["cnn"][sl].to(dev).requires_grad_()
    rnn_s = sc["rnn"][sl].to(dev).requires_grad_()
    h0, h1 = sharded_sent_loss(cnn_s, rnn_s, torch.arange(b, device=dev), sc["class_ids"][sl], b)
    (h0 + h1).backward()
    assert abs(h0.item() - g0.item()) <= 2e-6 * max(1.0, abs(g0.item())) and abs(h1.item() - g1.item()) <= 2e-6 * max(1.0, abs(g1.item()))
    assert relmax(cnn_s.grad, cnn.grad[sl]) <= 2e-5 and relmax(rnn_s.grad, rnn.grad[sl]) <= 2e-5
    # SyncBN over NCCL vs the reference's N-replica formula on the full batch:
    # inv_std = clamp(biased var, eps) ** -0.5   (sync_batchnorm/batchnorm.py:113-125), not 1/sqrt(var + eps)
    from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
    g = torch.Generator().manual_seed(5)
    x_full = torch.randn(4 * world, 32, 16, 16, generator=g)
    go_full = torch.randn(4 * world, 32, 16, 16, generator=g)
    bn = SynchronizedBatchNorm2d(32).to(dev)
    bn.train()
    xs = x_full[4 * rank:4 * rank + 4].to(dev).requires_grad_()
    y = bn(xs)
    y.backward(go_full[4 * rank:4 * rank + 4].to(dev))
    xr = x_full.to(dev).requires_grad_()
    mean = xr.mean(dim=(0, 2, 3), keepdim=True)
    var = ((xr - mean) ** 2).mean(dim=(0, 2, 3), keepdim=True)
    yr = (xr - mean) * var.clamp(min=bn.eps) ** -0.5 * bn.weight.detach().view(1, -1, 1, 1) + bn.bias.detach().view(1, -1, 1, 1)
    yr.backward(go_full.to(dev))
    assert float((y.detach() - yr.detach()[4 * rank:4 * rank + 4]).abs().max()) <= 2e-5
    assert relmax(xs.grad, xr.grad[4 * rank:4 * rank + 4]) <= 1e-4
    # fused affine_ssa over NCCL vs the N-replica restatement on the full batch
    from oracle import damsm_oracle as O
    norm = SynchronizedBatchNorm2d(32, affine=False).to(dev)
    norm.train()
    wf = torch.randn(4 * world, 32, generator=g) * 0.5
    bf = torch.randn(4 * world, 32, generator=g) * 0.5
    mf = torch.sigmoid(torch.randn(4 * world, 1, 16, 16, generator=g))
    sl4 = slice(4 * rank, 4 * rank + 4)
    ts = [v[sl4].to(dev).requires_grad_() for v in (x_full, wf, bf, mf)]
    ys = E.ssa_modulate(ts[0], ts[1], ts[2], ts[3], norm)
    ys.backward(go_full[sl4].to(dev))
    tf = [v.to(dev).double().requires_grad_() for v in (x_full, wf, bf, mf)]
    yf = O.port_affine_ssa(tf[0], tf[1], tf[2], tf[3], eps=norm.eps, n_replica_formula=True)
    yf.backward(go_full.to(dev).double())
    assert float((ys.detach().double() - yf.detach()[sl4]).abs().max()) <= 5e-5
    for a, b_ in zip(ts, tf):
        assert relmax(a.grad.double(), b_.grad[sl4]) <= 1e-4
    dist.barrier()
    if rank == 0:
        print("multigpu_check ok: world %d, words %.6f/%.6f sent %.6f/%.6f" % (world, s0.item(), s1.item(), h0.item(), h1.item()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — EE-GAN DAMSM hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one synthetic CUB-shaped batch: words_loss
(miscc/DAMSM_losses.py:272-342) forward + backward with BOTH input gradients, i.e. the
B x B grid of 289-region x <=18-word attentions, the gamma-LSE aggregation and the two
cross-entropies.  Metric: attention pairs/s (one pair = one (caption, image) cell, fwd+bwd).

 * value      — device-resident inputs, CUDA events around each step (L2 flushed between
                steps, outside the timed spans), max over ranks.
 * e2e        — same step through the public API from pinned HOST buffers: H2D of the
                inputs and D2H of the two losses inside the timed span.
 * roofline   — the dominant kernel's algorithmic FLOP/s from stage events recorded inside
                the timed steps (eegan_profile_*), against MEASURED_PEAKS.json.
 * cpu_baseline — the oracle's loop-structured fp32 port (the reference's algorithm and op
                mix) timed on this host's cores on a bounded sample (rank 0, N=1 only).
 * --impl reference — that same CPU arm as its own JSON line (the reference is pure Python
                on torch CPU ops; /root/reference does not exist on the GPU box, so the port
                under oracle/ that was validated bit-exact against it is what runs).
N>1 (torchrun, one rank per GPU): each rank keeps B=48 captions + their images, the grid is
caption-row-sharded (eegan_b200/sharded.py): all-gather of region features, local column
block, all-gather of the blocks, redundant CE, reduce-scatter of d_img.  "weak": per-GPU
caption rows fixed.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

B_PER_GPU, T_MAX, D, HW = 48, 18, 256, 17
R = HW * HW
METRIC = "attn+DAMSM fwd/bwd pairs/s at CUB shape"
UNIT = "pairs/s"
# kernels launched per step by OUR library, default engine (3, half-pair operands):
#   pair fwd 6 (pre: scan + maxima, pack: words + image features -> fp16 hi/lo, gemm S + attention fwd, gemm U, cos/lse,
#   att_maps) + CE fwd 2 + CE bwd 1 + pair bwd 6 (dU scale, dU, gemm dA + attention bwd, gemm dC, gemm dW, unpack)
LAUNCHES_PER_STEP = {3: 15, 2: 14}
# dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, averaged over the five launches of one step, from the
# committed ncu --set full captures: engine 3 profiles/r1_h_gemm_ncu_full_summary.csv (30.6 / 53.2 / 88.6 / 107.9 /
# 50.3 MB), engine 2 profiles/r1_v3_fused_step_ncu_full_summary.csv (32.3 / 52.9 / 89.2 / 107.4 / 50.3 MB);
# algorithmic bytes of the whole step are 45.3 MB — the rest is the stash round trips
NCU_TRAFFIC_BYTES_PER_GEMM_LAUNCH = {3: 66.1e6, 2: 66.4e6}
ENGINE_NOTE = {
    3: ("h_gemm_kernel (tcgen05.mma.kind::f16 on operands stored as fp16 hi/lo pairs with power-of-two scales; TMA -> MMA, no "
        "in-kernel split; 5 launches/step: S + attention fwd epilogue, U, dA + attention bwd epilogue, dC (two accumulators), dW)",
        3.0, "fp32-accurate 3xFP16 (SURVEY D7: single-pass TF32/BF16 flips argmax words): 3 tcgen05.mma.kind::f16 per K-step at "
             "the bf16 rate, so the engine's own ceiling is peak/3"),
    2: ("ts_gemm_kernel (tcgen05 3xTF32, A operand staged in TMEM; 5 launches/step: S + attention fwd epilogue, U, dA + attention "
        "bwd epilogue, dC, dW)",
        6.0, "fp32-accurate 3xTF32 (SURVEY D7): 3 tcgen05.mma.kind::tf32 per K-step at half the bf16 rate, so the engine's own "
             "ceiling is peak/6"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.err = [], None, ""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0 and self.proc.poll() is None:
                time.sleep(0.02)  # nvidia-smi needs a few hundred ms before its first sample
        except OSError as e:
            self.proc, self.err = None, str(e)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable: " + self.err]}
        time.sleep(0.06)
        self.proc.terminate()
        if not self.rows:
            try:
                self.err = (self.proc.stderr.read() or "")[:200]
            except Exception:
                pass
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi gave no samples: " + self.err]}
        sm, mx, reasons = [], None, set()
        rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.15] or self.rows[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(B, seed):
    """Synthetic CUB-shaped batch (SURVEY.md §8d "realistic set"): img = relu(randn) * 0.3 - 0.05, words = tanh(randn) * 0.5
    (score std ~1), ragged caption lengths in [5, T_MAX] with at least one of each extreme, CUB-like class ids with
    collisions (same-class cells become -inf).  Same recipe and RNG order as the tests' seeded cases."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    img = torch.relu(torch.randn(B, D, HW, HW, generator=g)) * 0.3 - 0.05
    words = torch.tanh(torch.randn(B, D, T_MAX, generator=g)) * 0.5
    cap = torch.randint(5, T_MAX + 1, (B,), generator=g)
    cap[0] = T_MAX
    if B > 1:
        cap[1] = 5
    cls = torch.randint(1, max(2, min(201, B // 2 + 2)), (B,), generator=g)
    return dict(img=img, words=words, cap_lens=cap, labels=torch.arange(B), class_ids=cls)


def algorithmic_flops(cap_lens_sum, B_img):
    """12*R*D*sum(T_i) per image row of the grid: fwd 4RDT + bwd 8RDT."""
    return 12.0 * R * D * float(cap_lens_sum) * B_img


# ---------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------
def cpu_step(c, B):
    from oracle import damsm_oracle as O
    img = c["img"].clone().requires_grad_()
    words = c["words"].clone().requires_grad_()
    l0, l1, _ = O.port_words_loss(img, words, c["labels"], c["cap_lens"], c["class_ids"], B)
    (l0 + l1).backward()
    return float((l0 + l1).detach())


def cpu_arm(steps, warmup, budget_s=25.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = B_PER_GPU
    c = make_inputs(B, 3407)
    for _ in range(max(1, min(warmup, 2))):
        cpu_step(c, B)
    times, t_begin = [], time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(c, B)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    return dict(value=B * B / (ms / 1e3), unit=UNIT, cores=cores, kind="port",
                sample="%d full steps of the B=%d batch (fwd+bwd, both grads), oracle/damsm_oracle.py port_words_loss, "
                       "torch %s CPU, %d threads" % (len(times), B, torch.__version__, cores)), ms, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms, n = cpu_arm(args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": n, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CUB bird DAMSM words_loss fwd+bwd, B=%d, T<=%d ragged, D=%d, %dx%d regions, CPU host cores"
                                   % (B_PER_GPU, T_MAX, D, HW, HW)},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# secondary kernels of the path (reported under "extra"; N=1 only)
# ---------------------------------------------------------------------------------------
def _time_cuda(fn, iters, flush):
    ms = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / iters


def _graph_replay(fn, dev):
    """Capture fn (warmed up on a side stream) into a CUDA graph and return its replay: the side measurements time device
    work, and at these sizes an eager autograd call costs the host more than the kernels cost the GPU."""
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def gag_extra(dev, flush, hbm_gbs):
    """GlobalAttentionGeneral fwd+bwd (grads on both outputs) at the synthetic generator shapes of
    SURVEY.md §8a-a7: B=48, T=18, (64^2,128ch) (128^2,64ch) (256^2,32ch).  HBM-bound: algorithmic
    bytes per (sample, pixel) row fwd+bwd = (5*idf + 2*T)*4."""
    import eegan_b200 as E
    out = []
    g = torch.Generator(device="cpu").manual_seed(7)
    Bq, T = 48, 18
    lens = torch.randint(5, T + 1, (Bq,), generator=g)
    mask = (torch.arange(T)[None, :] >= lens[:, None]).to(dev)
    for res, idf in ((64, 128), (128, 64), (256, 32)):
        x = torch.randn(Bq, idf, res, res, device=dev).requires_grad_()
        key = (torch.randn(Bq, idf, T, device=dev) * idf ** -0.5).requires_grad_()
        val = torch.randn(Bq, idf, T, device=dev).requires_grad_()
        mod = E.GlobalAttentionGeneral(idf, 256)
        mod.applyMask(mask)
        go = torch.randn(Bq, idf, res, res, device=dev)
        ga = torch.randn(Bq, T, res, res, device=dev)

        def fwd_bwd():
            x.grad = key.grad = val.grad = None
            o, a = mod(x, key, val)
            torch.autograd.backward([o, a], [go, ga])

        def fwd_only():
            with torch.no_grad():
                mod(x, key, val)

        ms = _time_cuda(_graph_replay(fwd_bwd, dev), 10, flush)
        ms_f = _time_cuda(_graph_replay(fwd_only, dev), 10, flush)
        rows = Bq * res * res
        by = (5 * idf + 2 * T) * 4 * rows
        by_f = (2 * idf + T) * 4 * rows
        out.append({"res": res, "idf": idf, "timing": "CUDA-graph replay", "rows_per_s": rows / (ms / 1e3), "ms_fwd_bwd": ms, "ms_fwd": ms_f,
                    "hbm_gbs_fwd_bwd": by / (ms / 1e3) / 1e9, "hbm_frac_fwd_bwd": by / (ms / 1e3) / 1e9 / hbm_gbs,
                    "hbm_gbs_fwd": by_f / (ms_f / 1e3) / 1e9, "hbm_frac_fwd": by_f / (ms_f / 1e3) / 1e9 / hbm_gbs})
        del x, key, val, go, ga
    return out


def syncbn_extra(dev, flush, hbm_gbs):
    """SynchronizedBatchNorm2d fwd+bwd, single replica, at generator shapes (SURVEY App. C, B=32).
    Algorithmic bytes: fwd read x twice + write y; bwd read x, dy twice + write dx = 8 passes."""
    from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
    out = []
    for C, hw in ((256, 16), (64, 128), (32, 256)):
        x = torch.randn(32, C, hw, hw, device=dev).requires_grad_()
        bn = SynchronizedBatchNorm2d(C).to(dev)
        gy = torch.randn_like(x)

        def fwd_bwd():
            x.grad = None
            bn(x).backward(gy)

        for _ in range(3):
            fwd_bwd()
        ms = _time_cuda(fwd_bwd, 10, flush)
        by = 8 * x.numel() * 4
        out.append({"shape": [32, C, hw, hw], "ms_fwd_bwd": ms, "hbm_gbs": by / (ms / 1e3) / 1e9,
                    "hbm_frac": by / (ms / 1e3) / 1e9 / hbm_gbs})
        del x, gy
    return out


def ssa_extra(dev, flush, hbm_gbs):
    """Fused affine_ssa (models.py:43-86: SyncBN(affine=False) + mask-gated modulation) fwd+bwd, single replica,
    at generator shapes (B=32), next to the reference's own op sequence run eagerly by torch on the same GPU.
    Bytes counted: the 8 full-tensor passes the fused kernels make (fwd: stats read, apply read + write;
    bwd: reduce reads x, dy; apply reads x, dy, writes dx); the reference sequence makes ~19."""
    import eegan_b200 as E
    from eegan_b200.sync_batchnorm import SynchronizedBatchNorm2d
    out = []
    for C, hw in ((256, 16), (64, 128), (32, 256)):
        x = torch.randn(32, C, hw, hw, device=dev).requires_grad_()
        w = (torch.randn(32, C, device=dev) * 0.3).requires_grad_()
        b = (torch.randn(32, C, device=dev) * 0.3).requires_grad_()
        m = torch.sigmoid(torch.randn(32, 1, hw, hw, device=dev)).requires_grad_()
        gy = torch.randn_like(x)
        norm = SynchronizedBatchNorm2d(C, affine=False).to(dev)

        def fwd_bwd():
            x.grad = w.grad = b.grad = m.grad = None
            E.ssa_modulate(x, w, b, m, norm).backward(gy)

        def torch_seq():  # models.py:69-86 as written, with torch's own batch norm
            x.grad = w.grad = b.grad = m.grad = None
            f = torch.nn.functional.batch_norm(x, None, None, None, None, True, 0.1, 1e-5)
            ww = w.unsqueeze(-1).unsqueeze(-1).expand(f.size()) * m + 1
            bb = b.unsqueeze(-1).unsqueeze(-1).expand(f.size()) * m
            (ww * f + bb).backward(gy)

        for _ in range(3):
            fwd_bwd()
            torch_seq()
        ms = _time_cuda(fwd_bwd, 10, flush)
        ms_t = _time_cuda(torch_seq, 10, flush)
        by = 8 * x.numel() * 4
        out.append({"shape": [32, C, hw, hw], "ms_fwd_bwd": ms, "hbm_gbs": by / (ms / 1e3) / 1e9,
                    "hbm_frac": by / (ms / 1e3) / 1e9 / hbm_gbs, "ms_torch_eager_same_gpu": ms_t})
        del x, gy, w, b, m
    return out


def aux_rows_extra(dev, flush, bf16_tflops):
    """SURVEY.md 8f ranks 1-2 at CUB sizes (B=48), fwd+bwd with all gradients, next to torch's own eager ops on the
    same GPU (fp32 matmul precision 'highest', i.e. the reference's arithmetic).
    emb_features: conv1x1 768 -> 256 on the 17x17 map (DAMSM.py:162, 229); algorithmic FLOP = 3 GEMMs of 2*B*R*Cin*Cout.
    ATTR_Enhance: 4 tokens x 256 channels (models.py:146-169): launch-latency-bound, reported in microseconds."""
    import eegan_b200 as E
    out = {}
    B, Cin, Cout, H = 48, 768, 256, 17
    x = torch.relu(torch.randn(B, Cin, H, H, device=dev)).requires_grad_()
    mod = E.EmbFeatures(Cin, Cout).to(dev)
    conv = torch.nn.Conv2d(Cin, Cout, 1, bias=False).to(dev)
    gy = torch.randn(B, Cout, H, H, device=dev)

    def ours():
        x.grad = None
        mod.weight.grad = None
        mod(x).backward(gy)

    def ref():
        x.grad = None
        conv.weight.grad = None
        conv(x).backward(gy)

    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(3):
        ours()
        ref()
    ms, ms_t = _time_cuda(ours, 10, flush), _time_cuda(ref, 10, flush)
    flop = 3 * 2.0 * B * H * H * Cin * Cout
    out["emb_features_conv1x1"] = {"shape": [B, Cin, H, H], "cout": Cout, "ms_fwd_bwd": ms, "tflops": flop / (ms / 1e3) / 1e12,
                                   "frac_of_3xtf32_ceiling": flop / (ms / 1e3) / 1e12 / (bf16_tflops / 6.0),
                                   "ms_torch_eager_fp32_same_gpu": ms_t}
    D, A = 256, 3
    ae = E.ATTR_Enhance(ntf=D).to(dev)
    sent = torch.randn(B, D, device=dev).requires_grad_()
    attrs = torch.randn(B, A, D, device=dev).requires_grad_()
    gs, ga = torch.randn(B, D, device=dev), torch.randn(B, A + 1, D, device=dev)
    import torch.nn.functional as F

    def ours2():
        sent.grad = attrs.grad = None
        ae.zero_grad(set_to_none=True)
        a, b = ae(sent, attrs)
        ((a * gs).sum() + (b * ga).sum()).backward()

    def ref2():
        sent.grad = attrs.grad = None
        ae.zero_grad(set_to_none=True)
        # models.py:161-168 op for op (cat, three Linear, softmax(q k^T) * norm, bmm), run by torch on the same GPU
        combine = torch.cat([sent.unsqueeze(1), attrs], dim=1)
        q, k, v = F.linear(combine, ae.attr_query.weight, ae.attr_query.bias), F.linear(combine, ae.attr_key.weight, ae.attr_key.bias), \
            F.linear(combine, ae.attr_value.weight, ae.attr_value.bias)
        b = torch.bmm(torch.softmax(torch.bmm(q, k.permute(0, 2, 1)), dim=-1) * ae._norm_fact, v)
        a = b[:, 0, :]
        ((a * gs).sum() + (b * ga).sum()).backward()

    # a few microseconds of device work behind ~30 Python-level ops: time CUDA-graph replays of both, so that the number
    # is the device time of the launches and not the interpreter
    r_ours, r_ref = _graph_replay(ours2, dev), _graph_replay(ref2, dev)
    out["attr_enhance"] = {"shape": [B, A + 1, D], "us_fwd_bwd": 1e3 * _time_cuda(r_ours, 20, flush),
                           "us_torch_ops_same_gpu": 1e3 * _time_cuda(r_ref, 20, flush),
                           "timing": "CUDA-graph replay of forward + backward (device time of the launches)"}
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    return out


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    import eegan_b200 as E
    from eegan_b200 import _lib
    from eegan_b200.sharded import sharded_words_loss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    B = B_PER_GPU
    Btot = B * world
    c = make_inputs(B, 3407 + rank)
    img_h = c["img"].pin_memory()
    words_h = c["words"].pin_memory()
    lens_h = c["cap_lens"].pin_memory()
    cls = c["class_ids"]  # CPU LongTensor as in train.py:423
    labels = c["labels"].to(dev)
    img_d = img_h.to(dev).requires_grad_()
    words_d = words_h.to(dev).requires_grad_()
    lens_d = lens_h.to(dev)
    lens_sum = torch.tensor([float(lens_h.sum())], device=dev)
    if world > 1:
        dist.all_reduce(lens_sum)
    # per-rank algorithmic work: its column block = B_tot images x local captions
    flops_rank = algorithmic_flops(float(lens_h.sum()), Btot)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def step(img, words, lens):
        img.grad = None
        words.grad = None
        if world > 1:
            l0, l1, _ = sharded_words_loss(img, words, labels, lens, cls, B)
        else:
            l0, l1, _ = E.words_loss(img, words, labels, lens, cls, B)
        (l0 + l1).backward()
        return l0, l1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N>1: capturing the autograd route with its NCCL collectives into a graph hung on this stack (torch 2.11 / NCCL 2.28.9;
    # the backward runs on autograd's thread), so N>1 uses the autograd-free ShardedWordsLossStep below
    engine = int(L.eegan_get_contraction_engine())
    use_graph = (world == 1) and not args.eager
    graphed = None
    sstep = None
    if use_graph:
        from eegan_b200.graphed import GraphedWordsLoss
        graphed = GraphedWordsLoss(B, D, HW, HW, T_MAX, dev, use_class_ids=True, words_grad=True, sharded=world > 1)
        cls_d = cls.to(dev)
        graphed(img_d.detach(), words_d.detach(), lens_d, cls_d)  # capture

        def run_step():
            graphed.graph.replay()
    elif world > 1 and not args.autograd:
        # N > 1: the autograd-free sharded step (same collectives and kernels, enqueued directly on static buffers);
        # --sharded-graph additionally captures it, collectives included, into one CUDA graph
        from eegan_b200.sharded import OverlappedShardedWordsLossStep, ShardedWordsLossStep
        if args.sharded_overlap:  # opt-in: local image block first, collectives hidden behind it (not yet measured on GPUs)
            sstep = OverlappedShardedWordsLossStep(B, D, HW, HW, T_MAX, dev, use_class_ids=True, words_grad=True)
        else:
            sstep = ShardedWordsLossStep(B, D, HW, HW, T_MAX, dev, use_class_ids=True, words_grad=True, graph=args.sharded_graph)
        cls_d = cls.to(dev)
        sstep.load(img_d.detach(), words_d.detach(), lens_d, cls_d)

        def run_step():
            sstep.run()
    else:
        def run_step():
            step(img_d, words_d, lens_d)

    # ---- device-resident timing ---------------------------------------------------
    for _ in range(args.warmup):
        run_step()
        flush.fill_(1.0)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t_wall0 = time.time()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1.0)  # evict L2 (126 MB) between timed steps; outside the timed span
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall1 = time.time()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    pairs_per_step = float(Btot) * Btot
    value = pairs_per_step / (ms_step / 1e3)

    # ---- per-stage device times (eager launches of the same kernels, stage events) ----
    stage = None
    eager_ms = None
    if world == 1:
        L.eegan_profile_enable(1)
        pe = []
        for _ in range(args.steps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(img_d, words_d, lens_d)
            e1.record()
            pe.append((e0, e1))
        torch.cuda.synchronize()
        eager_ms = sum(a.elapsed_time(b) for a, b in pe) / args.steps
        n = L.eegan_profile_nstages()
        ms_arr, cnt_arr = (ctypes.c_double * n)(), (ctypes.c_int * n)()
        _lib.check(L.eegan_profile_collect(ms_arr, cnt_arr), "profile_collect")
        L.eegan_profile_enable(0)
        stage = [(L.eegan_profile_stage_name(i).decode(), ms_arr[i], cnt_arr[i]) for i in range(n)]

    # ---- end-to-end from pinned host buffers ---------------------------------------
    # Every step: H2D of that step's inputs (pinned -> device) and D2H of its two losses, through
    # the public API.  With the graphed API the H2D of step k+1 is prefetched on a copy stream
    # while step k computes (double-buffered staging), as an input pipeline would do.
    h2d = img_h.numel() * 4 + words_h.numel() * 4 + lens_h.numel() * 8
    loss_h = torch.empty(2, dtype=torch.float32).pin_memory()
    barrier()
    nrun = args.warmup + args.steps
    if use_graph:
        copy_stream = torch.cuda.Stream(device=dev)
        comp = torch.cuda.current_stream()
        stg = [dict(img=torch.empty_like(img_d), words=torch.empty_like(words_d), lens=torch.empty_like(lens_d),
                    ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        for sbuf in stg:
            sbuf["free"].record(comp)

        def prefetch(k):
            sbuf = stg[k % 2]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(sbuf["free"])
                sbuf["img"].copy_(img_h, non_blocking=True)
                sbuf["words"].copy_(words_h, non_blocking=True)
                sbuf["lens"].copy_(lens_h, non_blocking=True)
                sbuf["ready"].record(copy_stream)

        prefetch(0)
        t_start = None
        for k in range(nrun):
            if k == args.warmup:
                torch.cuda.synchronize()
                t_start = torch.cuda.Event(enable_timing=True)
                t_start.record()
                prefetch(k)  # the first timed step pays its own copy in full
            if k + 1 < nrun and k + 1 != args.warmup:
                prefetch(k + 1)
            sbuf = stg[k % 2]
            comp.wait_event(sbuf["ready"])
            l0, l1, _, _ = graphed(sbuf["img"], sbuf["words"], sbuf["lens"], cls_d)
            sbuf["free"].record(comp)
            loss_h.copy_(torch.stack([l0, l1]), non_blocking=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_end.record()
        t_end.synchronize()
        e2e_total = t_start.elapsed_time(t_end)
        e2e_mode = "graphed API; H2D of step k+1 prefetched on a copy stream (double-buffered) while step k computes"
    else:
        e2e_evs = []
        for k in range(nrun):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if sstep is not None:  # H2D straight into the step's static input buffers
                l0, l1, _, _ = sstep(img_h, words_h, lens_h, cls_d)
            else:
                img = img_h.to(dev, non_blocking=True).requires_grad_()
                words = words_h.to(dev, non_blocking=True).requires_grad_()
                lens = lens_h.to(dev, non_blocking=True)
                l0, l1 = step(img, words, lens)
            loss_h.copy_(torch.stack([l0.detach(), l1.detach()]), non_blocking=True)
            e1.record()
            e1.synchronize()  # the caller sees the loss on the host
            if k >= args.warmup:
                e2e_evs.append(e0.elapsed_time(e1))
        e2e_total = sum(e2e_evs)
        e2e_mode = ("sharded step API (ShardedWordsLossStep); H2D, compute and D2H serial in every step" if sstep is not None
                    else "eager API; H2D, compute and D2H serial in every step")
    barrier()
    clocks = sampler.summary(t_wall0, time.time()) if sampler else None
    t = torch.tensor([e2e_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    e2e = {"value": pairs_per_step / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": 8 * world, "ms_per_step": e2e_ms, "mode": e2e_mode}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "CUB bird (cfg/bird.yml shape) DAMSM words_loss fwd+bwd, both grads: B=%d per GPU "
                                   "(global %d), T<=%d ragged (sum=%d), D=%d, %dx%d regions; caption-row-sharded for N>1"
                                   % (B, Btot, T_MAX, int(lens_sum.item()), D, HW, HW),
                       "l2": "256 MB fill between timed steps (outside the timed spans)",
                       "pairs_per_step": pairs_per_step,
                       "launch": ("one CUDA graph per step (eegan_b200.graphed.GraphedWordsLoss)" if use_graph else
                                  ("sharded step, %s (eegan_b200.sharded.ShardedWordsLossStep)" % ("one CUDA graph" if args.sharded_graph else "direct launches")
                                   if sstep is not None else "eager autograd launches")),
                       "eager_ms_per_step": eager_ms},
            "e2e": e2e, "gpu_launches": LAUNCHES_PER_STEP.get(engine, 15) * args.steps, "clocks": clocks}
    line["config"]["contraction_engine"] = engine
    if stage is not None:
        gemm = [s for s in stage if s[0].startswith("gemm")]
        gemm_ms = sum(s[1] for s in gemm)
        gemm_launch_count = args.steps * 5  # S, U, dA, dC, dW
        achieved = flops_rank * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        kname, div, why = ENGINE_NOTE.get(engine, ENGINE_NOTE[3])
        ceiling = pk["bf16_tflops"] / div
        line["roofline"] = {
            "bound": "tensor", "kernel": kname,
            "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops"] if achieved else None, "traffic": NCU_TRAFFIC_BYTES_PER_GEMM_LAUNCH.get(engine),
            "peak_source": pk["source"] + " cuBLAS bf16 burst (MEASURED_PEAKS.json)",
            "engine_ceiling_tflops": ceiling, "frac_of_engine_ceiling": (achieved / ceiling) if achieved else None,
            "note": "%s = %.0f TFLOP/s; avg launch %.1f us; achieved = algorithmic 12*R*D*sum(T)*B FLOP per step / summed "
                    "duration of the 5 GEMM launches (their epilogues carry the softmax work of the path), measured with stage "
                    "events on eager launches of the same kernels (the headline value replays them as one CUDA graph with "
                    "programmatic dependent launch between them)"
                    % (why, ceiling, 1e3 * gemm_ms / max(1, gemm_launch_count)),
            "stage_ms_per_step": {s[0]: s[1] / args.steps for s in stage},
            "hbm_equiv": {"algorithmic_bytes_per_step": 3 * B * D * (R + T_MAX) * 4 + 3 * B * B * 4,
                          "achieved_gbs": (3 * B * D * (R + T_MAX) * 4 + 3 * B * B * 4) / (ms_step / 1e3) / 1e9,
                          "peak_gbs": pk["hbm_gbs"]}}
    if world == 1:
        if not args.no_extra:
            line["extra"] = {"global_attention_general": gag_extra(dev, flush, pk["hbm_gbs"]),
                             "sync_batchnorm_1replica": syncbn_extra(dev, flush, pk["hbm_gbs"]),
                             "affine_ssa_1replica": ssa_extra(dev, flush, pk["hbm_gbs"]),
                             "aux_rows_8f": aux_rows_extra(dev, flush, pk["bf16_tflops"])}
        base, _, _ = cpu_arm(args.steps, 1)
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the GlobalAttentionGeneral / SyncBN side measurements")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of the CUDA-graph replay")
    ap.add_argument("--autograd", action="store_true", help="N>1: time the autograd route (sharded_words_loss + backward)")
    ap.add_argument("--sharded-graph", action="store_true", help="N>1: capture the sharded step, collectives included, into a CUDA graph")
    ap.add_argument("--sharded-overlap", action="store_true",
                    help="N>1: OverlappedShardedWordsLossStep (all-gather / reduce-scatter overlapped with the local image block)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

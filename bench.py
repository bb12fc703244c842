#!/usr/bin/env python
"""bench.py — EE-GAN DAMSM hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]

A "step" is one pass of the hot path over one synthetic batch: words_loss (miscc/DAMSM_losses.py:272-342) forward
+ backward with BOTH input gradients, i.e. the B x B grid of 289-region x <=T-word attentions, the gamma-LSE
aggregation and the two cross-entropies.  Metric: attention pairs/s (one pair = one (caption, image) cell, fwd+bwd).

Workloads (BASELINE.json `configs`):
  c2 (default; the metric's config)  CUB cfg/bird.yml shape: B = 48 per GPU, T <= 18 ragged, CUB-like class ids.  N > 1: the
      grid is sharded over the ranks (eegan_b200/sharded.py; --shard images (default): own images x all captions, the word
      features travel; --shard captions: own captions x all images, the region features travel), per-GPU samples fixed ->
      pairs = (48 N)^2, "weak".
  c3  COCO cfg/coco.yml shape: GLOBAL B = 64, T <= 20, unique class ids, split over the N GPUs ("strong").
  c5  flower cfg/flower.yml shape: GLOBAL B in {32, 64, 128, 256, 512} split over the N GPUs, one sub-line per B.
  c4  the generator step: the reference's own models.Gen (24 SyncBN layers) + Trainer.DAMSM_loss, B = 32 per GPU, imported
      from the staged reference (baseline/_ref) with eegan_b200.install(); next to the reference's own modules on the same GPU.

Every line carries:
  value      device-resident inputs, CUDA events around each step (L2 flushed between steps, outside the timed spans), max over ranks
  e2e        the same step through the reference-facing API — `words_loss(...)` + `.backward()` — from pinned HOST buffers: H2D of
             the inputs and D2H of the two losses inside the timed span (gradients stay on the device: they feed the encoder's backward)
  parity     BEFORE timing: the sharded step / drop-in API, sent_loss and one SyncBN layer against the float64 oracle (oracle/parity.py)
  roofline   the dominant kernels' algorithmic FLOP/s from stage events recorded inside warmed steps, against MEASURED_PEAKS.json
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, kind "reference"; the validated port when it is not staged) on this host's cores
  N > 1 also: comm_free_same_shape (the rank's block of the grid timed without any collective) and nccl_ms_per_step.
`--impl reference` prints the CPU arm as its own line (rank 0 only).
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
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

D, HW = 256, 17
R = HW * HW
METRIC = "attn+DAMSM fwd/bwd pairs/s at CUB shape"
UNIT = "pairs/s"
WORKLOADS = {
    "c2": dict(name="CUB bird (cfg/bird.yml shape)", per_gpu=48, T=18, cls="cub", scaling="weak"),
    "c3": dict(name="COCO (cfg/coco.yml shape)", global_B=64, T=20, cls="unique", scaling="strong"),
    "c5": dict(name="Oxford flower (cfg/flower.yml shape) pair-grid sweep", sweep=(32, 64, 128, 256, 512), T=20, cls="cub", scaling="strong"),
}
# kernels launched per step by OUR library, default engine (3, half-pair operands):
#   pair fwd 6 (pre: scan + maxima, pack: words + image features -> fp16 hi/lo, gemm S + attention fwd, gemm U, cos/lse,
#   att_maps) + CE fwd 2 + CE bwd 1 + pair bwd 6 (dU scale, dU, gemm dA + attention bwd, gemm dC, gemm dW, unpack)
LAUNCHES_PER_STEP = {3: 15, 2: 14}
# dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, averaged over the five launches of one step at B = 48.  NOT measured by
# this run: a constant copied from the committed ncu --set full capture named in `traffic_source`.
NCU_TRAFFIC = {3: (66.4e6, "profiles/r2_h_gemm_ncu_full_summary.csv (31.1 / 52.8 / 88.7 / 107.3 / 50.4 MB for the five GEMM launches, B=48; round 1: 30.6 / 53.2 / 88.6 / 107.9 / 50.3)"),
               2: (66.4e6, "profiles/r1_v3_fused_step_ncu_full_summary.csv")}
ENGINE_NOTE = {
    3: ("h_gemm_kernel (tcgen05.mma.kind::f16 on operands stored as fp16 hi/lo pairs with power-of-two scales; TMA -> MMA, no "
        "in-kernel split; 5 launches/step: S + attention fwd epilogue, U, dA + attention bwd epilogue, dC (two accumulators), dW)",
        3.0, "fp32-accurate 3xFP16 (SURVEY D7: single-pass TF32/BF16 flips argmax words): 3 tcgen05.mma.kind::f16 per K-step at "
             "the bf16 rate, so the engine's own ceiling is peak/3"),
    2: ("ts_gemm_kernel (tcgen05 3xTF32, A operand staged in TMEM; 5 launches/step)", 6.0,
        "fp32-accurate 3xTF32 (SURVEY D7): 3 tcgen05.mma.kind::tf32 per K-step at half the bf16 rate, so the engine's own ceiling is peak/6"),
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


def make_inputs(B, T, seed, cls_mode="cub"):
    """Synthetic batch (SURVEY.md §8d "realistic set"): img = relu(randn) * 0.3 - 0.05, words = tanh(randn) * 0.5
    (score std ~1), ragged caption lengths in [5, T] with at least one of each extreme; class ids CUB-like with
    collisions (same-class cells become -inf) or unique (COCO).  Same recipe and RNG order as the tests' seeded cases."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    img = torch.relu(torch.randn(B, D, HW, HW, generator=g)) * 0.3 - 0.05
    words = torch.tanh(torch.randn(B, D, T, generator=g)) * 0.5
    cap = torch.randint(5, T + 1, (B,), generator=g)
    cap[0] = T
    if B > 1:
        cap[1] = 5
    cls = torch.randint(1, max(2, min(201, B // 2 + 2)), (B,), generator=g)
    if cls_mode == "unique":
        cls = torch.arange(B)
    return dict(img=img, words=words, cap_lens=cap, labels=torch.arange(B), class_ids=cls)


def algorithmic_flops(cap_lens_sum, B_img):
    """12*R*D*sum(T_i) per image row of the grid: fwd 4RDT + bwd 8RDT."""
    return 12.0 * R * D * float(cap_lens_sum) * B_img


def workload_string(cfg_name, B_total, T):
    """Identical in both arms (ours / --impl reference): the batch is the GLOBAL batch, whoever computes it."""
    w = WORKLOADS[cfg_name]
    return ("%s DAMSM words_loss fwd+bwd, both grads: global B=%d, T<=%d ragged, D=%d, %dx%d regions"
            % (w["name"], B_total, T, D, HW, HW))


# ---------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores
# ---------------------------------------------------------------------------------------
def reference_words_loss_fn():
    """(callable, kind): the UNMODIFIED reference's words_loss when it is staged under baseline/_ref (or mounted), else the port."""
    from oracle import ref_loader as RL
    if RL.reference_available():
        ref = RL.load_reference()
        return ref.losses.words_loss, "reference", ref
    from oracle import damsm_oracle as O
    return O.port_words_loss, "port", None


def cpu_step(fn, c, B):
    img = c["img"].clone().requires_grad_()
    words = c["words"].clone().requires_grad_()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        l0, l1, _ = fn(img, words, c["labels"], c["cap_lens"], c["class_ids"], B)
    (l0 + l1).backward()
    return float((l0 + l1).detach())


def cpu_arm(cfg_name, steps, warmup, budget_s=25.0, B=None):
    w = WORKLOADS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if B is None:
        B = w.get("per_gpu") or w.get("global_B") or 128  # c5: the largest batch of the sweep the CPU finishes in seconds
    fn, kind, ref = reference_words_loss_fn()
    if ref is not None:
        ref.cfg.CUDA = False
    c = make_inputs(B, w["T"], 3407, w["cls"])
    for _ in range(max(1, min(warmup, 2))):
        cpu_step(fn, c, B)
    times, t_begin = [], time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(fn, c, B)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    src = ("miscc/DAMSM_losses.py words_loss of the unmodified reference (baseline/_ref)" if kind == "reference"
           else "oracle/damsm_oracle.py port_words_loss (reference not staged)")
    return dict(value=B * B / (ms / 1e3), unit=UNIT, cores=cores, kind=kind,
                sample="%d full steps of the B=%d batch (fwd+bwd, both grads), %s, torch %s CPU, %d threads"
                       % (len(times), B, src, torch.__version__, cores)), ms, len(times), B, int(c["cap_lens"].sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "c4":
        return run_full_step_reference_cpu(args)
    w = WORKLOADS[args.config]
    # the same GLOBAL batch our arm computes at --gpus N: the reference evaluates it in one process (train.py:195, 419-435)
    Bg = w["per_gpu"] * max(1, args.gpus) if "per_gpu" in w else (w.get("global_B") or 128)
    base, ms, n, B, lens_sum = cpu_arm(args.config, args.steps, args.warmup if Bg <= 96 else min(args.warmup, 1), budget_s=120.0, B=Bg)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": n, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.config, B, w["T"]), "config_id": args.config,
                       "arm": "reference CPU implementation on the host cores: one process, the whole B=%d batch" % B},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# small timing helpers
# ---------------------------------------------------------------------------------------
class _L2Flush:
    """L2 flush between timed spans: a 256 MB WRITE (larger than the 126 MB L2), then a 256 MB READ sweep of another buffer.
    The write alone leaves the cache full of DIRTY lines, whose write-back (up to 126 MB = ~19 us of HBM time) then lands
    inside the next timed span — measured on GlobalAttentionGeneral at 64^2 x 128: 118 us after the write alone, 110 us after
    write + read sweep (profiles/README.md).  After the sweep the cache is cold and clean: the span times the kernels only."""

    def __init__(self, dev):
        n = 256 * 1024 * 1024 // 4
        self.w = torch.empty(n, dtype=torch.float32, device=dev)
        self.r = torch.zeros(n, dtype=torch.float32, device=dev)

    def fill_(self, v):
        self.w.fill_(v)
        self.r.sum()
        return self


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


# ---------------------------------------------------------------------------------------
# secondary kernels of the path (reported under "extra"; N=1 only)
# ---------------------------------------------------------------------------------------
def gag_extra(dev, flush, hbm_gbs, ref):
    """GlobalAttentionGeneral fwd+bwd (grads on both outputs) at the synthetic generator shapes of
    SURVEY.md §8a-a7: B=48, T=18, (64^2,128ch) (128^2,64ch) (256^2,32ch).  HBM-bound: algorithmic
    bytes per (sample, pixel) row fwd+bwd = (5*idf + 2*T)*4.  Next to it: the reference module run eagerly on the same GPU."""
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

        def fwd_bwd(m=mod):
            x.grad = key.grad = val.grad = None
            o, a = m(x, key, val)
            torch.autograd.backward([o, a], [go, ga])

        def fwd_only():
            with torch.no_grad():
                mod(x, key, val)

        ms = _time_cuda(_graph_replay(fwd_bwd, dev), 10, flush)
        ms_f = _time_cuda(_graph_replay(fwd_only, dev), 10, flush)
        rows = Bq * res * res
        by = (5 * idf + 2 * T) * 4 * rows
        by_f = (2 * idf + T) * 4 * rows
        row = {"res": res, "idf": idf, "timing": "CUDA-graph replay", "rows_per_s": rows / (ms / 1e3), "ms_fwd_bwd": ms, "ms_fwd": ms_f,
               "hbm_gbs_fwd_bwd": by / (ms / 1e3) / 1e9, "hbm_frac_fwd_bwd": by / (ms / 1e3) / 1e9 / hbm_gbs,
               "hbm_gbs_fwd": by_f / (ms_f / 1e3) / 1e9, "hbm_frac_fwd": by_f / (ms_f / 1e3) / 1e9 / hbm_gbs}
        if ref is not None:
            rmod = ref.losses.GlobalAttentionGeneral(idf, 256)
            rmod.applyMask(mask)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for _ in range(2):
                    fwd_bwd(rmod)
                row["ms_reference_eager_same_gpu"] = _time_cuda(lambda: fwd_bwd(rmod), 5, flush)
        out.append(row)
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
        tbn = torch.nn.BatchNorm2d(C).to(dev)
        gy = torch.randn_like(x)

        def fwd_bwd(m=bn):
            x.grad = None
            m(x).backward(gy)

        for _ in range(3):
            fwd_bwd()
            fwd_bwd(tbn)
        ms = _time_cuda(fwd_bwd, 10, flush)
        ms_t = _time_cuda(lambda: fwd_bwd(tbn), 10, flush)
        by = 8 * x.numel() * 4
        out.append({"shape": [32, C, hw, hw], "ms_fwd_bwd": ms, "hbm_gbs": by / (ms / 1e3) / 1e9,
                    "hbm_frac": by / (ms / 1e3) / 1e9 / hbm_gbs, "ms_torch_batchnorm_eager_same_gpu": ms_t})
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
    import torch.nn.functional as F
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
    Dm, A = 256, 3
    ae = E.ATTR_Enhance(ntf=Dm).to(dev)
    sent = torch.randn(B, Dm, device=dev).requires_grad_()
    attrs = torch.randn(B, A, Dm, device=dev).requires_grad_()
    gs, ga = torch.randn(B, Dm, device=dev), torch.randn(B, A + 1, Dm, device=dev)

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

    r_ours, r_ref = _graph_replay(ours2, dev), _graph_replay(ref2, dev)
    out["attr_enhance"] = {"shape": [B, A + 1, Dm], "us_fwd_bwd": 1e3 * _time_cuda(r_ours, 20, flush),
                           "us_torch_ops_same_gpu": 1e3 * _time_cuda(r_ref, 20, flush),
                           "timing": "CUDA-graph replay of forward + backward (device time of the launches)"}
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def reference_eager_same_gpu(dev, ref, c, B, flush):
    """The like-for-like GPU baseline: the UNMODIFIED reference words_loss + backward run eagerly on this
    B200 with cfg.CUDA = True — a Python loop over captions of ~75 small kernels each."""
    img = c["img"].to(dev).requires_grad_()
    words = c["words"].to(dev).requires_grad_()
    labels, lens = c["labels"].to(dev), c["cap_lens"].to(dev)
    prev = ref.cfg.CUDA
    ref.cfg.CUDA = True

    def step():
        img.grad = words.grad = None
        l0, l1, _ = ref.losses.words_loss(img, words, labels, lens, c["class_ids"], B)
        (l0 + l1).backward()

    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            step()
            ms = _time_cuda(step, 3, flush)
    finally:
        ref.cfg.CUDA = prev
    return {"what": "miscc/DAMSM_losses.py words_loss + backward of the unmodified reference, eager on this GPU (cfg.CUDA=True)",
            "ms_per_step": ms, "pairs_per_s": B * B / (ms / 1e3), "B": B}


# ---------------------------------------------------------------------------------------
# one pair-grid workload on this rank's GPU (all ranks call it together)
# ---------------------------------------------------------------------------------------
class Ctx:
    pass


def time_pair_workload(args, cx, cfg_name, B, T, cls_mode, full=True):
    """B = the rank's local batch.  Returns the dict of measured numbers for this (workload, batch)."""
    import eegan_b200 as E
    from eegan_b200 import _lib
    world, rank, dev, flush, L = cx.world, cx.rank, cx.dev, cx.flush, cx.L
    Btot = B * world
    c = make_inputs(B, T, 3407 + rank, cls_mode)
    if cls_mode == "unique":
        c["class_ids"] = c["class_ids"] + rank * B
    img_h, words_h, lens_h = c["img"].pin_memory(), c["words"].pin_memory(), c["cap_lens"].pin_memory()
    cls = c["class_ids"]  # CPU LongTensor as in train.py:423
    labels = c["labels"].to(dev)
    img_d, words_d, lens_d = img_h.to(dev).requires_grad_(), words_h.to(dev).requires_grad_(), lens_h.to(dev)
    cls_d = cls.to(dev)
    lens_sum = torch.tensor([float(lens_h.sum())], device=dev)
    if world > 1:
        dist.all_reduce(lens_sum)
    flops_rank = algorithmic_flops(float(lens_h.sum()), Btot)  # its column block = B_tot images x local captions
    engine = int(L.eegan_get_contraction_engine())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def api_step(img, words, lens):
        """the reference's own call (train.py:428 + backward)"""
        img.grad = None
        words.grad = None
        l0, l1, _ = E.words_loss(img, words, labels, lens, cls, B)
        loss = l0 + l1
        loss.backward()
        return loss

    graphed = sstep = None
    if world == 1:
        from eegan_b200.graphed import GraphedWordsLoss
        graphed = GraphedWordsLoss(B, D, HW, HW, T, dev, use_class_ids=True, words_grad=True)
        graphed(img_d.detach(), words_d.detach(), lens_d, cls_d)  # capture
        run_step = graphed.graph.replay
        shard_by = None
        launch = "one CUDA graph per step (eegan_b200.graphed.GraphedWordsLoss)"
    else:
        from eegan_b200.sharded import OverlappedShardedWordsLossStep, ShardedWordsLossStep
        if args.sharded_mode == "overlap":
            sstep = OverlappedShardedWordsLossStep(B, D, HW, HW, T, dev, use_class_ids=True, words_grad=True)
            shard_by = "captions"
        else:
            sstep = ShardedWordsLossStep(B, D, HW, HW, T, dev, use_class_ids=True, words_grad=True, graph=args.sharded_mode == "graph",
                                         shard=args.shard)
            shard_by = sstep.shard
        sstep.load(img_d.detach(), words_d.detach(), lens_d, cls_d)
        run_step = sstep.run
        launch = "sharded step, mode=%s, grid partitioned by %s (eegan_b200.sharded.%s)" % (args.sharded_mode, shard_by, type(sstep).__name__)

    # ---- device-resident timing ---------------------------------------------------
    for _ in range(args.warmup):
        run_step()
        flush.fill_(1.0)
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
    ms_step = float(t.item()) / args.steps
    pairs_per_step = float(Btot) * Btot
    res = dict(shard_by=shard_by, B_local=B, B_total=Btot, T=T, lens_sum=int(lens_sum.item()), ms_per_step=ms_step, value=pairs_per_step / (ms_step / 1e3),
               pairs_per_step=pairs_per_step, launch=launch, engine=engine, wall=(t_wall0, t_wall1))

    # ---- the drop-in API, device-resident (words_loss + backward through the reference's signature) ----
    if world == 1:
        for _ in range(max(3, args.warmup)):  # call 1 plain, call 2 captures the plan's graphs, then replays
            api_step(img_d, words_d, lens_d)
        torch.cuda.synchronize()
        n_api = max(10, args.steps)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n_api):
            api_step(img_d, words_d, lens_d)
        ev1.record()
        ev1.synchronize()
        res["api_ms_per_step"] = ev0.elapsed_time(ev1) / n_api
        res["api_note"] = ("eegan_b200.words_loss(...) + (loss0 + loss1).backward() — the reference's signature (train.py:428), planned "
                           "route (fastpath.py) — %d back-to-back calls, no L2 flush, device-resident inputs" % n_api)

    # ---- per-stage device times (warmed eager launches of the same kernels, stage events) ----
    if world == 1 and full:
        from eegan_b200 import damsm_losses as dl

        def plain_step():  # the un-planned route: direct launches, so that the library's stage events see single kernels
            img_d.grad = words_d.grad = None
            m, _ = dl.pair_grid(img_d, words_d, lens_d)
            l0, l1, _ = dl._PairCEFn.apply(m, 10.0, cls_d, labels)
            (l0 + l1).backward()

        for _ in range(3):
            plain_step()
        torch.cuda.synchronize()
        L.eegan_profile_enable(1)
        pe = []
        for _ in range(args.steps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plain_step()
            e1.record()
            pe.append((e0, e1))
        torch.cuda.synchronize()
        res["eager_ms_per_step"] = sum(a.elapsed_time(b) for a, b in pe) / args.steps
        n = L.eegan_profile_nstages()
        ms_arr, cnt_arr = (ctypes.c_double * n)(), (ctypes.c_int * n)()
        _lib.check(L.eegan_profile_collect(ms_arr, cnt_arr), "profile_collect")
        L.eegan_profile_enable(0)
        res["stage"] = [(L.eegan_profile_stage_name(i).decode(), ms_arr[i], cnt_arr[i]) for i in range(n)]
        res["flops_rank"] = flops_rank

    # ---- N > 1: the same-shape block without any collective, and the collectives alone ----
    if world > 1:
        p = _lib.ptr
        by_img = res.get("shard_by") == "images"
        Bi_blk, Bc_blk = (B, Btot) if by_img else (Btot, B)  # the rank's block: own images x all captions, or all images x own captions
        ws = torch.empty(L.eegan_damsm_pair_workspace_bytes(Bi_blk, Bc_blk, D, R, T), dtype=torch.uint8, device=dev)
        img_all = img_d.detach().reshape(B, D, R).repeat(1 if by_img else world, 1, 1).contiguous()
        wd = words_d.detach().repeat(world if by_img else 1, 1, 1).contiguous()
        lens32 = lens_d.to(torch.int32).repeat(world if by_img else 1).contiguous()
        m_blk, dm_blk = torch.empty(Bi_blk, Bc_blk, device=dev), torch.full((Bi_blk, Bc_blk), 1e-3, device=dev)
        att = torch.empty(Bc_blk, T, R, device=dev)
        d_img_all, d_words = torch.empty(Bi_blk, D, R, device=dev), torch.empty(Bc_blk, D, T, device=dev)

        def comm_free():
            st = _lib.stream_ptr()
            _lib.check(L.eegan_damsm_pair_fwd(p(img_all), p(wd), p(lens32), Bi_blk, Bc_blk, D, R, T, 5.0, 5.0, p(m_blk), p(att),
                                              -rank * B if by_img else rank * B, p(ws), ws.numel(), st), "pair_fwd")
            _lib.check(L.eegan_damsm_pair_bwd(p(img_all), p(wd), p(lens32), Bi_blk, Bc_blk, D, R, T, 5.0, 5.0, p(dm_blk), p(d_img_all),
                                              p(d_words), p(ws), ws.numel(), st), "pair_bwd")

        cf = _graph_replay(comm_free, dev)
        barrier()
        t = torch.tensor([_time_cuda(cf, max(5, args.steps // 2), flush)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cf_ms = float(t.item())
        res["comm_free_same_shape"] = {
            "what": "pair grid fwd + bwd of this rank's (%d images x %d captions) block, no collective, no CE (O(B^2), ~10 us), one CUDA graph, "
                    "max over ranks" % (Bi_blk, Bc_blk),
            "ms_per_step": cf_ms, "per_gpu_pairs_per_s": Btot * B / (cf_ms / 1e3),
            "sharded_step_over_comm_free": ms_step / cf_ms}
        del ws
        # the collectives of one step, alone, back to back
        g_cls = torch.empty(Btot, dtype=torch.int64, device=dev)
        if by_img:
            w_loc = words_d.detach().contiguous()
            g_w, g_l, g_m = torch.empty(Btot, D, T, device=dev), torch.empty(Btot, dtype=torch.int32, device=dev), torch.empty(Btot, Btot, device=dev)
            l_loc, rs_out = lens_d.to(torch.int32), torch.empty(B, D, T, device=dev)

            def colls():
                dist.all_gather_into_tensor(g_w, w_loc)
                dist.all_gather_into_tensor(g_l, l_loc)
                dist.all_gather_into_tensor(g_cls, cls_d)
                dist.all_gather_into_tensor(g_m, m_blk)
                dist.reduce_scatter_tensor(rs_out, d_words)

            nbytes = {"all_gather_words": Btot * D * T * 4, "reduce_scatter_d_words": Btot * D * T * 4, "all_gather_m": Btot * Btot * 4}
        else:
            img_loc = img_d.detach().reshape(B, D, R)
            g_img, g_m, g_mp = torch.empty(Btot, D, R, device=dev), torch.empty(Btot, B, device=dev), torch.empty(world * Btot, B, device=dev)
            rs_out = torch.empty(B, D, R, device=dev)

            def colls():
                dist.all_gather_into_tensor(g_img, img_loc)
                dist.all_gather_into_tensor(g_cls, cls_d)
                dist.all_gather_into_tensor(g_mp, g_m)
                dist.reduce_scatter_tensor(rs_out, d_img_all)

            nbytes = {"all_gather_img": Btot * D * R * 4, "reduce_scatter_d_img": Btot * D * R * 4, "all_gather_m": world * Btot * B * 4}

        for _ in range(3):
            colls()
        barrier()
        t = torch.tensor([_time_cuda(colls, max(5, args.steps // 2), None)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["nccl_ms_per_step"] = float(t.item())
        res["nccl_bytes_per_step"] = nbytes
        del img_all, d_img_all

    # ---- end to end from pinned host buffers, through the reference-facing API ----
    h2d = img_h.numel() * 4 + words_h.numel() * 4 + lens_h.numel() * 8
    loss_h = torch.empty(2, dtype=torch.float32).pin_memory()
    loss1_h = torch.empty(1, dtype=torch.float32).pin_memory()  # (1 element, not 0-dim: a 0-dim device -> host copy_ synchronises)
    barrier()
    nrun = args.warmup + args.steps
    if world == 1:
        copy_stream = torch.cuda.Stream(device=dev)
        comp = torch.cuda.current_stream()
        # the step's inputs travel as ONE pinned block (img | words | cap_lens) -> one H2D copy per step, as an input pipeline would stage them
        nb = (img_h.numel() * 4, words_h.numel() * 4, lens_h.numel() * 8)
        packed_h = torch.empty(sum(nb), dtype=torch.uint8).pin_memory()
        packed_h[:nb[0]].view(torch.float32).copy_(img_h.reshape(-1))
        packed_h[nb[0]:nb[0] + nb[1]].view(torch.float32).copy_(words_h.reshape(-1))
        packed_h[nb[0] + nb[1]:].view(torch.int64).copy_(lens_h.reshape(-1))

        def staging():
            buf = torch.empty(sum(nb), dtype=torch.uint8, device=dev)
            return dict(buf=buf, img=buf[:nb[0]].view(torch.float32).view(img_d.shape).requires_grad_(),
                        words=buf[nb[0]:nb[0] + nb[1]].view(torch.float32).view(words_d.shape).requires_grad_(),
                        lens=buf[nb[0] + nb[1]:].view(torch.int64), ready=torch.cuda.Event(), free=torch.cuda.Event())

        stg = [staging() for _ in range(2)]
        for sbuf in stg:
            sbuf["free"].record(comp)

        def prefetch(k):
            sbuf = stg[k % 2]
            copy_stream.wait_event(sbuf["free"])
            with torch.cuda.stream(copy_stream):
                sbuf["buf"].copy_(packed_h, non_blocking=True)
            sbuf["ready"].record(copy_stream)

        prefetch(0)
        t_start = None
        host_t0 = 0.0
        for k in range(nrun):
            if k == args.warmup:
                torch.cuda.synchronize()
                t_start = torch.cuda.Event(enable_timing=True)
                t_start.record()
                host_t0 = time.perf_counter()
                prefetch(k)  # the first timed step pays its own copy in full
            if k + 1 < nrun and k + 1 != args.warmup:
                prefetch(k + 1)
            sbuf = stg[k % 2]
            comp.wait_event(sbuf["ready"])
            loss = api_step(sbuf["img"], sbuf["words"], sbuf["lens"])
            sbuf["free"].record(comp)
            loss1_h.copy_(loss.detach().reshape(1), non_blocking=True)  # the step's result: loss0 + loss1
        host_ms = 1e3 * (time.perf_counter() - host_t0) / args.steps
        t_end = torch.cuda.Event(enable_timing=True)
        t_end.record()
        t_end.synchronize()
        e2e_total = t_start.elapsed_time(t_end)
        e2e_mode = ("drop-in API: eegan_b200.words_loss(img, words, labels, cap_lens, class_ids, B) + (loss0 + loss1).backward() per step "
                    "(reference signature, train.py:428); H2D of step k+1 prefetched on a copy stream (double-buffered) while step k computes; "
                    "D2H = the step's loss (loss0 + loss1); both gradients (%.1f MB) stay on the device, where the encoder's backward consumes "
                    "them; host time of the loop %.3f ms per step" % ((img_h.numel() + words_h.numel()) * 4 / 1e6, host_ms))
    else:
        e2e_evs = []
        for k in range(nrun):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            l0, l1, _, _ = sstep(img_h, words_h, lens_h, cls_d)  # H2D straight into the step's static input buffers
            loss_h.copy_(torch.stack([l0.detach(), l1.detach()]), non_blocking=True)
            e1.record()
            e1.synchronize()  # the caller sees the loss on the host
            if k >= args.warmup:
                e2e_evs.append(e0.elapsed_time(e1))
        e2e_total = sum(e2e_evs)
        e2e_mode = "sharded step API (%s); H2D, compute, collectives and D2H of the losses serial in every step; gradients stay on the device" % type(sstep).__name__
    barrier()
    t = torch.tensor([e2e_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    res["e2e"] = {"value": pairs_per_step / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                  "d2h_bytes_per_step": (4 if world == 1 else 8) * world, "ms_per_step": e2e_ms, "mode": e2e_mode}
    res["inputs"] = c
    if sstep is not None and hasattr(sstep, "release_graph"):
        sstep.release_graph()
    del graphed, sstep
    return res


def roofline_block(res, steps):
    pk = peaks()
    stage = res["stage"]
    engine = res["engine"]
    gemm = [s for s in stage if s[0].startswith("gemm")]
    gemm_ms = sum(s[1] for s in gemm)
    nlaunch = sum(s[2] for s in gemm)
    achieved = res["flops_rank"] * steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    kname, div, why = ENGINE_NOTE.get(engine, ENGINE_NOTE[3])
    ceiling = pk["bf16_tflops"] / div
    traffic, tsrc = NCU_TRAFFIC.get(engine, (None, None))
    B, T = res["B_local"], res["T"]
    alg_bytes = 3 * B * D * (R + T) * 4 + 3 * B * B * 4
    return {
        "bound": "tensor", "kernel": kname, "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved / pk["bf16_tflops"] if achieved else None,
        "traffic": traffic if B == 48 else None,
        "traffic_source": ("constant, NOT measured by this run: " + tsrc) if (traffic and B == 48) else None,
        "peak_source": pk["source"] + " cuBLAS bf16 burst (MEASURED_PEAKS.json)",
        "engine_ceiling_tflops": ceiling, "frac_of_engine_ceiling": (achieved / ceiling) if achieved else None,
        "note": "%s = %.0f TFLOP/s; avg GEMM launch %.1f us (%d launches); achieved = algorithmic 12*R*D*sum(T)*B FLOP per step / summed duration "
                "of the GEMM launches (their epilogues carry the softmax work of the path), stage events on WARMED direct launches of the same "
                "kernels (the headline value replays them as one CUDA graph with programmatic dependent launch between them)"
                % (why, ceiling, 1e3 * gemm_ms / max(1, nlaunch), nlaunch),
        "stage_ms_per_step": {s[0]: s[1] / steps for s in stage},
        "whole_step_tflops": res["flops_rank"] / (res["ms_per_step"] / 1e3) / 1e12,
        "hbm_equiv": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (res["ms_per_step"] / 1e3) / 1e9,
                      "peak_gbs": pk["hbm_gbs"]}}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def setup_ours(args):
    from eegan_b200 import _lib
    cx = Ctx()
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and cx.world == 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(cx.local)
    cx.dev = torch.device("cuda", cx.local)
    if cx.world > 1:
        dist.init_process_group("nccl", device_id=cx.dev)
    cx.L = _lib.lib()
    cx.flush = _L2Flush(cx.dev)
    return cx


def parity_block(cx, args):
    from oracle import parity
    try:
        return parity.run(cx.world, cx.rank, cx.dev, mode=args.sharded_mode if cx.world > 1 else "serial", shard=args.shard)
    except Exception as e:  # a parity block that cannot run is a failed one, and says why
        return {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}


def run_ours(args):
    if args.config == "c4":
        return run_full_step(args)
    cx = setup_ours(args)
    world, rank, dev = cx.world, cx.rank, cx.dev
    w = WORKLOADS[args.config]
    parity = parity_block(cx, args)
    sampler = ClockSampler(cx.local) if rank == 0 else None
    t_begin = time.time()
    sweep = None
    if args.config == "c5":
        sweep = []
        for Bg in w["sweep"]:
            if Bg % world or Bg // world < 2:
                continue
            r = time_pair_workload(args, cx, "c5", Bg // world, w["T"], w["cls"], full=(world == 1 and Bg == w["sweep"][-1]))
            sweep.append(r)
        res = sweep[-1]
    elif args.config == "c3":
        assert w["global_B"] % world == 0
        res = time_pair_workload(args, cx, "c3", w["global_B"] // world, w["T"], w["cls"])
    else:
        res = time_pair_workload(args, cx, "c2", w["per_gpu"], w["T"], w["cls"])
    clocks = sampler.summary(t_begin, time.time()) if sampler else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    engine = res["engine"]
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args.config, res["B_total"], res["T"]),
                       "config_id": args.config, "sum_cap_lens": res["lens_sum"], "per_gpu_B": res["B_local"],
                       "sharding": ("%s rows of the grid over %d GPUs (eegan_b200/sharded.py)" % ({"images": "image", "captions": "caption"}.get(
                           res.get("shard_by"), "caption"), world)) if world > 1 else "none",
                       "l2": "256 MB write, then a 256 MB read sweep, between timed steps (outside the timed spans): cold cache, no dirty lines left to write back inside a span",
                       "pairs_per_step": res["pairs_per_step"], "launch": res["launch"], "contraction_engine": engine,
                       "eager_ms_per_step": res.get("eager_ms_per_step")},
            "parity": parity, "e2e": res["e2e"], "gpu_launches": LAUNCHES_PER_STEP.get(engine, 15) * args.steps, "clocks": clocks}
    if "api_ms_per_step" in res:
        line["drop_in_api"] = {"ms_per_step": res["api_ms_per_step"], "pairs_per_s": res["pairs_per_step"] / (res["api_ms_per_step"] / 1e3),
                               "note": res["api_note"]}
    if "stage" in res:
        line["roofline"] = roofline_block(res, args.steps)
    for k in ("comm_free_same_shape", "nccl_ms_per_step", "nccl_bytes_per_step"):
        if k in res:
            line[k] = res[k]
    if sweep is not None:
        line["sweep"] = [{"global_B": r["B_total"], "per_gpu_B": r["B_local"], "ms_per_step": r["ms_per_step"], "pairs_per_s": r["value"],
                          "e2e_pairs_per_s": r["e2e"]["value"], "tflops_algorithmic": algorithmic_flops(r["lens_sum"], r["B_total"]) / (r["ms_per_step"] / 1e3) / 1e12,
                          "frac_of_bf16_peak": algorithmic_flops(r["lens_sum"], r["B_total"]) / (r["ms_per_step"] / 1e3) / 1e12 / (pk["bf16_tflops"] * world),
                          "hbm_frac_algorithmic": (3 * r["B_total"] * D * (R + r["T"]) * 4 + 3 * r["B_total"] ** 2 * 4) / (r["ms_per_step"] / 1e3) / 1e9 / (pk["hbm_gbs"] * world),
                          "drop_in_api_ms": r.get("api_ms_per_step"),
                          "comm_free_ms": (r.get("comm_free_same_shape") or {}).get("ms_per_step"), "nccl_ms": r.get("nccl_ms_per_step")}
                         for r in sweep]
    if world == 1:
        from oracle import ref_loader as RL
        ref = RL.load_reference() if RL.reference_available() else None
        if not args.no_extra and args.config == "c2":
            extra = {"global_attention_general": gag_extra(dev, cx.flush, pk["hbm_gbs"], ref),
                     "sync_batchnorm_1replica": syncbn_extra(dev, cx.flush, pk["hbm_gbs"]),
                     "affine_ssa_1replica": ssa_extra(dev, cx.flush, pk["hbm_gbs"]),
                     "aux_rows_8f": aux_rows_extra(dev, cx.flush, pk["bf16_tflops"])}
            if ref is not None:
                extra["reference_eager_same_gpu"] = {"words_loss": reference_eager_same_gpu(dev, ref, res["inputs"], res["B_local"], cx.flush),
                                                     "global_attention_general": "see extra.global_attention_general[*].ms_reference_eager_same_gpu"}
            line["extra"] = extra
        base, _, _, _, _ = cpu_arm(args.config, args.steps, 1, B=min(res["B_local"], 128))
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
# config 4: the generator step — the reference's own Gen + Trainer.DAMSM_loss with install()
# ---------------------------------------------------------------------------------------
C4_METRIC = "EE-GAN generator step (Gen fwd/bwd + DAMSM_loss) samples/s, CUB shape"


def _c4_inputs(B, T, seed, dev):
    g = torch.Generator(device="cpu").manual_seed(seed)
    d = dict(noise=torch.randn(B, 100, generator=g), sent=torch.randn(B, 256, generator=g) * 0.5,
             words=torch.tanh(torch.randn(B, 256, T, generator=g)) * 0.5, attrs=torch.randn(B, 3, 256, generator=g) * 0.5,
             lens=torch.randint(5, T + 1, (B,), generator=g), cls=torch.randint(1, 18, (B,), generator=g).numpy())
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


def _c4_build(ns, dev, seed, fuse):
    """netG, attr_enhance, image_encoder as train.py:216-262 builds them (random init: no checkpoints offline)."""
    import contextlib
    import io
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        netG = ns.models.Gen(32, 100)
        attr = ns.models.ATTR_Enhance()
        enc = ns.DAMSM.CNN_ENCODER(256)
    # train.py zero-initialised gates (models.py:63-66, 103) would switch the SyncBN / affine paths off in the output: randomise them
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in netG.named_parameters():
            if n.endswith("gamma") or "linear2" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    for p in enc.parameters():
        p.requires_grad = False
    enc.eval()
    if fuse:  # the opt-in one-liners of INTEGRATION.md: same parameter tensors, fused kernels
        import eegan_b200 as E
        E.fuse_emb_features(enc)
        E.fuse_affine_ssa(netG)
    return netG.to(dev).train(), attr.to(dev).train(), enc.to(dev)


def _c4_step(ns, nets, opt, x, B, labels, wrap_cfg):
    netG, attr, enc = nets
    _, attn_attr = attr(x["sent"], x["attrs"])
    attn_attr = ns.models.ATTR_Enhance.attr_merge(attn_attr)  # train.py:188-189
    fake = netG(x["noise"], x["sent"], attn_attr)          # train.py:190
    with wrap_cfg():
        w_loss, s_loss, a_loss = ns.train.Trainer.DAMSM_loss(fake[-1], x["sent"], x["words"], attn_attr, x["cls"], B, labels, x["lens"], enc)
    g_loss = 0.05 * (s_loss + w_loss + a_loss)              # train.py:493 with the default --sim_coe
    opt.zero_grad()
    g_loss.backward()
    opt.step()
    return g_loss.detach(), w_loss.detach(), s_loss.detach(), a_loss.detach()


def run_full_step(args):
    import contextlib
    from oracle import ref_loader as RL
    cx = setup_ours(args)
    world, rank, dev = cx.world, cx.rank, cx.dev
    if not RL.reference_models_available():
        if rank == 0:
            print(json.dumps({"metric": C4_METRIC, "unavailable": "reference models.py / DAMSM.py / train.py not staged under baseline/_ref "
                                                                 "(run oracle/stage_reference.py in the build container)"}))
        return
    import eegan_b200 as E
    ref = RL.load_reference_models()
    inst = RL.load_reference_installed()
    E.damsm_losses.AUTO_SHARD = True  # what eegan_b200.install() sets: global-batch losses from per-rank shards under a process group
    B, T = 32, 18
    parity = parity_block(cx, args)
    x = _c4_inputs(B, T, 100 + rank, dev)
    labels = torch.arange(B, device=dev)

    @contextlib.contextmanager
    def cuda_cfg():
        prev = ref.cfg.CUDA
        ref.cfg.CUDA = True
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                yield
        finally:
            ref.cfg.CUDA = prev

    def build_and_time(ns, fuse, tag):
        nets = _c4_build(ns, dev, 7, fuse)
        if world > 1:  # train.py:220: DataParallelWithCallback(netG); here one process per GPU, gradients all-reduced by the wrapper
            wrapped = (ns.sync_batchnorm.DataParallelWithCallback(nets[0]), ns.sync_batchnorm.DataParallelWithCallback(nets[1]), nets[2])
            params = list(nets[0].parameters()) + list(nets[1].parameters())
        else:
            wrapped = nets
            params = list(nets[0].parameters()) + list(nets[1].parameters())
        opt = torch.optim.Adam(params, lr=2e-4, betas=(0.0, 0.9))  # train.py:264-275
        losses = None
        for _ in range(args.warmup):
            losses = _c4_step(ns, wrapped, opt, x, B, labels, cuda_cfg)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            losses = _c4_step(ns, wrapped, opt, x, B, labels, cuda_cfg)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the DAMSM part alone (image encoder excluded): Trainer.DAMSM_loss's three losses on fixed features
        feats = torch.relu(torch.randn(B, 256, 17, 17, device=dev)).requires_grad_()
        code = torch.randn(B, 256, device=dev).requires_grad_()
        attn_attr = torch.randn(B, 256, device=dev).requires_grad_()

        class _Fixed(torch.nn.Module):
            def forward(self, _x):
                return feats, code

        def damsm_only():
            feats.grad = code.grad = attn_attr.grad = None
            with cuda_cfg():
                wl, sl, al = ns.train.Trainer.DAMSM_loss(None, x["sent"], x["words"], attn_attr, x["cls"], B, labels, x["lens"], _Fixed())
            (wl + sl + al).backward()

        for _ in range(3):
            damsm_only()
        ms_damsm = _time_cuda(damsm_only, 5, None)
        del opt, nets, wrapped
        torch.cuda.empty_cache()
        return {"arm": tag, "ms_per_step": float(t.item()), "samples_per_s": B * world / (float(t.item()) / 1e3),
                "ms_damsm_loss_fwd_bwd_alone": ms_damsm, "losses_last_step": [float(v) for v in losses]}

    sampler = ClockSampler(cx.local) if rank == 0 else None
    t0 = time.time()
    ours = build_and_time(inst, True, "reference Gen / ATTR_Enhance / Trainer.DAMSM_loss imported with eegan_b200.install(): 24 eegan_b200 SyncBN "
                                      "layers (14 of them inside fused affine_ssa, eegan_b200.fuse_affine_ssa), eegan_b200 words_loss / sent_loss "
                                      "(global-batch losses over NCCL when N > 1), fused emb_features")
    refarm = None
    if world == 1:
        refarm = build_and_time(ref, False, "the unmodified reference modules and losses, eager on the same GPU (cfg.CUDA=True)")
    clocks = sampler.summary(t0, time.time()) if sampler else None
    if rank == 0:
        line = {"metric": C4_METRIC, "value": ours["samples_per_s"], "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ours["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[3]: EE-GAN generator step, CUB shape, B=%d per GPU (global %d): ATTR_Enhance -> Gen(ngf=32) "
                                       "-> 256x256 image -> CNN_ENCODER (Inception v3, frozen, random init) -> Trainer.DAMSM_loss (words + 2 x sent) "
                                       "-> backward -> Adam step; discriminators not included" % (B, B * world),
                           "config_id": "c4", "sync_batchnorm_layers": 24,
                           "syncbn_all_reduces_per_step": 48 if world > 1 else 0},
                "parity": parity, "ours": ours, "reference_same_gpu": refarm, "clocks": clocks,
                "speedup_vs_reference_same_gpu": (refarm["ms_per_step"] / ours["ms_per_step"]) if refarm else None,
                "e2e": {"value": ours["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "mode": "inputs (noise, embeddings) are produced on the device by the text encoder in train.py; nothing to copy"},
                "gpu_launches": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_full_step_reference_cpu(args):
    from oracle import ref_loader as RL
    if not RL.reference_models_available():
        print(json.dumps({"impl": "reference", "unavailable": "reference models not staged under baseline/_ref"}))
        return
    import contextlib
    ref = RL.load_reference_models()
    ref.cfg.CUDA = False
    torch.set_num_threads(os.cpu_count() or 1)
    B, T = 4, 18
    dev = torch.device("cpu")
    x = _c4_inputs(B, T, 100, dev)
    nets = _c4_build(ref, dev, 7, False)
    opt = torch.optim.Adam(list(nets[0].parameters()) + list(nets[1].parameters()), lr=2e-4, betas=(0.0, 0.9))
    labels = torch.arange(B)

    @contextlib.contextmanager
    def nocfg():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yield

    _c4_step(ref, nets, opt, x, B, labels, nocfg)
    times = []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        _c4_step(ref, nets, opt, x, B, labels, nocfg)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > 90:
            break
    ms = 1e3 * sum(times) / len(times)
    base = dict(value=B / (ms / 1e3), unit="samples/s", cores=os.cpu_count() or 1, kind="reference",
                sample="%d generator steps at B=%d (bounded sample of the B=32 step), unmodified reference modules on the host cores" % (len(times), B))
    print(json.dumps({"impl": "reference", "metric": C4_METRIC, "value": base["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": len(times),
                      "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": "BASELINE.json configs[3] generator step, B=%d sample on the CPU" % B, "config_id": "c4"},
                      "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--no-extra", action="store_true", help="skip the GlobalAttentionGeneral / SyncBN / reference-eager side measurements")
    ap.add_argument("--shard", default=None, choices=["images", "captions"],
                    help="N > 1: partition of the pair grid (default: eegan_b200.sharded.SHARD_BY = images: own images x all captions, the word "
                         "features travel; captions: own captions x all images, the region features travel)")
    ap.add_argument("--sharded-mode", default=os.environ.get("EEGAN_SHARDED_MODE", "serial"), choices=["serial", "overlap", "graph"],
                    help="N>1: ShardedWordsLossStep with direct launches (serial), the same captured into one CUDA graph, collectives "
                         "included (graph), or OverlappedShardedWordsLossStep (collectives hidden behind the local image block)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.config == "c4" and args.steps == 100:
        args.steps, args.warmup = 10, 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

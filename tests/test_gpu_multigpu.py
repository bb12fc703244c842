"""GPU, N >= 2 devices on one box (skipped on a single-GPU box): tests/multigpu_check.py under torchrun — the
caption-row-sharded words_loss / sent_loss (autograd route, step class, overlapped step, graph-captured step),
SynchronizedBatchNorm2d and fused affine_ssa over NCCL against the single-device full-batch result — and the
bench's own parity block at N ranks.  `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multigpu.py -m gpu` runs it;
the round's logs are committed under profiles/."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(n, script_args, env=None, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(_port())] + script_args
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box")
def test_multigpu_check_under_torchrun():
    n = min(_ngpu(), 8)
    r = _torchrun(n, ["tests/multigpu_check.py"], env={"EEGAN_CHECK_SHARDED_OVERLAP": "1", "EEGAN_CHECK_SHARDED_GRAPH": "1"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multigpu_check ok: world %d" % n in r.stdout


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box")
@pytest.mark.parametrize("mode", ["serial", "overlap", "graph"])
def test_bench_parity_block_at_n_ranks(mode):
    n = min(_ngpu(), 8)
    r = _torchrun(n, ["bench.py", "--gpus", str(n), "--steps", "3", "--warmup", "3", "--sharded-mode", mode])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == n and line["parity"]["ok"], line["parity"]
    assert line["parity"]["collectives"] == "nccl"

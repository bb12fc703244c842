"""GEMM2-shaped launch (M=640, N=256, K=289, batch 48) through the stand-alone entry: 128-wide vs 256-wide tiles (EEGAN_H_WIDE)."""
import os, sys, torch
sys.path.insert(0, ".")
from eegan_b200 import _lib
L = _lib.lib()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def run(M, N, K, batch, n=20):
    pad8 = lambda v: (v + 7) // 8 * 8
    lda, ldb = pad8(M), pad8(K)
    A = torch.randn(batch, K, lda, device="cuda"); B = torch.randn(batch, N, ldb, device="cuda"); C = torch.empty(batch, M, N, device="cuda")
    ws = torch.empty(2 * (A.numel() * 2 + B.numel() * 2 + 512) + 512, dtype=torch.uint8, device="cuda")
    def call(flag):
        _lib.check(L.eegan_gemm_f16x3(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, lda, ldb, N, A.stride(0), B.stride(0), C.stride(0), batch, 64.0, 64.0, flag, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    call(0); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(4); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    ref = torch.bmm(A[:1, :, :M].transpose(1, 2).double(), B[:1, :, :K].transpose(1, 2).double())
    print("WIDE=%s M=%d N=%d K=%d b=%d: median %.1f us min %.1f us  err %.2e" % (os.environ.get("EEGAN_H_WIDE"), M, N, K, batch, ts[len(ts)//2], ts[0], (C[:1].double() - ref).abs().max().item()))
run(640, 256, 289, 48)
run(640, 256, 1156, 48)
run(1024, 512, 2048, 8)

"""Per-launch cost model of the half-pair GEMM: fixed overhead + blocks * cycles."""
import sys, torch
sys.path.insert(0, ".")
from eegan_b200 import _lib
L = _lib.lib()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def run(M, N, K, batch, n=20, cold=False):
    pad8 = lambda v: (v + 7) // 8 * 8
    lda, ldb = pad8(M), pad8(K)
    A = torch.randn(batch, K, lda, device="cuda")
    B = torch.randn(batch, N, ldb, device="cuda")
    C = torch.empty(batch, M, N, device="cuda")
    ws = torch.empty(2 * (A.numel() * 2 + B.numel() * 2 + 512) + 512, dtype=torch.uint8, device="cuda")
    def call(flag):
        _lib.check(L.eegan_gemm_f16x3(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, lda, ldb, N, A.stride(0), B.stride(0), C.stride(0), batch, 64.0, 64.0, flag, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    call(0); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(4); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    tiles = ((M + 127) // 128) * ((N + 127) // 128) * batch
    kb = (K + 31) // 32
    per_sm = -(-tiles // 148) * kb
    print("M=%4d N=%4d K=%4d b=%3d %s tiles=%4d blocks/SM(max)=%3d  median %.1f us  min %.1f us  -> %.0f cyc/block" %
          (M, N, K, batch, "cold" if cold else "warm", tiles, per_sm, ts[len(ts) // 2], ts[0], ts[len(ts) // 2] * 1.965e3 / per_sm))
for cold in (False, True):
    run(128, 128, 32, 148, cold=cold)
    run(128, 128, 256, 148, cold=cold)
    run(128, 128, 1024, 148, cold=cold)
    run(640, 256, 289, 48, cold=cold)
    run(640, 256, 288, 48, cold=cold)
    run(289, 640, 256, 48, cold=cold)
    run(256, 289, 1280, 48, cold=cold)
    run(256, 256, 1280, 48, cold=cold)

"""Main-loop probe: half-pair engine vs the TMEM-staged 3xTF32 engine at K = 2048 (epilogue negligible)."""
import sys, torch
sys.path.insert(0, ".")
from eegan_b200 import _lib
L = _lib.lib()
M, N, K, batch = 1024, 1024, 2048, 16
A = torch.randn(batch, K, M, device="cuda")
B = torch.randn(batch, N, K, device="cuda")
C = torch.empty(batch, M, N, device="cuda")
ws = torch.empty(2 * (A.numel() * 2 + B.numel() * 2 + 512) + 512, dtype=torch.uint8, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def ts():
    _lib.check(L.eegan_gemm_tf32x3(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, 0, 1, M, K, N, A.stride(0), B.stride(0), C.stride(0), batch, 1, _lib.stream_ptr()))
def hh(k=K):
    _lib.check(L.eegan_gemm_f16x3(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, k, M, K, N, A.stride(0), B.stride(0), C.stride(0), batch, 64.0, 64.0, 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
blocks = (M // 128) * (N // 128) * batch * (K // 32)
t_ts = t(ts)
t_h = t(hh)
t_h0 = t(lambda: hh(32))   # same split kernels, 1/64 of the main loop
flop = 2.0 * M * N * K * batch
cyc = lambda ms: ms * 1e-3 * 1.965e9 / (blocks / 148)
print("ts   %.3f ms  %.1f TFLOP/s  %.0f cyc/block" % (t_ts, flop / t_ts / 1e9, cyc(t_ts)))
print("half %.3f ms total, %.3f ms with K=32 (split kernels + epilogue) -> main loop %.3f ms  %.1f TFLOP/s  %.0f cyc/block"
      % (t_h, t_h0, t_h - t_h0, flop / (t_h - t_h0) / 1e9, cyc(t_h - t_h0)))
hh()
ref = torch.bmm(A[:2].transpose(1, 2).double(), B[:2].transpose(1, 2).double())
print("half err", (C[:2].double() - ref).abs().max().item())

import sys, torch
sys.path.insert(0, '/root/repo')
from eegan_b200 import _lib
L = _lib.lib()
def run(M, N, K, batch, shared, staging, iters=20):
    nb = 1 if shared else batch
    A = torch.randn(nb, K, M, device='cuda')      # MN-major A: [K][M]
    B = torch.randn(nb, N, K, device='cuda')      # K-major B: [N][K]
    C = torch.empty(batch, M, N, device='cuda')
    bsA = 0 if shared else A.stride(0); bsB = 0 if shared else B.stride(0)
    def go():
        _lib.check(L.eegan_gemm_tf32x3(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, 0, 1, M, K, N, bsA, bsB, C.stride(0), batch, staging, _lib.stream_ptr()))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): go()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tiles = ((M + 127) // 128) * ((N + 127) // 128) * batch
    rounds = -(-tiles // 148)
    kb = (K + 31) // 32
    cyc_per_kb = us * 1965.0 / (rounds * kb)
    print("M%d N%d K%d batch%d shared=%d staging=%d: %.1f us, %.0f cycles/k-block (floor 768), %.0f TF eff" % (M, N, K, batch, shared, staging, us, cyc_per_kb, 2.0 * M * N * K * batch / us * 1e-6), flush=True)
for staging in (0, 1):
    for shared in (1, 0):
        run(128, 128, 2048, 148, shared, staging)
    run(128, 128, 256, 148 * 5, 0, staging)
    run(128, 128, 256, 148 * 5, 1, staging)

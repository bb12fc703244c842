import sys, os, torch
sys.path.insert(0, '/root/repo')
from eegan_b200 import _lib
L = _lib.lib()
def run(M, N, K, batch, shared, staging, iters=20):
    nb = 1 if shared else batch
    A = torch.randn(nb, K, M, device='cuda'); B = torch.randn(nb, N, K, device='cuda')
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
    rounds = -(-tiles // 148); kb = (K + 31) // 32
    print("dbg=%s K%d batch%d shared=%d: %.1f us, %.0f cycles/k-block" % (os.environ.get('EEGAN_TS_DBG'), K, batch, shared, us, us * 1965.0 / (rounds * kb)), flush=True)
run(128, 128, 2048, 148, 1, 1)
run(128, 128, 256, 740, 1, 1)

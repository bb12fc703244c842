import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from eegan_b200 import _lib
L = _lib.lib()
def run(M, N, K, a_k, b_k):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(1, M, K, generator=g); B = torch.randn(1, N, K, generator=g)
    Ab = A.clone() if a_k else A.transpose(1, 2).contiguous()
    Bb = B.clone() if b_k else B.transpose(1, 2).contiguous()
    lda = K if a_k else M; ldb = K if b_k else N
    Ad, Bd = Ab.cuda(), Bb.cuda(); C = torch.zeros(1, M, N, device='cuda')
    _lib.check(L.eegan_gemm_tf32x3(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(C), M, N, K, int(a_k), int(b_k), lda, ldb, N, Ad.stride(0), Bd.stride(0), C.stride(0), 1, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))[0]
    return float((C[0].cpu().double() - ref).abs().max())
for tr in ('0', '1'):
    os.environ['EEGAN_TC_TRUNC_HI'] = tr
    print('trunc_hi', tr, [('%.2e' % run(256, 256, 256, a, b)) for a, b in ((True, True), (True, False), (False, False))], flush=True)

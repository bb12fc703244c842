import sys, torch
sys.path.insert(0, '/root/repo')
import eegan_b200 as E
dev = torch.device('cuda')
g = torch.Generator().manual_seed(7)
Bq, T = 48, 18
lens = torch.randint(5, T + 1, (Bq,), generator=g)
mask = (torch.arange(T)[None, :] >= lens[:, None]).to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def timeit(fn, n=6, rep=4):
    # rep back-to-back calls per timed span so that the span is GPU-bound, not launch-latency-bound
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(rep): fn()
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) / rep
    return tot / n
for res, idf in ((64, 128), (128, 64), (256, 32)):
    x = torch.randn(Bq, idf, res, res, device=dev).requires_grad_()
    key = (torch.randn(Bq, idf, T, device=dev) * idf ** -0.5).requires_grad_()
    val = torch.randn(Bq, idf, T, device=dev).requires_grad_()
    mod = E.GlobalAttentionGeneral(idf, 256); mod.applyMask(mask)
    go = torch.randn(Bq, idf, res, res, device=dev); ga = torch.randn(Bq, T, res, res, device=dev)
    o, a = mod(x, key, val)
    def bwd():
        x.grad = key.grad = val.grad = None
        torch.autograd.backward([o, a], [go, ga], retain_graph=True)
    def fwd():
        with torch.no_grad(): mod(x, key, val)
    rows = Bq * res * res
    mf, mb = timeit(fwd), timeit(bwd)
    by_f = (2 * idf + T) * 4 * rows; by_b = (3 * idf + T) * 4 * rows
    print("res %d idf %d: fwd %.3f ms (%.0f GB/s, %.2f of HBM)  bwd %.3f ms (%.0f GB/s, %.2f)" % (res, idf, mf, by_f / mf / 1e6, by_f / mf / 1e6 / 6549, mb, by_b / mb / 1e6, by_b / mb / 1e6 / 6549), flush=True)

import sys, torch
sys.path.insert(0, '/root/repo')
import eegan_b200 as E
dev = torch.device('cuda')
g = torch.Generator().manual_seed(7)
Bq, T = 48, 18
res, idf = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 128)
lens = torch.randint(5, T + 1, (Bq,), generator=g)
mask = (torch.arange(T)[None, :] >= lens[:, None]).to(dev)
x = torch.randn(Bq, idf, res, res, device=dev).requires_grad_()
key = (torch.randn(Bq, idf, T, device=dev) * idf ** -0.5).requires_grad_()
val = torch.randn(Bq, idf, T, device=dev).requires_grad_()
mod = E.GlobalAttentionGeneral(idf, 256); mod.applyMask(mask)
go = torch.randn(Bq, idf, res, res, device=dev); ga = torch.randn(Bq, T, res, res, device=dev)
for _ in range(2):
    o, a = mod(x, key, val)
    torch.autograd.backward([o, a], [go, ga])
torch.cuda.synchronize()
print("ok")

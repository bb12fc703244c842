"""Stage times of the pair-grid forward / backward at B=48 (direct launches, stage events): python profiles/tools/fwd_stage.py [B] [T]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from eegan_b200 import _lib, damsm_losses as dl
from oracle import cases
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
T = int(sys.argv[2]) if len(sys.argv) > 2 else 18
c = cases.words_case(B, T, seed=3407)
img, words = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
lens, cls, labels = c["cap_lens"].cuda(), c["class_ids"].cuda(), c["labels"].cuda()
L = _lib.lib()
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def step():
    img.grad = words.grad = None
    m, _ = dl.pair_grid(img, words, lens)
    l0, l1, _ = dl._PairCEFn.apply(m, 10.0, cls, labels)
    (l0 + l1).backward()
for _ in range(3): step()
torch.cuda.synchronize()
L.eegan_profile_enable(1)
N = 20
for _ in range(N):
    flush.fill_(1.0); step()
torch.cuda.synchronize()
n = L.eegan_profile_nstages()
ms, cnt = (ctypes.c_double * n)(), (ctypes.c_int * n)()
L.eegan_profile_collect(ms, cnt)
print(os.environ.get("EEGAN_HF_DBG", "-"), " ".join("%s=%.1f" % (L.eegan_profile_stage_name(i).decode().split("(")[0], 1e3 * ms[i] / N) for i in range(n) if cnt[i]))

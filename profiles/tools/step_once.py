"""A few direct-launch words_loss steps (forward + backward, B=48 CUB shape) for ncu captures:
    ncu ... python profiles/tools/step_once.py [B] [T] [nsteps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from eegan_b200 import damsm_losses as dl
from oracle import cases
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
T = int(sys.argv[2]) if len(sys.argv) > 2 else 18
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
c = cases.words_case(B, T, seed=3407)
img, words = c["img"].cuda().requires_grad_(), c["words"].cuda().requires_grad_()
lens, cls, labels = c["cap_lens"].cuda(), c["class_ids"].cuda(), c["labels"].cuda()
for _ in range(n):
    img.grad = words.grad = None
    m, _ = dl.pair_grid(img, words, lens)
    l0, l1, _ = dl._PairCEFn.apply(m, 10.0, cls, labels)
    (l0 + l1).backward()
torch.cuda.synchronize()
print("ok", float(l0), float(l1))

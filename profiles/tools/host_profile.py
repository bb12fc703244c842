"""Where the HOST time of the drop-in call goes: words_loss(...) + backward at B=48 (run on the GPU box).
    python profiles/tools/host_profile.py [B]
Prints per-step wall time of the whole call and of its parts, and a cProfile table of the steady state."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import eegan_b200 as E
from eegan_b200 import fastpath
from oracle import cases

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
T = 18
c = cases.words_case(B, T, seed=1)
dev = torch.device("cuda")
img, words = c["img"].to(dev).requires_grad_(), c["words"].to(dev).requires_grad_()
labels, lens, cls = c["labels"].to(dev), c["cap_lens"].to(dev), c["class_ids"]


def step():
    img.grad = None
    words.grad = None
    l0, l1, _ = E.words_loss(img, words, labels, lens, cls, B)
    (l0 + l1).backward()


def fwd_only():
    with torch.no_grad():
        E.words_loss(img, words, labels, lens, cls, B)


def timeit(fn, n=300):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    return 1e6 * t_host / n, 1e6 * t_all / n


print("full step          host %.1f us  wall %.1f us" % timeit(step))
print("forward (no_grad)  host %.1f us  wall %.1f us" % timeit(fwd_only))
plan = next(iter(fastpath._plans.values()))
print("plan.load          host %.1f us  wall %.1f us" % timeit(lambda: plan.load(img.detach().reshape(B, 256, -1), words.detach(), lens, cls, labels)))
print("fwd graph replay   host %.1f us  wall %.1f us" % timeit(plan.fwd_graph.replay))
print("bwd graph replay   host %.1f us  wall %.1f us" % timeit(plan.bwd_graphs[(True, True)].replay))
print("clone d_img        host %.1f us  wall %.1f us" % timeit(lambda: plan.d_img.clone()))
x = torch.zeros((), device=dev, requires_grad=True)


def tiny():
    x.grad = None
    (x * 2).backward()


print("tiny autograd step host %.1f us  wall %.1f us" % timeit(tiny))
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

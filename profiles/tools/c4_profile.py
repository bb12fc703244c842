"""Kernel-time census of the config-4 generator step (ours arm): python profiles/tools/c4_profile.py [fuse_ssa]"""
import os, sys, contextlib, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from oracle import ref_loader as RL
import eegan_b200 as E
dev = torch.device("cuda")
ref = RL.load_reference_models(); inst = RL.load_reference_installed()
B, T = 32, 18
x = bench._c4_inputs(B, T, 100, dev)
labels = torch.arange(B, device=dev)
nets = bench._c4_build(inst, dev, 7, True)
if len(sys.argv) > 1 and sys.argv[1] == "fuse_ssa":
    print("fused affine_ssa layers:", E.fuse_affine_ssa(nets[0]))
opt = torch.optim.Adam(list(nets[0].parameters()) + list(nets[1].parameters()), lr=2e-4, betas=(0.0, 0.9))
@contextlib.contextmanager
def nocfg():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore"); yield
for _ in range(3): bench._c4_step(inst, nets, opt, x, B, labels, nocfg)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): bench._c4_step(inst, nets, opt, x, B, labels, nocfg)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print("ms/step wall %.2f  host-issue %.2f" % (1e3 * (time.perf_counter() - t0) / 5, 1e3 * t_host / 5))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): bench._c4_step(inst, nets, opt, x, B, labels, nocfg)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("GPU kernel time per step: %.2f ms" % tot)
for k, ms, n in rows[:32]: print("%7.3f ms %5d  %s" % (ms, n, k[:110]))

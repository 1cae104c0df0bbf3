"""Where does the step time go outside the kernels when the logits are kept?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from bench import WORKLOADS, synth
lib = _lib.load()
dev = torch.device("cuda", 0)
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)

def ev():
    return torch.cuda.Event(enable_timing=True)

for keep in (False, True):
    F.set_keep_activations(keep)
    tf = tb = 0.0
    n = 12
    for i in range(n + 4):
        e0, e1, e2 = ev(), ev(), ev()
        fd.grad = gd.grad = Wd.grad = bd.grad = None
        e0.record()
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        s = loss.sum()
        e1.record()
        s.backward()
        e2.record()
        torch.cuda.synchronize()
        if i >= 4:
            tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    print(f"keep={keep}: forward op {tf / n:.3f} ms, backward {tb / n:.3f} ms", flush=True)
    # raw allocation cost
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        x = torch.empty(3489660928, dtype=torch.uint8, device=dev); del x
    torch.cuda.synchronize(); print(f"  torch.empty(3.5 GB) + free: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms", flush=True)
from torch.profiler import profile, ProfilerActivity
F.set_keep_activations(True)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        fd.grad = gd.grad = Wd.grad = bd.grad = None
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        loss.sum().backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

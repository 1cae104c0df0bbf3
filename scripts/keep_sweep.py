"""Kept-activation schedule: step time against ring slots and K-groups (usage: keep_sweep.py target "2,4 3,4 4,4 2,3")."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
dev = torch.device("cuda", 0)
wl = sys.argv[1]
B, T, U, V, H, _ = WORKLOADS[wl]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)

def step():
    fd.grad = gd.grad = Wd.grad = bd.grad = None
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    loss.sum().backward()

for combo in sys.argv[2].split():
    ns, kg = (int(x) for x in combo.split(","))
    lib.rnnt_debug_set(b"ring_slots", ns)
    lib.rnnt_debug_set(b"mega_kg_kept", kg)
    for _ in range(6):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record(); torch.cuda.synchronize()
    print(f"{wl} ring_slots={ns} kg_kept={kg}: {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)

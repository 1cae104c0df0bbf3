"""Backward mega-kernel role split: sweep the number of K-groups of dW consumers (rnnt_debug_set("mega_kg", k)) for a
workload and report fwd+bwd ms per step (CUDA events, 20 steps after 5 warm-up)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
kgs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
if "," in name:   # ad-hoc shape "B,T,U,V,H"
    B, T, U, V, H = (int(x) for x in name.split(","))
else:
    B, T, U, V, H, desc = WORKLOADS[name]
dev = torch.device("cuda", 0)
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
lib = _lib.load()
def step():
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    loss.sum().backward()
    out = (float(loss.sum()), float(Wd.grad.abs().sum()))
    fd.grad = gd.grad = Wd.grad = bd.grad = None
    return out
ref = None
for kg in kgs:
    lib.rnnt_debug_set(b"mega_kg", kg)
    for _ in range(5):
        chk = step()
    ref = ref or chk
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        loss.sum().backward()
        fd.grad = gd.grad = Wd.grad = bd.grad = None
    e1.record(); torch.cuda.synchronize()
    print(f"{name} KG={kg}: {e0.elapsed_time(e1) / 20:.3f} ms/step  loss {chk[0]:.4f} |dW|_1 {chk[1]:.4f} (ref {ref[0]:.4f} {ref[1]:.4f})", flush=True)

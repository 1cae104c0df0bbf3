"""fwd+bwd ms per step at a wide vocabulary (default B=8 T=500 U=100 V=4096 H=1024) on both schedules of the library."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import synth
B, T, U, V, H = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,500,100,4096,1024").split(","))
dev = torch.device("cuda", 0)
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
lib = _lib.load()
res = {}
for name, path in (("persistent", 1), ("per-slab", 0)):
    lib.rnnt_debug_set(b"path", path)
    for _ in range(5):
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1); loss.sum().backward()
        chk = (float(loss.sum()), float(Wd.grad.float().abs().sum()))
        fd.grad = gd.grad = Wd.grad = bd.grad = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1); loss.sum().backward()
        fd.grad = gd.grad = Wd.grad = bd.grad = None
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tf = 6.0 * B * T * (U + 1) * H * V / (ms * 1e-3) / 1e12
    print(f"B={B} T={T} U={U} V={V} H={H} {name}: {ms:.3f} ms/step = {tf:.0f} TFLOP/s algorithmic, loss {chk[0]:.3f} |dW|_1 {chk[1]:.1f}", flush=True)
lib.rnnt_debug_set(b"path", 1)

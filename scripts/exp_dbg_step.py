"""Timing experiment: fwd+bwd ms per step with bring-up switches (gemm_dbg bits) set for BOTH passes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
wl = sys.argv[2] if len(sys.argv) > 2 else "target"
B, T, U, V, H, _ = WORKLOADS[wl]
dev = torch.device("cuda", 0)
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
lib = _lib.load()
for dbg in [int(x) for x in sys.argv[1].split(",")]:
    lib.rnnt_debug_set(b"gemm_dbg", dbg)
    def step():
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        loss.sum().backward()
        fd.grad = gd.grad = Wd.grad = bd.grad = None
        return loss
    for _ in range(8):
        l = step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30):
        l = step()
    e1.record(); torch.cuda.synchronize()
    print(f"{wl} gemm_dbg={dbg}: {e0.elapsed_time(e1) / 30:.3f} ms/step, loss sum {float(l.sum()):.3f}", flush=True)
lib.rnnt_debug_set(b"gemm_dbg", 0)

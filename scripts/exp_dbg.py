"""Timing experiment: backward ms per step with bring-up switches (gemm_dbg bits) set; results are garbage for non-zero bits."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
B, T, U, V, H, _ = WORKLOADS["target"]
dev = torch.device("cuda", 0)
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
lib = _lib.load()
for dbg in [int(x) for x in sys.argv[1].split(",")]:
    lib.rnnt_debug_set(b"gemm_dbg", 0)
    losses = []
    for _ in range(8):
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        lib.rnnt_debug_set(b"gemm_dbg", dbg)
        loss.sum().backward()
        lib.rnnt_debug_set(b"gemm_dbg", 0)
        fd.grad = gd.grad = Wd.grad = bd.grad = None
    tot = 0.0
    for _ in range(30):
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        s = loss.sum()
        lib.rnnt_debug_set(b"gemm_dbg", dbg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.backward(); e1.record()
        torch.cuda.synchronize()
        lib.rnnt_debug_set(b"gemm_dbg", 0)
        tot += e0.elapsed_time(e1)
        fd.grad = gd.grad = Wd.grad = bd.grad = None
    print(f"gemm_dbg={dbg}: backward {tot / 30:.3f} ms", flush=True)

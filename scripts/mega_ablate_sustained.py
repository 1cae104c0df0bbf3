"""Sustained (power-capped) cost of parts of the backward mega-kernel: 25 back-to-back backward passes per
configuration with pieces of the producer epilogue switched off via gemm_dbg bits (results are garbage)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
for t in (fd, gd, Wd, bd):
    t.requires_grad_(True)
cfgs = [(0, "full"), (32, "no db sums"), (64, "no dh reduce"), (64 | 128, "no dh tile+reduce"), (32 | 64 | 128, "no db, no dh tile+reduce"), (0, "full again")]
N = 25
for dbg, label in cfgs:
    lib.rnnt_debug_set(b"gemm_dbg", dbg)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    tot = loss.sum()
    for it in range(3):
        tot.backward(retain_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(N):
        tot.backward(retain_graph=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{label:28s} {e0.elapsed_time(e1) / N:7.3f} ms per backward (sustained, {N} passes)", flush=True)
lib.rnnt_debug_set(b"gemm_dbg", 0)

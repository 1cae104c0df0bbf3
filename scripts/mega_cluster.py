"""Backward mega-kernel with CTA pairs (148 CTAs) vs 4-clusters with operand multicast (co-resident 132 CTAs)."""
import os, sys
import numpy as np
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
ref = None
for cs in (2, 4, 2, 4):
    lib.rnnt_debug_set(b"cluster_bwd", cs)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    tot = loss.sum()
    ms = []
    for it in range(12):
        fd.grad = gd.grad = Wd.grad = bd.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tot.backward(retain_graph=True); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    dW = Wd.grad.float().cpu().numpy()
    if ref is None:
        ref = dW
    err = float(np.linalg.norm(dW - ref) / np.linalg.norm(ref))
    print(f"cluster_bwd {cs}: backward ms " + " ".join(f"{m:.2f}" for m in ms[2:]) + f"   dW rel diff vs first {err:.2e}", flush=True)

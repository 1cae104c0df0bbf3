"""A/B of the exact tile list (prune 0) against walking every tile (prune -1), alternating, per schedule."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from bench import WORKLOADS, synth
lib = _lib.load()
dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "target"
B, T, U, V, H, _ = WORKLOADS[wl]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, yd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True), y.to(dev)
Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
lib.rnnt_debug_set(b"time_kernels", 1)
import ctypes
def step():
    fd.grad = gd.grad = Wd.grad = bd.grad = None
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    loss.sum().backward()
def kernel_ms():
    ms = (ctypes.c_double * 9)(); cnt = (ctypes.c_longlong * 9)()
    lib.rnnt_debug_kernel_times(ms, cnt, 9)
    return {n: ms[i] / max(cnt[i], 1) for i, n in enumerate(["hgen", "fwd", "dz", "dh", "dw", "lattice", "coefs", "misc", "mega"]) if cnt[i]}
for keep in (True, False):
    F.set_keep_activations(keep)
    for rnd in range(3):
        for prune in (0, -1):
            lib.rnnt_debug_set(b"prune", prune)
            for _ in range(3):
                step()
            torch.cuda.synchronize(); kernel_ms()
            for _ in range(10):
                step()
            torch.cuda.synchronize()
            k = kernel_ms()
            print(f"{wl} keep={keep} prune={prune:2d}: mega {k['mega']:.3f} ms  fwd {k['fwd']:.3f} ms", flush=True)
lib.rnnt_debug_set(b"prune", 0)

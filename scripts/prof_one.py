"""Two fwd+bwd passes at the BASELINE target shape (for ncu: pass 1 warms up, pass 2 is profiled)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M  # noqa: E402
from bench import WORKLOADS, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "target"
if "--plain-launch" in sys.argv:  # ncu --set full cannot replay the cooperative launch of the backward mega-kernel
    from myrtlespeech_b200 import _lib
    _lib.load().rnnt_debug_set(b"mega_cooperative", 0)
B, T, U, V, H, _ = WORKLOADS[name]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd = f.cuda().requires_grad_(True); gd = g.cuda().requires_grad_(True)
Wd = W.cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
for it in range(2):
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), fl, yl, V - 1)
    loss.sum().backward()
    torch.cuda.synchronize()
    fd.grad = gd.grad = Wd.grad = bd.grad = None
print("ok", float(loss[0]))

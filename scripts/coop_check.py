import os, sys
sys.path.insert(0, "/root/repo")
import torch
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
for t in (fd, gd, Wd, bd): t.requires_grad_(True)
for it in range(4):
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.sum().backward(); e1.record(); torch.cuda.synchronize()
    print("backward ms", round(e0.elapsed_time(e1), 3), "cooperative launch in use:", lib.rnnt_debug_get(b"mega_cooperative"), "db.sum", float(bd.grad.sum()))
    fd.grad = gd.grad = Wd.grad = bd.grad = None

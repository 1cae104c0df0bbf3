"""One small fwd+bwd on both persistent schedules (for compute-sanitizer)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import functional as F
from bench import synth
dev = torch.device("cuda", 0)
for shape in [(3, 37, 11, 300, 128), (2, 20, 9, 29, 512)]:
    B, T, U, V, H = shape
    f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1, dev)
    fl = torch.tensor([max(1, T - 5 * i) for i in range(B)], dtype=torch.int32)
    yl = torch.tensor([max(0, U - 3 * i) for i in range(B)], dtype=torch.int32)
    for keep in (True, False):
        F.set_keep_activations(keep)
        fd, gd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True)
        Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.to(dev), fl, yl, V - 1)
        loss.sum().backward()
        torch.cuda.synchronize()
        print(shape, "keep" if keep else "recompute", float(loss.sum()), float(fd.grad.abs().sum()), flush=True)

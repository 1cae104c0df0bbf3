"""Kept logits against recomputed logits: gradient agreement and step time, per workload.
usage: keep_vs_recompute.py small|c2|target|c3 [...]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from bench import WORKLOADS, synth

lib = _lib.load()
dev = torch.device("cuda", 0)
SHAPES = dict(WORKLOADS)
SHAPES["small"] = (3, 37, 11, 300, 128, None)
SHAPES["odd"] = (5, 70, 23, 1000, 520, None)


def run(wl, keep, steps):
    B, T, U, V, H, _ = SHAPES[wl]
    f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
    if wl in ("small", "odd"):
        fl = torch.tensor([max(1, T - 3 * i) for i in range(B)], dtype=torch.int32)
        yl = torch.tensor([max(0, U - 2 * i) for i in range(B)], dtype=torch.int32)
    fd, gd, yd = f.to(dev).float().requires_grad_(True), g.to(dev).float().requires_grad_(True), y.to(dev)
    Wd, bd = W.to(dev).float().requires_grad_(True), bias.to(dev).float().requires_grad_(True)
    F.set_keep_activations(keep)

    def step():
        fd.grad = gd.grad = Wd.grad = bd.grad = None
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        loss.sum().backward()
        return loss
    for _ in range(3):
        l = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        l = step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, [l.detach().clone(), fd.grad.clone(), gd.grad.clone(), Wd.grad.clone(), bd.grad.clone()]


for wl in sys.argv[1:]:
    steps = 20 if wl in ("target", "c3") else 10
    t_r, r = run(wl, False, steps)
    t_k, k = run(wl, True, steps)
    errs = []
    for name, a, b in zip(["loss", "df", "dg", "dW", "db"], r, k):
        fro = float((a - b).norm() / (a.norm() + 1e-30))
        mx = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        errs.append(f"{name} {fro:.2e}/{mx:.2e}")
        assert torch.isfinite(b).all(), name
    print(f"{wl}: recompute {t_r:.3f} ms  kept {t_k:.3f} ms   kept vs recompute (Frobenius/max-norm): " + "  ".join(errs), flush=True)

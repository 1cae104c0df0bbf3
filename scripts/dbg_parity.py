"""Scratch: where does the 1.7e-3 full-size disagreement come from?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import myrtlespeech_b200 as M
from oracle import rnnt_oracle as O
from tests import torch_reference as R

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))

def run(B, T, U, V, H, seed=1234, gl_lin=True):
    gen = torch.Generator().manual_seed(seed)
    f = torch.randn(B, T, H, generator=gen).bfloat16()
    g = torch.randn(B, U + 1, H, generator=gen).bfloat16()
    W = ((torch.rand(V, H, generator=gen) * 2 - 1) / H ** 0.5).bfloat16()
    bias = (torch.rand(V, generator=gen) * 2 - 1) / H ** 0.5
    y = torch.randint(0, V - 1, (B, U), generator=gen, dtype=torch.int32)
    fl = torch.full((B,), T, dtype=torch.int64); yl = torch.full((B,), U, dtype=torch.int64)
    gl = torch.linspace(0.5, 1.5, B) if gl_lin else torch.ones(B)
    fd = f.float().cuda().requires_grad_(True); gd = g.float().cuda().requires_grad_(True)   # fp32 leaves: fp32 gradients
    Wd = W.float().cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), fl, yl, V - 1)
    loss.backward(gl.cuda()); torch.cuda.synchronize()
    got = dict(loss=loss.detach().cpu().numpy(), df=fd.grad.cpu().numpy(), dg=gd.grad.cpu().numpy(), dW=Wd.grad.cpu().numpy(), db=bd.grad.cpu().numpy())
    args = (f.float().numpy(), g.float().numpy(), W.float().numpy(), bias.numpy(), y.numpy(), fl.numpy(), yl.numpy(), V - 1)
    out = {}
    for name, fn in (("torch_cuda", lambda: R.rnnt_joint_loss(*args, grad_loss=gl.numpy(), faithful=True, device="cuda")),
                     ("torch_cpu", lambda: R.rnnt_joint_loss(*args, grad_loss=gl.numpy(), faithful=True, device="cpu")),
                     ("numpy", lambda: O.rnnt_joint_loss(*args, grad_loss=gl.numpy(), faithful=True))):
        if name == "numpy" and B * T * (U + 1) * max(V, H) > 3e8:
            continue
        if name == "torch_cpu" and B * T * (U + 1) * max(V, H) > 3e9:
            continue
        ref = fn()
        out[name] = {k: rel(got[k], ref[k]) for k in ("loss", "df", "dg", "dW", "db")}
    print((B, T, U, V, H), {n: {k: f"{v:.1e}" for k, v in e.items()} for n, e in out.items()}, flush=True)

run(2, 50, 20, 29, 512)
run(2, 200, 50, 29, 512)
run(2, 500, 100, 29, 512)
run(2, 500, 100, 29, 512, gl_lin=False)
run(4, 200, 100, 29, 512)
run(2, 500, 20, 29, 512)
run(1, 500, 100, 1024, 1024)

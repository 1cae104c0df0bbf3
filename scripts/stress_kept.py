"""Randomised stress of the kept-activation schedule against the recompute schedule (and the oracle on small cases):
random shapes exercise the role split (consumer / producer counts), box and chunk edges, odd tile counts, ragged lengths."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from oracle import rnnt_oracle as O          # scripts are test infrastructure: the oracle is the checker here
lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 120
worst = {}
t0 = time.time()
for case in range(n_cases):
    B = int(rng.integers(1, 7)); T = int(rng.integers(1, 140)); U = int(rng.integers(0, 70))
    V = int(rng.choice([3, 17, 29, 64, 65, 100, 255, 256, 257, 300, 512, 777, 1000, 1024, 1500, 2048, 3000, 4096]))
    H = int(rng.choice([8, 24, 64, 72, 128, 200, 256, 320, 512, 520, 768, 1024]))
    if V * H > 1 << 21 and T * U > 2000:
        T, U = max(1, T // 3), U // 3
    blank = int(rng.integers(0, V))
    f = torch.tensor(rng.normal(size=(B, T, H)), dtype=torch.float32).bfloat16().float()
    g = torch.tensor(rng.normal(size=(B, U + 1, H)), dtype=torch.float32).bfloat16().float()
    W = torch.tensor(rng.uniform(-1, 1, size=(V, H)) / np.sqrt(H), dtype=torch.float32).bfloat16().float()
    bias = torch.tensor(rng.uniform(-1, 1, size=V) / np.sqrt(H), dtype=torch.float32)
    labels = np.array([k for k in range(V) if k != blank] or [0])
    y = torch.tensor(rng.choice(labels, size=(B, max(U, 1)))[:, :U] if U > 0 else np.zeros((B, 0), dtype=np.int64), dtype=torch.int64)
    fl = rng.integers(1, T + 1, size=B); fl[0] = T
    yl = rng.integers(0, U + 1, size=B); yl[0] = U
    gl = torch.tensor(rng.uniform(-1.5, 1.5, size=B), dtype=torch.float32)      # negative upstream gradients included
    out = {}
    for keep in (True, False):
        F.set_keep_activations(keep)
        fd, gd = f.to(dev).requires_grad_(True), g.to(dev).requires_grad_(True)
        Wd, bd = W.to(dev).requires_grad_(True), bias.to(dev).requires_grad_(True)
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.to(dev), torch.tensor(fl), torch.tensor(yl), blank)
        loss.backward(gl.to(dev))
        torch.cuda.synchronize()
        out[keep] = [loss.detach().cpu(), fd.grad.cpu(), gd.grad.cpu(), Wd.grad.cpu(), bd.grad.cpu()]
    for name, a, b in zip(["loss", "df", "dg", "dW", "db"], out[True], out[False]):
        assert torch.isfinite(a).all() and torch.isfinite(b).all(), (case, name, (B, T, U, V, H))
        err = float((a - b).norm() / (b.norm() + 1e-20))
        if err > worst.get(name, (0,))[0]:
            worst[name] = (err, (B, T, U, V, H, blank))
        assert err < (1e-6 if name == "loss" else 5e-4), (case, name, err, (B, T, U, V, H, blank))
    if B * T * (U + 1) * V * H < 3e7:          # small enough for the numpy oracle: both schedules within the parity tolerance
        ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, blank, grad_loss=gl.numpy(),
                                faithful=True)
        n_oracle = globals().get("n_oracle", 0) + 1
        for keep in (True, False):
            for name, a in zip(["loss", "df", "dg", "dW", "db"], out[keep]):
                r = np.asarray(ref[name], dtype=np.float64)
                err = float(np.linalg.norm(a.double().numpy() - r) / (np.linalg.norm(r) + 1e-20))
                key = ("oracle", "kept" if keep else "recompute", name)
                if err > worst.get(key, (0,))[0]:
                    worst[key] = (err, (B, T, U, V, H, blank))
                assert err < 1e-3, (case, key, err, (B, T, U, V, H, blank))
F.set_keep_activations(True)
print(f"{n_cases} random cases in {time.time() - t0:.0f} s; worst kept-vs-recompute relative error per output:")
for k, v in worst.items():
    print(f"  {str(k):40s} {v[0]:.2e} at (B,T,U,V,H,blank) = {v[1]}")
print("cases checked against the oracle:", globals().get("n_oracle", 0))

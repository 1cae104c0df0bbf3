"""Bring-up checks on a real B200, one stage per process so a faulting kernel cannot hide later stages.

usage: python scripts/gpu_check.py <stage> [args]     (stages: lattice fwd bwd mid big greedy all)
Everything is compared with the CPU oracle (oracle/rnnt_oracle.py) or a torch fp32 reference.
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def make(seed, B, T, U, V, H, blank, ragged):
    import torch
    rng = np.random.default_rng(seed)
    f = torch.tensor(rng.normal(size=(B, T, H)), dtype=torch.float32).bfloat16()
    g = torch.tensor(rng.normal(size=(B, U + 1, H)), dtype=torch.float32).bfloat16()
    W = torch.tensor(rng.uniform(-1, 1, size=(V, H)) / np.sqrt(H), dtype=torch.float32).bfloat16()
    bias = torch.tensor(rng.uniform(-1, 1, size=V) / np.sqrt(H), dtype=torch.float32)
    labels = np.array([k for k in range(V) if k != blank])
    y = torch.tensor(rng.choice(labels, size=(B, max(U, 1)))[:, :U].reshape(B, U), dtype=torch.int32)
    fl = np.full(B, T); yl = np.full(B, U)
    if ragged and B > 1:
        fl = np.sort(rng.integers(max(1, T // 2), T + 1, size=B))[::-1].copy(); fl[0] = T
        yl = rng.integers(U // 2, U + 1, size=B); yl[0] = U
    return f, g, W, bias, y, fl.astype(np.int32), yl.astype(np.int32)


def stage_lattice():
    import torch
    from oracle import rnnt_oracle as O
    import myrtlespeech_b200 as M
    for (B, T, U, V, blank) in [(1, 2, 2, 5, 0), (3, 7, 4, 6, 5), (4, 40, 17, 9, 8), (2, 33, 0, 4, 0), (2, 300, 120, 5, 4)]:
        rng = np.random.default_rng(B * 100 + T)
        z = rng.normal(size=(B, T, U + 1, V))
        labels = np.array([k for k in range(V) if k != blank])
        y = rng.choice(labels, size=(B, max(U, 1)))[:, :U].reshape(B, U)
        fl = rng.integers(max(1, T // 2), T + 1, size=B); fl[0] = T
        yl = rng.integers(U // 2, U + 1, size=B); yl[0] = U
        loss, dz = O.rnnt_loss_from_logits(z, y, fl, yl, blank)
        zt = torch.tensor(z, dtype=torch.float32, device="cuda", requires_grad=True)
        l = M.rnnt_loss_from_logits(zt, torch.tensor(y, dtype=torch.int32), torch.tensor(fl), torch.tensor(yl), blank)
        l.sum().backward()
        print(f"lattice B={B} T={T} U={U} V={V}: loss rel {rel(l.detach().cpu().numpy(), loss):.2e} "
              f"grad rel {rel(zt.grad.cpu().numpy(), dz):.2e}", flush=True)


def _run_fused(cfg, check_stats=True, faithful=True, do_bwd=True):
    import torch
    from oracle import rnnt_oracle as O
    import myrtlespeech_b200 as M
    seed, B, T, U, V, H, blank, ragged = cfg
    f, g, W, bias, y, fl, yl = make(*cfg)
    r = O.rnnt_joint_loss(f.float().numpy(), g.float().numpy(), W.float().numpy(), bias.numpy(), y.numpy(), fl, yl,
                          blank, faithful=faithful)
    # fp32 leaves holding bf16-representable values: gradients come back in fp32 (no bf16 cast of the result)
    fd = f.float().cuda().requires_grad_(True); gd = g.float().cuda().requires_grad_(True)
    Wd = W.float().cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), torch.tensor(fl), torch.tensor(yl), blank)
    torch.cuda.synchronize()
    print(f"cfg B={B} T={T} U={U} V={V} H={H} blank={blank} ragged={ragged}")
    print(f"  loss rel {rel(loss.detach().cpu().numpy(), r['loss']):.2e}   (oracle {r['loss'][:3]}, got {loss.detach().cpu().numpy()[:3]})", flush=True)
    if do_bwd:
        loss.sum().backward()
        torch.cuda.synchronize()
        for k, t in (("df", fd), ("dg", gd), ("dW", Wd), ("db", bd)):
            got = t.grad.float().cpu().numpy()
            print(f"  {k} rel {rel(got, r[k]):.2e}  |ref|max {np.abs(r[k]).max():.3e} |got|max {np.abs(got).max():.3e}", flush=True)
    return r


def stage_fwd():
    for cfg in [(1, 1, 2, 2, 5, 8, 0, False), (2, 2, 5, 3, 6, 8, 5, False), (3, 3, 9, 4, 7, 16, 0, True),
                (4, 2, 12, 6, 29, 24, 28, True), (5, 2, 20, 9, 40, 72, 39, True)]:
        _run_fused(cfg, do_bwd=False)


def stage_bwd():
    for cfg in [(2, 2, 5, 3, 6, 8, 5, False), (3, 3, 9, 4, 7, 16, 0, True), (4, 2, 12, 6, 29, 24, 28, True),
                (5, 2, 20, 9, 40, 72, 39, True), (6, 2, 37, 11, 300, 128, 299, True)]:
        _run_fused(cfg)


def stage_mid():
    # BASELINE config C1 (oracle-sized) and a V=H=1024 case, multi-slab
    for cfg in [(7, 4, 200, 50, 29, 512, 28, True), (8, 2, 60, 20, 1024, 1024, 1023, True)]:
        t = time.time()
        _run_fused(cfg)
        print(f"  ({time.time() - t:.1f}s incl. oracle)", flush=True)


def stage_big():
    """Target shape, timing only + size-independent properties (sum_v dz = 0 => db sums to ~0; dW finite)."""
    import torch
    import myrtlespeech_b200 as M
    B, T, U, V, H = 32, 500, 100, 1024, 1024
    f, g, W, bias, y, fl, yl = make(9, B, T, U, V, H, V - 1, False)
    fd = f.cuda().requires_grad_(True); gd = g.cuda().requires_grad_(True)
    Wd = W.cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
    yd = y.cuda(); flt = torch.tensor(fl); ylt = torch.tensor(yl)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, flt, ylt, V - 1)
        e1.record()
        loss.sum().backward()
        e2.record()
        torch.cuda.synchronize()
        print(f"big it{it}: fwd {e0.elapsed_time(e1):.2f} ms  bwd {e1.elapsed_time(e2):.2f} ms  wall {1e3 * (time.time() - t0):.1f} ms  "
              f"loss[0]={loss[0].item():.3f} db.sum={bd.grad.sum().item():.3e} |dW|max={Wd.grad.abs().max().item():.3e}", flush=True)
        fd.grad = gd.grad = Wd.grad = bd.grad = None


def stage_greedy():
    import torch
    import myrtlespeech_b200 as M
    B, T, V, H = 16, 30, 1024, 1024
    f, g, W, bias, y, fl, yl = make(11, B, T, 3, V, H, V - 1, False)
    fd, Wd, bd = f.cuda(), W.cuda(), bias.cuda()
    gd = g[:, 0].contiguous().cuda()
    t_idx = torch.arange(B, dtype=torch.int32, device="cuda") % T
    t_idx[3] = -1
    out = M.greedy_joint_argmax(fd, gd, Wd, bd, t_idx)
    h = torch.tanh(fd[torch.arange(B), t_idx.clamp(min=0).long()].float() + gd.float()).bfloat16().float()
    z = h @ Wd.float().T + bd
    ref = z.argmax(-1).int(); ref[3] = -1
    print("greedy match:", bool((out == ref).all().item()), out[:8].tolist(), ref[:8].tolist(), flush=True)


STAGES = dict(lattice=stage_lattice, fwd=stage_fwd, bwd=stage_bwd, mid=stage_mid,
              big=stage_big, greedy=stage_greedy)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        names = sys.argv[2:] or ["lattice", "fwd", "bwd", "mid", "greedy", "big"]
        for name in names:
            print(f"===== stage {name} =====", flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240)
                print(f"===== stage {name} exit {r.returncode} =====", flush=True)
            except subprocess.TimeoutExpired:
                print(f"===== stage {name} TIMEOUT =====", flush=True)
    else:
        STAGES[which]()

"""Fraction of lattice tiles the backward pass walks (tiles with non-zero arc occupancy) at a bench workload."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from bench import WORKLOADS, synth
lib = _lib.load()
for wl in (sys.argv[1:] or ["target"]):
    B, T, U, V, H, _ = WORKLOADS[wl]
    f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
    fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
    for t in (fd, gd, Wd, bd): t.requires_grad_(True)
    for eps in (None, -60, -30):
        lib.rnnt_debug_set(b"prune_log2_eps", -100000 if eps is None else eps)
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        loss.sum().backward(); torch.cuda.synchronize()
        ws = next(iter(F._ws_pool.values()))
        out = (ctypes.c_int * 2)()
        _lib.check(lib.rnnt_debug_read_active_tiles(ws.data_ptr(), B, T, U, V, H, out))
        print(f"{wl}: threshold {'exact zero' if eps is None else '2^%d' % eps}: {out[0]} of {out[1]} tiles active ({100.0 * out[0] / out[1]:.1f} %)", flush=True)
        fd.grad = gd.grad = Wd.grad = bd.grad = None
    lib.rnnt_debug_set(b"prune_log2_eps", -100000)

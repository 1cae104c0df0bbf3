"""Forward kernel with kept activations: what do the logit stores cost?  (gemm_dbg 64 = stage but do not issue the TMA stores)"""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from bench import WORKLOADS, synth
lib = _lib.load()
dev = torch.device("cuda", 0)
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, dev)
fd, gd, Wd, bd, yd = (x.to(dev) for x in (f, g, W, bias, y))
lib.rnnt_debug_set(b"time_kernels", 1)
def fwd_ms(n=15):
    for _ in range(4):
        with torch.no_grad():
            M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 9)(); cnt = (ctypes.c_longlong * 9)()
    lib.rnnt_debug_kernel_times(ms, cnt, 9)
    for _ in range(n):
        with torch.no_grad():
            M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    torch.cuda.synchronize()
    lib.rnnt_debug_kernel_times(ms, cnt, 9)
    return ms[1] / max(cnt[1], 1)
for rnd in range(2):
    for name, keep, dbg in (("recompute (nothing kept)", False, 0), ("kept", True, 0), ("kept, logit stores not issued", True, 64), ("kept, no logit staging at all", True, 128)):
        F.set_keep_activations(keep)
        lib.rnnt_debug_set(b"gemm_dbg", dbg)
        print(f"{name:34s} forward {fwd_ms():.3f} ms", flush=True)
lib.rnnt_debug_set(b"gemm_dbg", 0)

"""RNNT_PROFILE build: where the mega-kernel's dh epilogue (warp 0 of every producer CTA) spends its cycles, per chunk."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
WL = sys.argv[1] if len(sys.argv) > 1 else "target"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 50
B, T, U, V, H, _ = WORKLOADS[WL]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
for t in (fd, gd, Wd, bd): t.requires_grad_(True)
buf = (ctypes.c_ulonglong * (160 * 8))()
for it in range(3):
    lib.rnnt_debug_read_prof3(buf, 160 * 8, 1)
    lib.rnnt_debug_set(b"gemm_dbg", 4)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    loss.sum().backward(); torch.cuda.synchronize()
lib.rnnt_debug_read_prof3(buf, 160 * 8, 0)
a = np.array(buf[:], dtype=np.float64).reshape(160, 8)[: 2 * P]
n = a[:, 7:8]
names = ["wait tfull", "tmem ld", "math (h, 1-h^2)", "butterflies + red/sts", "barrier 1", "final dg reduce", "barrier 2"]
per = np.median(a[:, :7] / n, axis=0)
print(f"{WL}: dh epilogue of warp 0, cycles per chunk (median over {2 * P} producer CTAs):")
for nm, v in zip(names, per):
    print(f"  {nm:24s} {v:8.0f}")
print(f"  {'sum':24s} {per.sum():8.0f}")

"""RNNT_PROFILE build: issue cycles per chunk of the producers' MMA warp, dz pass vs dh pass."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
WL = os.environ.get("WL", "target")
B, T, U, V, H, _ = WORKLOADS[WL]
P = int(os.environ.get("P", "50"))
CHV = (V + 255) // 256
CHH = (H + 255) // 256
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
for t in (fd, gd, Wd, bd): t.requires_grad_(True)
lib.rnnt_debug_set(b"gemm_dbg", 4 | (int(sys.argv[1]) if len(sys.argv) > 1 else 0))
for it in range(3):
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    loss.sum().backward(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (2 * 160 * 8))()
lib.rnnt_debug_read_prof(buf, 2 * 160 * 8)
a = np.array(buf[:], dtype=np.int64).reshape(2, 160, 8)
lead = a[1][0:2 * P:2]
n_ptiles = (B * ((T + 15) // 16) * ((U + 1 + 7) // 8) + 1) // 2
tiles = (n_ptiles + P - 1) // P
print("per-chunk issue cycles (median over producer leaders; ideal 8192):")
print("  dz pass mean %.0f  max %d" % (np.median(lead[:, 0]) / (tiles * CHV), np.median(lead[:, 2])))
print("  dh pass mean %.0f  max %d" % (np.median(lead[:, 1]) / (tiles * CHH), np.median(lead[:, 3])))
print("  first chunk of a pass mean %.0f" % (np.median(lead[:, 4]) / (tiles * 2)))

for name, col in (("epilogue set 0 (cols 0-127)", 5), ("epilogue set 1 (cols 128-255)", 6)):
    v = a[1][0:2 * P, col].astype(np.uint64)
    hold_dz = ((v >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.float64) * 1024 / (tiles * CHV)
    tot_dz = ((v >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.float64) * 1024 / (tiles * CHV)
    hold_dh = ((v >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.float64) * 1024 / (tiles * CHH)
    tot_dh = (v & np.uint64(0xFFFF)).astype(np.float64) * 1024 / (tiles * CHH)
    print(f"{name}: per chunk cycles (median over CTAs)  dz: TMEM hold {np.median(hold_dz):.0f}, total {np.median(tot_dz):.0f};"
          f"  dh: TMEM hold {np.median(hold_dh):.0f}, total {np.median(tot_dh):.0f}")

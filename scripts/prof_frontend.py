"""Front-end warps of the mega-kernel's producers (RNNT_PROFILE build): cycles waiting for a free slot, in hgen, in the dz pass."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib, functional as F
from bench import WORKLOADS, synth
lib = _lib.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "target"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 42
F.set_keep_activations(len(sys.argv) <= 3 or sys.argv[3] != "0")
B, T, U, V, H, _ = WORKLOADS[wl]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
fd.requires_grad_(True); gd.requires_grad_(True); Wd.requires_grad_(True); bd.requires_grad_(True)
lib.rnnt_debug_set(b"gemm_dbg", 4)
zb = (ctypes.c_ulonglong * (2 * 160 * 8))()
for it in range(4):
    if it == 3:
        lib.rnnt_debug_read_prof3(zb, 2 * 160 * 8, 1)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.sum().backward(); e1.record(); torch.cuda.synchronize()
    print("backward ms", e0.elapsed_time(e1))
buf = (ctypes.c_ulonglong * (2 * 160 * 8))()
lib.rnnt_debug_read_prof(buf, 2 * 160 * 8)
a = np.array(buf[:], dtype=np.uint64).reshape(2, 160, 8)
v = a[1, : 2 * P, 7]
wait, hgen, dz = (v >> np.uint64(40)).astype(np.int64) << 10, ((v >> np.uint64(20)) & np.uint64(0xFFFFF)).astype(np.int64) << 10, (v & np.uint64(0xFFFFF)).astype(np.int64) << 10
total = a[0, : 2 * P : 2, 0].astype(np.int64)
print(f"{wl} P={P}: MMA loop cycles median {int(np.median(total))}; front-end median cycles: slot wait {int(np.median(wait))}  hgen {int(np.median(hgen))}  dz {int(np.median(dz))}")
lib.rnnt_debug_read_prof3(zb, 2 * 160 * 8, 0)
z = np.array(zb[:], dtype=np.float64).reshape(2, 160, 8)[1, : 2 * P]
nb = np.maximum(z[:, 7], 1)
names = ["wait zfull", "transform in place", "fence + arrive"]
print("dz pass of thread 0, cycles per box (median over producer CTAs):")
for i, nm in enumerate(names):
    print(f"  {nm:34s} {np.median(z[:, i] / nb):8.0f}")

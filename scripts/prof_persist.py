"""Wait-cycle counters of the persistent forward kernel (gemm_dbg=4): which role starves which."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = f.cuda(), g.cuda(), W.cuda(), bias.cuda(), y.cuda()
lib.rnnt_debug_set(b"gemm_dbg", 4)
for it in range(3):
    with torch.no_grad():
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (160 * 8))()
lib.rnnt_debug_read_prof(buf, 160 * 8)
a = np.array(buf[:], dtype=np.int64).reshape(160, 8)[:148]
lead = a[0::2]
names = ["mma_loop_cyc", "mma_wait_full", "mma_wait_tempty", "mma_loop_ns", "tma_wait_hfull", "tma_wait_empty", "hgen_busy", "epi_busy"]
for i, nm in enumerate(names):
    col = lead[:, i] if i < 4 else a[:, i]
    print(f"{nm:16s} min {col.min():10d} median {int(np.median(col)):10d} max {col.max():10d}")
print("effective SM clock in MMA loop: %.3f GHz" % (np.median(lead[:, 0] / np.maximum(lead[:, 3], 1))))

"""Wait-cycle counters of the backward mega-kernel (gemm_dbg=4), producers and consumers separately."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "target"]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if len(sys.argv) > 3:
    lib.rnnt_debug_set(b"ring_slots", int(sys.argv[3]))
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
fd.requires_grad_(True); gd.requires_grad_(True); Wd.requires_grad_(True); bd.requires_grad_(True)
lib.rnnt_debug_set(b"gemm_dbg", 4)
for it in range(3):
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.sum().backward(); e1.record(); torch.cuda.synchronize()
    print("backward ms", e0.elapsed_time(e1))
buf = (ctypes.c_ulonglong * (160 * 8))()
lib.rnnt_debug_read_prof(buf, 160 * 8)
a = np.array(buf[:], dtype=np.int64).reshape(160, 8)[:148]
def show(title, rows, names, cols):
    print(title)
    for nm, i in zip(names, cols):
        col = rows[:, i]
        print(f"  {nm:18s} min {col.min():10d} median {int(np.median(col)):10d} max {col.max():10d}")
prod_lead, cons_lead = a[0:2 * P:2], a[2 * P::2]
show("producer leaders (MMA warp)", prod_lead, ["mma_loop_cyc", "wait_full", "wait_tempty", "loop_ns"], [0, 1, 2, 3])
show("producer CTAs (TMA warp)", a[:2 * P], ["wait_hfull", "wait_empty", "wait_dzr"], [4, 5, 6])
show("consumer leaders (MMA warp)", cons_lead, ["mma_loop_cyc", "wait_full", "loop_ns"], [0, 1, 3])
show("consumer CTAs (TMA warp)", a[2 * P:], ["wait_ready", "wait_empty"], [4, 5])
print("effective SM clock: producers %.3f GHz, consumers %.3f GHz" % (
    np.median(prod_lead[:, 0] / np.maximum(prod_lead[:, 3], 1)), np.median(cons_lead[:, 0] / np.maximum(cons_lead[:, 3], 1))))

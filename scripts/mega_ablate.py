"""Backward mega-kernel with parts of the producer epilogue switched off (results are garbage): what bounds it?"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS["target"]
P = 50
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
for t in (fd, gd, Wd, bd):
    t.requires_grad_(True)
cfgs = [(4, "full"), (4 | 32, "no db sums"), (4 | 64, "no dh reduce"), (4 | 64 | 128, "no dh tile+reduce"),
        (4 | 32 | 64 | 128, "no db, no dh tile+reduce")]
if len(sys.argv) > 1:
    cfgs = [(4 | int(a), f"dbg bits {a}") for a in sys.argv[1:]]
for dbg, label in cfgs:
    lib.rnnt_debug_set(b"gemm_dbg", dbg)
    ms = []
    for it in range(3):
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss.sum().backward(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    buf = (ctypes.c_ulonglong * (160 * 8))()
    lib.rnnt_debug_read_prof(buf, 160 * 8)
    a = np.array(buf[:], dtype=np.int64).reshape(160, 8)[:148]
    pl, cl = a[0:2 * P:2], a[2 * P::2]
    print(f"{label:28s} bwd {ms[-1]:6.2f} ms | prod loop {int(np.median(pl[:,0]))/1e6:6.2f}M wait_full {int(np.median(pl[:,1]))/1e6:5.2f}M "
          f"wait_tempty {int(np.median(pl[:,2]))/1e6:5.2f}M clk {np.median(pl[:,0]/np.maximum(pl[:,3],1)):.2f} GHz | cons loop "
          f"{int(np.median(cl[:,0]))/1e6:6.2f}M wait_full {int(np.median(cl[:,1]))/1e6:5.2f}M wait_ready {int(np.median(a[2*P:,4]))/1e6:5.2f}M", flush=True)

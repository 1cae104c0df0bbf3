"""Persistent forward kernel with the A and/or B TMA loads skipped: how much of its time is L2 feed?"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = f.cuda(), g.cuda(), W.cuda(), bias.cuda(), y.cuda()
names = ["mma_loop_cyc", "mma_wait_full", "mma_wait_tempty", "mma_loop_ns"]
for dbg, label in ((4, "all loads"), (4 | 8, "no A loads"), (4 | 16, "no B loads"), (4 | 256, "no loads after ring fill")):
    lib.rnnt_debug_set(b"gemm_dbg", dbg)
    for it in range(3):
        with torch.no_grad():
            loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (160 * 8))()
    lib.rnnt_debug_read_prof(buf, 160 * 8)
    a = np.array(buf[:], dtype=np.int64).reshape(160, 8)[:148][0::2]
    print(f"{label:12s}: " + "  ".join(f"{n}={int(np.median(a[:, i]))}" for i, n in enumerate(names)) +
          "  clk=%.3f GHz" % np.median(a[:, 0] / np.maximum(a[:, 3], 1)), flush=True)

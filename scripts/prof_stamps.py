"""In-kernel %globaltimer stamps of the last joint_fwd launch (gemm_dbg=4): where a launch's time goes."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS["target"]
B = 37  # 416 tiles per utterance * 37 = 104 full slabs of 148 tiles, so the last launch is a full one
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = f.cuda(), g.cuda(), W.cuda(), bias.cuda(), y.cuda()
# only 1 slab + so the last launch is a full one: use the first 148 tiles worth of utterances -> B=... simply run full and read last (partial) launch
for dbg in (4, 4 | 1 | 2):
    lib.rnnt_debug_set(b"gemm_dbg", dbg)
    for it in range(2):
        with torch.no_grad():
            loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (160 * 8))()
    n = lib.rnnt_debug_read_prof(buf, 160 * 8)
    a = np.array(buf[:], dtype=np.int64).reshape(160, 8)[:148]
    a = a[a[:, 0] > 0]
    t0 = a[:, 0].min()
    names = ["entry", "setup_done", "first_mma", "last_mma_issued", "epi_chunk0_ready", "epi_lastchunk_ready", "epi_done", "exit"]
    print(f"dbg={dbg}: {len(a)} CTAs; ns relative to the earliest CTA entry (min / median / max)")
    for i, nm in enumerate(names):
        col = a[:, i][a[:, i] > 0] - t0
        if len(col):
            print(f"  {nm:22s} {col.min():8d} {int(np.median(col)):8d} {col.max():8d}   (n={len(col)})")

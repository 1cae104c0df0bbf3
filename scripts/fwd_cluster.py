"""Forward kernel time (CUDA events) with clusters of 2 (148 CTAs) and 4 (W multicast, co-resident grid only)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = f.cuda(), g.cuda(), W.cuda(), bias.cuda(), y.cuda()
for cs in (2, 4, 2, 4):
    lib.rnnt_debug_set(b"cluster", cs)
    ms = []
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            e0.record(); loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"cluster {cs}: forward call (fwd kernel + lattice) ms: " + " ".join(f"{m:.3f}" for m in ms[2:]), flush=True)

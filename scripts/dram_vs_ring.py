"""Two fwd+bwd passes at the target shape with a given number of dW-consumer K-groups (ring footprint = P x NS x 1 MB,
P = 74 - 8 KG) and ring slots: run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M  # noqa: E402
from myrtlespeech_b200 import _lib  # noqa: E402
from bench import WORKLOADS, synth  # noqa: E402

lib = _lib.load()
lib.rnnt_debug_set(b"mega_cooperative", 0)
kg = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.rnnt_debug_set(b"mega_kg", kg)
if len(sys.argv) > 2:
    lib.rnnt_debug_set(b"ring_slots", int(sys.argv[2]))
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd = f.cuda().requires_grad_(True); gd = g.cuda().requires_grad_(True)
Wd = W.cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
for it in range(2):
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), fl, yl, V - 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.sum().backward(); e1.record()
    torch.cuda.synchronize()
    fd.grad = gd.grad = Wd.grad = bd.grad = None
print("ok kg", kg, "bwd ms", e0.elapsed_time(e1), float(loss[0]))

"""Times the forward pass (kernel-class event timing) under the bring-up switches of the GEMM kernel."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from bench import WORKLOADS, synth, KCLASSES
lib = _lib.load()
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = f.cuda(), g.cuda(), W.cuda(), bias.cuda(), y.cuda()
for dbg in (0, 1, 2, 3, 0):
    lib.rnnt_debug_set(b"gemm_dbg", dbg)
    for it in range(2):
        lib.rnnt_debug_set(b"time_kernels", it)
        with torch.no_grad():
            loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
        torch.cuda.synchronize()
    kms = (ctypes.c_double * 16)(); kn = (ctypes.c_longlong * 16)()
    lib.rnnt_debug_kernel_times(kms, kn, 16)
    print(f"dbg={dbg}: " + "  ".join(f"{KCLASSES[i]}={kms[i]:.3f}ms/{kn[i]}" for i in range(len(KCLASSES)) if kn[i]), flush=True)

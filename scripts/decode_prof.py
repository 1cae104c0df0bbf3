"""Per-stage cycle breakdown of the cluster decode's epilogue (cluster 0, rank 0) at configs[4]."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myrtlespeech_b200 import _lib
from myrtlespeech_b200.model import RNNTJoint
from myrtlespeech_b200.model.rnn_t import RNNT, RNNTPredictionNet
from myrtlespeech_b200.post_process import RNNTGreedyDecoder
B, T, V, H, S = 128, 500, 1024, 1024, 4
E, HP = int(os.environ.get("PRED_E", 256)), int(os.environ.get("PRED_H", 512))
torch.manual_seed(0)
model = RNNT(torch.nn.Identity(), RNNTPredictionNet(V, E, HP, 1, H), RNNTJoint(H, V)).cuda()
dec = RNNTGreedyDecoder(V - 1, model, max_symbols_per_step=S)
f = torch.randn(B, T, H, device="cuda").bfloat16()
lens = torch.full((B,), T, dtype=torch.int32)
lib = _lib.load()
for c in [int(x) for x in os.environ.get("CLUSTERS", "8").split(",")]:
    lib.rnnt_debug_set(b"decode_cluster", c)
    lib.rnnt_debug_set(b"decode_prof", int(os.environ.get("DEC_PROF", 1)))
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = dec(f, lens); e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    lib.rnnt_debug_decode_prof(buf, 16)
    n = max(1, buf[9])
    names = ["gather issue", "wait L acc", "L epilogue+push", "wait P acc", "P epilogue+push", "wait J acc", "J epilogue+send",
             "wait keys", "bookkeeping"]
    tot = sum(buf[i] for i in range(9))
    print(f"cluster size {c}: {e0.elapsed_time(e1):.1f} ms, {n} steps, {tot / n:.0f} cycles/step")
    for i, nm in enumerate(names):
        print(f"  {nm:18s} {buf[i] / n:8.0f} cycles/step  {100.0 * buf[i] / tot:5.1f} %")
    print(f"  producer: waits for a free ring slot {buf[10] / n:.0f}, for the step decision {buf[11] / n:.0f} cycles/step")
    print(f"  MMA warp: waits for weights {buf[12] / n:.0f}, for exchanged activations {buf[13] / n:.0f} cycles/step")

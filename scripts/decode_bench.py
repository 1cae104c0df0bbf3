"""BASELINE.json configs[4]: greedy decode over the fused joint, B=128 T=500 V=H=1024 max-symbols-per-step=4.
Random-init LSTM prediction network (there is no checkpoint); reports utterances/s and joint steps/s."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myrtlespeech_b200.model import RNNTJoint
from myrtlespeech_b200.model.rnn_t import RNNT, RNNTPredictionNet
from myrtlespeech_b200.post_process import RNNTGreedyDecoder
B, T, V, H, S = 128, 500, 1024, 1024, 4
torch.manual_seed(0)
joint = RNNTJoint(H, V)
pred = RNNTPredictionNet(V, 256, 512, 1, H)
model = RNNT(torch.nn.Identity(), pred, joint).cuda()
dec = RNNTGreedyDecoder(V - 1, model, max_symbols_per_step=S)
f = torch.randn(B, T, H, device="cuda").bfloat16()
lens = torch.full((B,), T, dtype=torch.int32)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = dec(f, lens)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n_sym = sum(len(o) for o in out)
    print(f"decode pass {it}: {dt*1e3:.1f} ms, {B/dt:.1f} utt/s, {n_sym} symbols emitted ({n_sym/B/T:.2f} per frame)", flush=True)

"""One greedy decode of configs[4] through the one-launch path (target for `ncu -k regex:greedy_decode`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myrtlespeech_b200.model import RNNTJoint
from myrtlespeech_b200.model.rnn_t import RNNT, RNNTPredictionNet
from myrtlespeech_b200.post_process import RNNTGreedyDecoder
B, T, V, H, S = 128, int(os.environ.get("DEC_T", 500)), 1024, 1024, 4
torch.manual_seed(0)
model = RNNT(torch.nn.Identity(), RNNTPredictionNet(V, 256, 512, 1, H), RNNTJoint(H, V)).cuda()
dec = RNNTGreedyDecoder(V - 1, model, max_symbols_per_step=S)
f = torch.randn(B, T, H, device="cuda").bfloat16()
out = dec(f, torch.full((B,), T, dtype=torch.int32))
torch.cuda.synchronize()
print("symbols", sum(len(o) for o in out))

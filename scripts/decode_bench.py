"""BASELINE.json configs[4]: greedy decode over the fused joint, B=128 T=500 V=H=1024 max-symbols-per-step=4.
Random-init LSTM prediction network (there is no checkpoint); reports utterances/s for the one-launch decode
(rnnt_greedy_decode_lstm) and for the per-step CUDA-graph loop, CUDA-event timed around the decoder call."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myrtlespeech_b200.model import RNNTJoint
from myrtlespeech_b200.model.rnn_t import RNNT, RNNTPredictionNet
from myrtlespeech_b200.post_process import RNNTGreedyDecoder
B, T, V, H, S = int(os.environ.get("DEC_B", 128)), 500, 1024, 1024, 4
E, HP = int(os.environ.get("PRED_E", 256)), int(os.environ.get("PRED_H", 512))
BLANK_BIAS = float(os.environ.get("BLANK_BIAS", 0.0))
LAYERS = int(os.environ.get("PRED_LAYERS", 1))
torch.manual_seed(0)
joint = RNNTJoint(H, V)
pred = RNNTPredictionNet(V, E, HP, LAYERS, H)
with torch.no_grad():
    joint.fc.bias[V - 1] += BLANK_BIAS
model = RNNT(torch.nn.Identity(), pred, joint).cuda()
dec = RNNTGreedyDecoder(V - 1, model, max_symbols_per_step=S)
f = torch.randn(B, T, H, device="cuda").bfloat16()
lens = torch.full((B,), T, dtype=torch.int32)
outs = {}
from myrtlespeech_b200 import _lib
for _k in ("decode_l_late", "decode_resident", "decode_cluster"):
    if _k.upper() in os.environ:
        _lib.load().rnnt_debug_set(_k.encode(), int(os.environ[_k.upper()]))
for fused in (("cluster", "gridsync", False) if LAYERS == 1 else ("cluster", False)):
    dec.USE_FUSED_LOOP = bool(fused)
    _lib.load().rnnt_debug_set(b"decode_variant", 1 if fused == "cluster" else 0)
    for it in range(int(os.environ.get("PASSES", 3))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
        out = dec(f, lens)
        e1.record(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        n_sym = sum(len(o) for o in out)
        steps = max(len(o) for o in out) + T
        print(f"{('one-launch/' + fused) if fused else 'graph-step'} pass {it}: wall {dt*1e3:.1f} ms (device {e0.elapsed_time(e1):.1f} ms), "
              f"{B/dt:.1f} utt/s, {n_sym} symbols ({n_sym/B/T:.2f} per frame), <= {steps} steps "
              f"-> {e0.elapsed_time(e1)*1e3/steps:.1f} us/step", flush=True)
    outs[fused] = out
if "gridsync" in outs:
    same = sum(a == b for a, b in zip(outs["cluster"], outs["gridsync"]))
    print(f"transcripts identical between the two one-launch schedules: {same}/{B}")
same = sum(a == b for a, b in zip(outs["cluster"], outs[False]))
print(f"transcripts identical between one-launch (bf16 LSTM operands) and graph-step (cuDNN fp32 LSTM): {same}/{B}")

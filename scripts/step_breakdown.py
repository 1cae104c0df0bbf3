"""Where does a step go?  Event-timed segments of the sustained step loop: library forward, torch glue, backward."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import myrtlespeech_b200 as M
from bench import WORKLOADS, synth
B, T, U, V, H, _ = WORKLOADS["target"]
f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234, "cuda")
fd, gd, Wd, bd, yd = (x.cuda() for x in (f, g, W, bias, y))
for t in (fd, gd, Wd, bd):
    t.requires_grad_(True)
def step(evs=None):
    if evs: evs[0].record()
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, V - 1)
    if evs: evs[1].record()
    tot = loss.sum()
    if evs: evs[2].record()
    tot.backward()
    if evs: evs[3].record()
    fd.grad = gd.grad = Wd.grad = bd.grad = None
for _ in range(10):
    step()
torch.cuda.synchronize()
N = 40
all_evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(N)]
t0 = time.perf_counter()
for i in range(N):
    step(all_evs[i])
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
seg = [sum(e[k].elapsed_time(e[k + 1]) for e in all_evs) / N for k in range(3)]
gap = sum(all_evs[i][3].elapsed_time(all_evs[i + 1][0]) for i in range(N - 1)) / (N - 1)
print(f"host enqueue time per step {t_host / N * 1e3:.2f} ms; wall per step {t_all / N * 1e3:.2f} ms")
print(f"GPU segments per step: forward call {seg[0]:.3f} ms | loss.sum {seg[1]:.3f} ms | backward {seg[2]:.3f} ms | between steps {gap:.3f} ms")

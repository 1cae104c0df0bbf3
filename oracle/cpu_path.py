"""CPU stand-in for "the reference's own implementation" of the path, used only as a timed baseline.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/rnnt_oracle.py header).  The reference snapshot has no
RNN-T code; what a myrtlespeech user would run on the CPU is stock-PyTorch eager ops plus
a library transducer loss, exactly as its CTC path is stock ``LogSoftmax`` + ``torch.nn.CTCLoss``
(``loss/ctc_loss.py:45-48,95-101``).  So the baseline is:

    logits = Linear(H, V)(tanh(f.unsqueeze(2) + g.unsqueeze(1)))      # model/fully_connected.py:164 analog
    loss   = torchaudio.functional.rnnt_loss(logits, y, f_lens, y_lens, blank, reduction="sum")
    loss.backward()

in fp32 on the host cores.  It materialises the (B,T,U+1,V) tensor, as the reference
convention would.
"""
import os
import time
from typing import Dict, Optional

import torch


def step(f, g, W, bias, y, f_lens, y_lens, blank: int) -> Dict[str, torch.Tensor]:
    import torchaudio

    f = f.detach().float().requires_grad_(True)
    g = g.detach().float().requires_grad_(True)
    W = W.detach().float().requires_grad_(True)
    bias = bias.detach().float().requires_grad_(True)
    logits = torch.nn.functional.linear(torch.tanh(f.unsqueeze(2) + g.unsqueeze(1)), W, bias)
    loss = torchaudio.functional.rnnt_loss(
        logits, y.int(), f_lens.int(), y_lens.int(), blank=blank, reduction="sum", fused_log_softmax=True
    )
    loss.backward()
    return dict(loss=loss.detach(), df=f.grad, dg=g.grad, dW=W.grad, db=bias.grad)


def time_steps(B: int, T: int, U: int, V: int, H: int, steps: int, warmup: int, seed: int = 1234,
               threads: Optional[int] = None) -> Dict:
    """Times `steps` fwd+bwd passes on synthetic inputs of the given shape; returns utt/s and metadata.

    ``threads`` (default: every host core, ``os.cpu_count()``) is set explicitly with ``torch.set_num_threads`` so that
    the figure does not depend on ``OMP_NUM_THREADS`` -- ``torchrun`` sets it to 1 for its workers."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(seed)
    f = torch.randn(B, T, H, generator=gen)
    g = torch.randn(B, U + 1, H, generator=gen)
    W = (torch.rand(V, H, generator=gen) * 2 - 1) / H ** 0.5
    bias = (torch.rand(V, generator=gen) * 2 - 1) / H ** 0.5
    y = torch.randint(0, V - 1, (B, U), generator=gen, dtype=torch.int32)
    fl = torch.full((B,), T, dtype=torch.int32)
    yl = torch.full((B,), U, dtype=torch.int32)
    for _ in range(warmup):
        step(f, g, W, bias, y, fl, yl, V - 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(f, g, W, bias, y, fl, yl, V - 1)
    dt = time.perf_counter() - t0
    return dict(utt_per_s=B * steps / dt, ms_per_step=1e3 * dt / steps, cores=torch.get_num_threads(),
                sample=f"{steps} fwd+bwd steps of B={B} T={T} U={U} V={V} H={H} fp32 (torch eager joint + torchaudio rnnt_loss)")

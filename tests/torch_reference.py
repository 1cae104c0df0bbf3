"""Chunked fp64 torch restatement of ``oracle/rnnt_oracle.py::rnnt_joint_loss`` for problem sizes the numpy oracle
cannot hold (BASELINE.json configs[1], configs[2] and the target shape: 1.6-1.9 M lattice rows).

TEST INFRASTRUCTURE ONLY -- a checker, like ``oracle/``; nothing under ``myrtlespeech_b200/`` imports it.

It walks the batch utterance by utterance (one utterance's ``(T, U+1, V)`` logits are 0.4 GB in fp64, the whole
``B x T x (U+1) x V`` tensor never exists), evaluates every product in fp64 on whatever device it is given (the GPU
in the ``-m gpu`` tests, the CPU in ``tests/test_torch_reference.py``) and rounds exactly where the numpy oracle's
``faithful=True`` mode rounds -- ``h`` to bf16 after an fp32 add, ``dz`` to bf16 -- so that it is the *same function*
as the pinned oracle, only scheduled differently.  ``tests/test_torch_reference.py`` pins it against the numpy
oracle (loss, df, dg, dW, db to 1e-10) on small ragged cases before any GPU test trusts it.

The formulas are SURVEY.md Appendix A (Graves 2012); the alpha/beta recurrences run as anti-diagonal wavefronts
(every cell of a diagonal depends only on the previous diagonal), which is only a reordering of the oracle's loops.
"""
from typing import Dict, Optional

import numpy as np
import torch

NEG = -1.0e30


def _bf16(x: torch.Tensor) -> torch.Tensor:
    """fp64 -> fp32 -> bf16 (round to nearest even) -> fp64, the oracle's ``bf16_round``."""
    return x.float().bfloat16().double()


def _lattice(lpb: torch.Tensor, lpl: torch.Tensor, T: int, U: int):
    """alpha, beta (T, U+1) fp64 and ln P for one utterance; ``lpl[:, U]`` is ignored."""
    dev = lpb.device
    U1 = U + 1
    alpha = torch.full((T, U1), NEG, dtype=torch.float64, device=dev)
    beta = torch.full((T, U1), NEG, dtype=torch.float64, device=dev)
    alpha[0, 0] = 0.0
    for d in range(1, T + U):
        u = torch.arange(max(0, d - (T - 1)), min(U, d) + 1, device=dev)
        t = d - u
        from_blank = torch.where(t > 0, alpha[(t - 1).clamp(min=0), u] + lpb[(t - 1).clamp(min=0), u],
                                 torch.full_like(u, NEG, dtype=torch.float64))
        from_label = torch.where(u > 0, alpha[t, (u - 1).clamp(min=0)] + lpl[t, (u - 1).clamp(min=0)],
                                 torch.full_like(u, NEG, dtype=torch.float64))
        alpha[t, u] = torch.logaddexp(from_blank, from_label)
    beta[T - 1, U] = lpb[T - 1, U]
    for d in range(T + U - 2, -1, -1):
        u = torch.arange(max(0, d - (T - 1)), min(U, d) + 1, device=dev)
        t = d - u
        by_blank = torch.where(t < T - 1, beta[(t + 1).clamp(max=T - 1), u] + lpb[t, u],
                               torch.full_like(u, NEG, dtype=torch.float64))
        by_label = torch.where(u < U, beta[t, (u + 1).clamp(max=U)] + lpl[t, u],
                               torch.full_like(u, NEG, dtype=torch.float64))
        beta[t, u] = torch.logaddexp(by_blank, by_label)
    lnp = alpha[T - 1, U] + lpb[T - 1, U]
    return alpha, beta, lnp


def rnnt_joint_loss(f, g, W, bias, y, f_lens, y_lens, blank: int, grad_loss: Optional[np.ndarray] = None,
                    faithful: bool = True, device: str = "cpu") -> Dict[str, np.ndarray]:
    """Same contract as ``oracle.rnnt_oracle.rnnt_joint_loss``: dict(loss (B,), df, dg, dW, db) as fp64 numpy arrays;
    ``grad_loss`` defaults to ones (reduction "sum")."""
    dev = torch.device(device)
    as64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)  # noqa: E731
    f, g, W = as64(f), as64(g), as64(W)
    B, Tm, H = f.shape
    U1m = g.shape[1]
    V = W.shape[0]
    bias_ = torch.zeros(V, dtype=torch.float64, device=dev) if bias is None else as64(bias)
    gl = np.ones(B) if grad_loss is None else np.asarray(grad_loss, dtype=np.float64)
    y = np.asarray(y)
    loss = np.zeros(B)
    df = torch.zeros(B, Tm, H, dtype=torch.float64, device=dev)
    dg = torch.zeros(B, U1m, H, dtype=torch.float64, device=dev)
    dW = torch.zeros(V, H, dtype=torch.float64, device=dev)
    db = torch.zeros(V, dtype=torch.float64, device=dev)
    for b in range(B):
        T, U = int(f_lens[b]), int(y_lens[b])
        U1 = U + 1
        x = f[b, :T, None, :] + g[b, None, :U1, :]
        if faithful:
            x = x.float().double()
        h = torch.tanh(x)
        if faithful:
            h = _bf16(h)
        del x
        h2 = h.reshape(T * U1, H)
        z = h2 @ W.t() + bias_
        lse = torch.logsumexp(z, dim=-1)
        lp = z - lse[:, None]
        del z
        lp3 = lp.view(T, U1, V)
        lpb = lp3[:, :, blank].contiguous()
        lpl = torch.full((T, U1), NEG, dtype=torch.float64, device=dev)
        if U > 0:
            yb = torch.as_tensor(y[b, :U].astype(np.int64), device=dev)
            lpl[:, :U] = lp3[:, :U, :].gather(-1, yb[None, :, None].expand(T, U, 1)).squeeze(-1)
        alpha, beta, lnp = _lattice(lpb, lpl, T, U)
        loss[b] = -float(lnp)
        # occupancies (Appendix A): c1 blank arc, c2 label arc
        c1 = torch.zeros(T, U1, dtype=torch.float64, device=dev)
        c2 = torch.zeros(T, U1, dtype=torch.float64, device=dev)
        if T > 1:
            c1[: T - 1] = torch.exp(alpha[: T - 1] + lpb[: T - 1] + beta[1:] - lnp)
        c1[T - 1, U] = torch.exp(alpha[T - 1, U] + lpb[T - 1, U] - lnp)
        if U > 0:
            c2[:, :U] = torch.exp(alpha[:, :U] + lpl[:, :U] + beta[:, 1:] - lnp)
        dz = torch.exp(lp3) * (c1 + c2)[:, :, None]
        del lp, lp3
        dz[:, :, blank] -= c1
        if U > 0:
            dz[:, :U, :].scatter_add_(-1, yb[None, :, None].expand(T, U, 1), -c2[:, :U, None])
        dz = dz * float(gl[b])
        if faithful:
            dz = _bf16(dz)
        dz2 = dz.view(T * U1, V)
        db += dz2.sum(0)
        dW += dz2.t() @ h2
        dh = dz2 @ W
        del dz, dz2
        dpre = (dh * (1.0 - h2 * h2)).view(T, U1, H)
        df[b, :T] = dpre.sum(1)
        dg[b, :U1] = dpre.sum(0)
        del dh, dpre, h, h2
    n = lambda t: t.cpu().numpy()  # noqa: E731
    return dict(loss=loss, df=n(df), dg=n(dg), dW=n(dW), db=n(db))

"""Autograd-level entry points over the C-ABI library (no torch types cross the boundary)."""
import os
from typing import Optional, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _on(t: torch.Tensor):
    """Device guard: the library launches on the *current* device, so make the tensors' device current (a process
    that drives several GPUs may call with tensors of any of them)."""
    return torch.cuda.device(t.device)


def _host_i32(x) -> torch.Tensor:
    """Lengths as a contiguous host int32 tensor (they originate on the host in the reference,
    data/batch.py:103-105; a device tensor costs one sync here)."""
    if isinstance(x, torch.Tensor):
        return x.detach().to(device="cpu", dtype=torch.int32).contiguous()
    return torch.tensor(list(x), dtype=torch.int32)


#: One scratch workspace per (device, stream).  A fused call needs ~0.4 GB of scratch at the target shape; only its first
#: ``rnnt_fused_state_bytes`` bytes (~40 MB) carry information from the forward to the backward call.  The forward op
#: returns that prefix as its own tensor (saved for backward by autograd) and the backward op copies it back to the
#: front of the pooled workspace, so any number of live graphs share ONE scratch buffer per stream and nothing of
#: 0.4 GB is allocated per step.  Calls on one stream are ordered, which is all the sharing needs.
_ws_pool: dict = {}


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_pool.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_pool[key] = ws
    return ws


def _align1k(x):
    return (x + 1023) // 1024 * 1024


def _state_bytes(B, Tmax, Umax, V, H):
    """``rnnt_fused_state_bytes`` restated in Python so that the fake (meta) implementation works on symbolic sizes:
    tile prefix + two length arrays + row lse + four per-cell fp32 arrays in the diagonal layout, each 1 KB aligned
    (``csrc/capi.cu::make_plan``; ``tests/test_modules_cpu.py`` checks it against the library)."""
    U1 = Umax + 1
    max_tiles = B * ((Tmax + 15) // 16) * ((U1 + 7) // 8)
    cells = B * (Tmax + U1) * U1
    return (_align1k(4 * (B + 1)) + 2 * _align1k(4 * B) + _align1k(4 * max_tiles * 128) + 4 * _align1k(4 * cells))


#: Keep the joint's activations -- logits (fp16) and h = tanh(f + g) (bf16) -- from the forward to the backward call
#: instead of recomputing them there: 6.8 GB per live graph at B=32 T=500 U=100 V=H=1024 against 2 of the 8 N*H*V flops of
#: a step (include/rnnt_b200.h, ``rnnt_fused_forward_keep``).  ``RNNT_KEEP_ACTIVATIONS=0`` or
#: ``set_keep_activations(False)`` selects the recompute schedule; so does an allocation that does not fit.
_keep = os.environ.get("RNNT_KEEP_ACTIVATIONS", "1") != "0"
#: Largest buffer (bytes) the operator keeps per live graph; beyond it the backward pass recomputes.  A pure function of the
#: shapes, so that the fake (meta) implementation and the real one agree.  32 GiB default: 4.5x the target shape.
_keep_cap = int(os.environ.get("RNNT_KEEP_MAX_BYTES", str(32 << 30)))


def set_keep_activations(flag: bool, max_bytes: Optional[int] = None) -> None:
    """Choose between keeping logits and h for the backward pass (default) and recomputing them (less memory);
    ``max_bytes`` changes the per-graph budget above which they are recomputed anyway."""
    global _keep, _keep_cap
    _keep = bool(flag)
    if max_bytes is not None:
        _keep_cap = int(max_bytes)


def _kept_bytes(B, Tmax, Umax, V, H):
    """``rnnt_fused_kept_bytes`` restated for symbolic sizes: per lattice tile 128 rows x (fp16 padded vocabulary + bf16 H);
    nothing is kept for more than 4096 vocabulary columns (``tests/test_modules_cpu.py`` checks it against the library)."""
    if not _keep:
        return 0
    Vp = (V + 63) // 64 * 64
    if Vp > 4096:
        return 0
    max_tiles = B * ((Tmax + 15) // 16) * ((Umax + 1 + 7) // 8)
    nbytes = _align1k(2 * max_tiles * 128 * Vp) + 2 * max_tiles * 128 * H
    return nbytes if nbytes <= _keep_cap else 0


def _bf16c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.bfloat16).contiguous()


@torch.library.custom_op("rnnt_b200::fused_joint_loss", mutates_args=(), device_types="cuda")
def _fused_joint_loss(f: torch.Tensor, g: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], y: torch.Tensor,
                      f_lens: torch.Tensor, y_lens: torch.Tensor, blank: int
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(loss (B,) fp32, state (rnnt_fused_state_bytes,) uint8, kept (rnnt_fused_kept_bytes,) uint8)``: joint +
    log-softmax + alpha/beta through ``rnnt_fused_forward_keep``; ``state`` and ``kept`` (empty when they are
    recomputed instead of kept) are what ``rnnt_b200::fused_joint_loss_backward`` needs back."""
    lib = _lib.load()
    B, Tmax, H = f.shape
    Umax = g.shape[1] - 1
    V = W.shape[0]
    fb, gb, Wb = _bf16c(f), _bf16c(g), _bf16c(W)
    bf = None if bias is None else bias.detach().to(torch.float32).contiguous()
    yi = y.detach().to(device=f.device, dtype=torch.int32).contiguous()
    fl, yl = _host_i32(f_lens), _host_i32(y_lens)
    nbytes = lib.rnnt_fused_workspace_bytes(B, Tmax, Umax, V, H)
    if nbytes == 0:
        raise ValueError(lib.rnnt_last_error().decode())
    sbytes = lib.rnnt_fused_state_bytes(B, Tmax, Umax, V, H)
    with _on(fb):
        ws = _workspace(nbytes, f.device)
        ws[:sbytes].zero_()       # the per-cell arrays have cells no kernel writes (outside the lattice): keep them defined
        loss = torch.empty(B, dtype=torch.float32, device=f.device)
        zbytes = _kept_bytes(B, Tmax, Umax, V, H)
        try:
            kept = torch.empty(zbytes, dtype=torch.uint8, device=f.device)
        except torch.OutOfMemoryError:
            kept = torch.empty(0, dtype=torch.uint8, device=f.device)
        _lib.check(lib.rnnt_fused_forward_keep(_ptr(fb), _ptr(gb), _ptr(Wb), _ptr(bf), _ptr(yi), _ptr(fl), _ptr(yl),
                                               B, Tmax, Umax, V, H, int(blank), _ptr(loss), _ptr(ws), nbytes,
                                               _ptr(kept) if kept.numel() else None, kept.numel(),
                                               _stream(f.device)))
        state = ws[:sbytes].clone()
    return loss, state, kept


@_fused_joint_loss.register_fake
def _(f, g, W, bias, y, f_lens, y_lens, blank):
    B, Tmax, H = f.shape
    return (f.new_empty((B,), dtype=torch.float32),
            f.new_empty((_state_bytes(B, Tmax, g.shape[1] - 1, W.shape[0], H),), dtype=torch.uint8),
            f.new_empty((_kept_bytes(B, Tmax, g.shape[1] - 1, W.shape[0], H),), dtype=torch.uint8))


@torch.library.custom_op("rnnt_b200::fused_joint_loss_backward", mutates_args=(), device_types="cuda")
def _fused_joint_loss_backward(grad_loss: torch.Tensor, f: torch.Tensor, g: torch.Tensor, W: torch.Tensor,
                               bias: Optional[torch.Tensor], y: torch.Tensor, f_lens: torch.Tensor, y_lens: torch.Tensor,
                               state: torch.Tensor, kept: torch.Tensor, blank: int
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``(df, dg, dW, db)`` in fp32 through ``rnnt_fused_backward_kept``; ``state`` and ``kept`` are the forward op's
    second and third outputs."""
    lib = _lib.load()
    B, Tmax, H = f.shape
    Umax = g.shape[1] - 1
    V = W.shape[0]
    dev = f.device
    fb, gb, Wb = _bf16c(f), _bf16c(g), _bf16c(W)
    bf = None if bias is None else bias.detach().to(torch.float32).contiguous()
    yi = y.detach().to(device=dev, dtype=torch.int32).contiguous()
    fl, yl = _host_i32(f_lens), _host_i32(y_lens)
    gl = grad_loss.detach().to(torch.float32).contiguous()
    nbytes = lib.rnnt_fused_workspace_bytes(B, Tmax, Umax, V, H)
    df = torch.empty(B, Tmax, H, dtype=torch.float32, device=dev)
    dg = torch.empty(B, Umax + 1, H, dtype=torch.float32, device=dev)
    dW = torch.empty(V, H, dtype=torch.float32, device=dev)
    db = torch.empty(V, dtype=torch.float32, device=dev)
    with _on(fb):
        ws = _workspace(nbytes, dev)
        ws[: state.numel()].copy_(state)
        _lib.check(lib.rnnt_fused_backward_kept(_ptr(fb), _ptr(gb), _ptr(Wb), _ptr(bf), _ptr(yi), _ptr(fl), _ptr(yl),
                                                B, Tmax, Umax, V, H, int(blank), _ptr(gl), _ptr(df), _ptr(dg), _ptr(dW),
                                                _ptr(db), _ptr(ws), nbytes,
                                                _ptr(kept) if kept.numel() else None, kept.numel(), _stream(dev)))
    return df, dg, dW, db


@_fused_joint_loss_backward.register_fake
def _(grad_loss, f, g, W, bias, y, f_lens, y_lens, state, kept, blank):
    B, Tmax, H = f.shape
    V = W.shape[0]
    new = lambda *shape: f.new_empty(shape, dtype=torch.float32)  # noqa: E731
    return new(B, Tmax, H), new(B, g.shape[1], H), new(V, H), new(V)


def _setup_context(ctx, inputs, output):
    f, g, W, bias, y, f_lens, y_lens, blank = inputs
    _, state, kept = output
    ctx.save_for_backward(f, g, W, bias, y, f_lens, y_lens, state, kept)
    # state and kept carry no gradient: without this autograd materialises zero "gradients" for them in every backward
    # pass -- a 3.5 GB fill at the target shape (0.9 ms of a 10 ms step, measured)
    ctx.mark_non_differentiable(state, kept)
    ctx.set_materialize_grads(False)
    ctx.blank = blank
    ctx.has_bias = bias is not None


def _backward(ctx, grad_loss, _grad_state, _grad_kept):
    f, g, W, bias, y, f_lens, y_lens, state, kept = ctx.saved_tensors
    if grad_loss is None:
        grad_loss = torch.zeros(f.shape[0], dtype=torch.float32, device=f.device)
    df, dg, dW, db = torch.ops.rnnt_b200.fused_joint_loss_backward(grad_loss, f, g, W, bias, y, f_lens, y_lens, state,
                                                                    kept, ctx.blank)
    return (df.to(f.dtype), dg.to(g.dtype), dW.to(W.dtype), db.to(bias.dtype) if ctx.has_bias else None,
            None, None, None, None)


_fused_joint_loss.register_autograd(_backward, setup_context=_setup_context)


def rnnt_joint_loss(f: torch.Tensor, g: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor],
                    y: torch.Tensor, f_lens, y_lens, blank: int) -> torch.Tensor:
    """Per-utterance ``-ln P(y|x)`` of the additive-tanh joint, fused with its lattice.

    f (B,T,H), g (B,U+1,H), W (V,H), bias (V) or None, y (B,U) int; returns (B,) fp32.
    The (B,T,U+1,V) logits exist only as the fp16 copy kept for the backward pass (``set_keep_activations(False)``: never
    materialised, recomputed there instead).  The arithmetic is the registered operator
    ``torch.ops.rnnt_b200.fused_joint_loss`` (CUDA only; fake implementation and autograd formula registered, so it
    traces under ``torch.compile`` / ``torch.export`` as one opaque node).
    """
    _lib.load()
    if not f.is_cuda:
        raise _lib.RNNTLibraryError("rnnt fused joint+loss needs CUDA tensors; there is no CPU fallback")
    B = f.shape[0]
    Umax = g.shape[1] - 1
    if y.dim() != 2 or y.shape[0] != B or (Umax > 0 and y.shape[1] != Umax):
        raise ValueError(f"targets must have shape ({B}, {Umax}), got {tuple(y.shape)}")
    fl = f_lens if isinstance(f_lens, torch.Tensor) else torch.tensor(list(f_lens), dtype=torch.int32)
    yl = y_lens if isinstance(y_lens, torch.Tensor) else torch.tensor(list(y_lens), dtype=torch.int32)
    if fl.numel() != B or yl.numel() != B:
        raise ValueError(f"length tensors must have {B} entries")
    return torch.ops.rnnt_b200.fused_joint_loss(f, g, W, bias, y, fl, yl, int(blank))[0]


class _LatticeLoss(torch.autograd.Function):
    """RNNTLoss on an explicit (B,T,U+1,V) logits tensor: torch does the log-softmax/gather plumbing,
    the alpha/beta lattice runs in the CUDA library."""

    @staticmethod
    def forward(ctx, logits, y, f_lens, y_lens, blank):
        lib = _lib.load()
        if not logits.is_cuda:
            raise _lib.RNNTLibraryError("rnnt lattice loss needs CUDA tensors; there is no CPU fallback")
        B, Tmax, U1, V = logits.shape
        Umax = U1 - 1
        lp = torch.log_softmax(logits.detach().float(), dim=-1)
        yi = y.detach().to(device=logits.device, dtype=torch.int64)
        lpb = lp[..., blank].contiguous()
        lpl = torch.zeros(B, Tmax, U1, dtype=torch.float32, device=logits.device)
        if Umax > 0:
            idx = yi[:, None, :, None].expand(B, Tmax, Umax, 1)
            lpl[:, :, :Umax] = lp[:, :, :Umax].gather(-1, idx).squeeze(-1)
        fl, yl = _host_i32(f_lens), _host_i32(y_lens)
        nbytes = lib.rnnt_lattice_workspace_bytes(B, Tmax, Umax)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
        loss = torch.empty(B, dtype=torch.float32, device=logits.device)
        c1 = torch.empty(B, Tmax, U1, dtype=torch.float32, device=logits.device)
        c2 = torch.empty_like(c1)
        with _on(logits):
            _lib.check(lib.rnnt_lattice_forward(_ptr(lpb), _ptr(lpl), _ptr(fl), _ptr(yl), B, Tmax, Umax, _ptr(loss),
                                                _ptr(c1), _ptr(c2), _ptr(ws), nbytes, _stream(logits.device)))
        ctx.save_for_backward(lp, c1, c2, yi)
        ctx.blank = blank
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lp, c1, c2, yi = ctx.saved_tensors
        B, Tmax, U1, V = lp.shape
        Umax = U1 - 1
        dz = torch.exp(lp) * (c1 + c2)[..., None]
        dz[..., ctx.blank] -= c1
        if Umax > 0:
            idx = yi[:, None, :, None].expand(B, Tmax, Umax, 1)
            dz[:, :, :Umax].scatter_add_(-1, idx, -c2[:, :, :Umax, None])
        dz = dz * grad_loss.float()[:, None, None, None]
        return dz.to(ctx.in_dtype), None, None, None, None


def rnnt_loss_from_logits(logits, y, f_lens, y_lens, blank: int) -> torch.Tensor:
    return _LatticeLoss.apply(logits, y, f_lens, y_lens, blank)


def greedy_joint_argmax(f: torch.Tensor, g: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor],
                        t_idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b] = argmax_v joint(f[b, t_idx[b]], g[b]); -1 where t_idx[b] < 0.  All bf16/int32 CUDA tensors."""
    lib = _lib.load()
    B, Tmax, H = f.shape
    V = W.shape[0]
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=f.device)
    with _on(f):
        _lib.check(lib.rnnt_greedy_joint_argmax(_ptr(f), _ptr(g), _ptr(W), _ptr(bias), _ptr(t_idx), _ptr(out),
                                                B, Tmax, V, H, _stream(f.device)))
    return out


def greedy_step(f: torch.Tensor, g: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], lens: torch.Tensor,
                t_cur: torch.Tensor, emitted: torch.Tensor, n_sym: torch.Tensor, sym: torch.Tensor,
                is_sym: torch.Tensor, label: torch.Tensor, active: torch.Tensor, blank: int, max_symbols: int) -> None:
    """One greedy decode step with its bookkeeping, in place (see ``rnnt_greedy_step`` in include/rnnt_b200.h).
    f bf16 (B,T,H), g fp32 (B,H), W bf16 (V,H); every state tensor is int32 on the device."""
    lib = _lib.load()
    B, Tmax, H = f.shape
    V = W.shape[0]
    with _on(f):
        _lib.check(lib.rnnt_greedy_step(_ptr(f), _ptr(g), _ptr(W), _ptr(bias), _ptr(lens), _ptr(t_cur), _ptr(emitted),
                                        _ptr(n_sym), _ptr(sym), sym.shape[1], _ptr(is_sym), _ptr(label), _ptr(active),
                                        B, Tmax, V, H, int(blank), int(max_symbols), _stream(f.device)))


def greedy_decode_lstm_supported(B: int, V: int, H: int, Hp: int, n_layers: int = 1) -> bool:
    """Whether the one-launch decode (``rnnt_greedy_decode_lstm_stack``) covers this shape."""
    return _lib.load().rnnt_greedy_decode_stack_workspace_bytes(B, V, H, Hp, n_layers) > 0


def greedy_decode_lstm(f: torch.Tensor, lens: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor],
                       gate_table: torch.Tensor, W_hh: torch.Tensor, W_proj: torch.Tensor,
                       bias_proj: Optional[torch.Tensor], blank: int, max_symbols: int,
                       W_upper: Optional[torch.Tensor] = None, bias_upper: Optional[torch.Tensor] = None,
                       cell: str = "lstm"):
    """Whole greedy decode of a batch in one launch (see ``rnnt_greedy_decode_lstm_stack`` / ``rnnt_greedy_decode_gru_stack``
    in include/rnnt_b200.h; ``cell="gru"`` expects the four-rows-per-unit GRU packing described there).

    f bf16 (B,T,H); lens int32 (B) on the device; W bf16 (V,H); gate_table fp32 (V+1, 4*Hp); W_hh bf16 (4*Hp, Hp);
    W_proj bf16 (H, Hp); for an n-layer LSTM ``W_upper`` bf16 (n-1, 4*Hp, 2*Hp) = [W_ih_l | W_hh_l] and ``bias_upper``
    fp32 (n-1, 4*Hp).  Returns ``(sym int32 (B, T*max_symbols), n_sym int32 (B))`` on the device.
    """
    lib = _lib.load()
    B, Tmax, H = f.shape
    V = W.shape[0]
    Hp = W_hh.shape[1]
    n_layers = 1 if W_upper is None else W_upper.shape[0] + 1
    if gate_table.shape != (V + 1, 4 * Hp) or W_hh.shape[0] != 4 * Hp or tuple(W_proj.shape) != (H, Hp):
        raise ValueError(f"gate_table {tuple(gate_table.shape)}, W_hh {tuple(W_hh.shape)}, W_proj {tuple(W_proj.shape)} "
                         f"do not match V={V} H={H} Hp={Hp}")
    if W_upper is not None and (tuple(W_upper.shape[1:]) != (4 * Hp, 2 * Hp) or bias_upper is None
                                or tuple(bias_upper.shape) != (n_layers - 1, 4 * Hp)):
        raise ValueError(f"W_upper {tuple(W_upper.shape)} / bias_upper do not match Hp={Hp}")
    nbytes = lib.rnnt_greedy_decode_stack_workspace_bytes(B, V, H, Hp, n_layers)
    if nbytes == 0:
        raise ValueError(f"fused greedy decode does not cover B={B} V={V} H={H} Hp={Hp} n_layers={n_layers}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=f.device)
    cap = max(1, Tmax * max_symbols)
    sym = torch.zeros(B, cap, dtype=torch.int32, device=f.device)
    n_sym = torch.zeros(B, dtype=torch.int32, device=f.device)
    if cell not in ("lstm", "gru"):
        raise ValueError(f"cell={cell!r} must be 'lstm' or 'gru'")
    entry = lib.rnnt_greedy_decode_lstm_stack if cell == "lstm" else lib.rnnt_greedy_decode_gru_stack
    with _on(f):
        _lib.check(entry(_ptr(f), _ptr(lens), _ptr(W), _ptr(bias), _ptr(gate_table), _ptr(W_hh),
                         n_layers, _ptr(W_upper), _ptr(bias_upper), _ptr(W_proj), _ptr(bias_proj),
                         B, Tmax, V, H, Hp, int(blank), int(max_symbols), _ptr(sym), cap,
                         _ptr(n_sym), _ptr(ws), nbytes, _stream(f.device)))
    return sym, n_sym

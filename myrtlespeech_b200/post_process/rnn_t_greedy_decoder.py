"""RNN-T greedy decoder with the interface of ``post_process/ctc_greedy_decoder.py:17-97``."""
from typing import List

import torch

from ..functional import greedy_joint_argmax


class RNNTGreedyDecoder(torch.nn.Module):
    """Decodes RNN-T output using a greedy strategy.

    For every frame, symbols are emitted while the joint's argmax is not blank, up to
    ``max_symbols_per_step`` per frame.  The joint step (tanh + projection + argmax) runs in the
    CUDA library for the whole batch at once; only the emitted ids (``batch`` int32) come back to the
    host per step, not logits.

    Args:
        blank_index: Index of the "blank" symbol.
        model: An :py:class:`myrtlespeech_b200.model.RNNT`.
        max_symbols_per_step: Maximum number of non-blank symbols per encoder frame.
    """

    def __init__(self, blank_index: int, model: torch.nn.Module, max_symbols_per_step: int = 4):
        super().__init__()
        if max_symbols_per_step < 1:
            raise ValueError(f"max_symbols_per_step={max_symbols_per_step} must be >= 1")
        self.blank_index = blank_index
        self.max_symbols_per_step = max_symbols_per_step
        # not registered as a submodule: the decoder does not own the model's parameters
        object.__setattr__(self, "_model", model)

    @property
    def model(self):
        return self._model

    @torch.no_grad()
    def forward(self, x: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
        r"""Decodes using a greedy strategy.

        Args:
            x: ``(batch, seq_len, hidden)`` encoder output ``f`` (already projected to the joint
                width), or raw features if ``model.encoder`` should be applied -- decided by the
                last dimension matching ``model.joint.hidden_size``.
            lengths: 1D integer tensor of valid frames per sequence.

        Returns:
            ``List[List[int]]`` of emitted symbol ids per sequence.

        Raises:
            :py:class:`ValueError`: if ``lengths.dtype`` is not an integer type, if the batch sizes
                differ, or if any length exceeds ``seq_len`` (same checks and messages as
                ``post_process/ctc_greedy_decoder.py:49-72``).
        """
        supported_dtypes = [torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64]
        if lengths.dtype not in supported_dtypes:
            raise ValueError(f"lengths.dtype={lengths.dtype} must be in {supported_dtypes}")
        x_batch, seq_len, _ = x.size()
        l_batch = len(lengths)
        if x_batch != l_batch:
            raise ValueError(f"batch size of x ({x_batch}) and lengths {l_batch} must be equal")
        if not (lengths <= seq_len).all():
            raise ValueError("length values must be less than or equal to x seq_len")

        model = self._model
        was_training = model.training
        model.eval()
        try:
            if x.size(2) != model.joint.hidden_size:
                x, lengths = model.encode(x, lengths)
            return self._decode(x, lengths)
        finally:
            model.train(was_training)

    def _decode(self, f: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
        model = self._model
        dev = model.joint.fc.weight.device
        B, T, H = f.shape
        fb = f.to(dev, torch.bfloat16).contiguous()
        Wb = model.joint.fc.weight.detach().to(torch.bfloat16).contiguous()
        bias = model.joint.fc.bias
        bias = None if bias is None else bias.detach().float().contiguous()
        lens = lengths.to("cpu", torch.int64)
        pred = model.prediction

        g, hid = pred.step(None, None, B, dev)
        out: List[List[int]] = [[] for _ in range(B)]
        t_host = torch.zeros(B, dtype=torch.int64)          # current frame of each utterance
        emitted = torch.zeros(B, dtype=torch.int64)          # symbols emitted at the current frame
        active = t_host < lens
        t_idx = torch.empty(B, dtype=torch.int32, device=dev)
        k_dev = torch.empty(B, dtype=torch.int32, device=dev)
        while bool(active.any()):
            t_idx.copy_(torch.where(active, t_host, torch.full_like(t_host, -1)).to(torch.int32))
            greedy_joint_argmax(fb, g.to(torch.bfloat16).contiguous(), Wb, bias, t_idx, k_dev)
            k = k_dev.cpu().to(torch.int64)
            is_sym = active & (k != self.blank_index)
            for b in torch.nonzero(is_sym).flatten().tolist():
                out[b].append(int(k[b]))
            if bool(is_sym.any()):
                # advance the prediction network only where a symbol was emitted
                g_new, hid_new = pred.step(k.clamp(min=0).to(dev), hid, B, dev)
                m = is_sym.to(dev)
                g = torch.where(m[:, None], g_new, g)
                hid = _select_hidden(m, hid_new, hid)
            emitted = torch.where(is_sym, emitted + 1, emitted)
            advance = active & (~is_sym | (emitted >= self.max_symbols_per_step))
            t_host = torch.where(advance, t_host + 1, t_host)
            emitted = torch.where(advance, torch.zeros_like(emitted), emitted)
            active = t_host < lens
        return out

    def extra_repr(self) -> str:
        return f"blank_index={self.blank_index}, max_symbols_per_step={self.max_symbols_per_step}"


def _select_hidden(mask: torch.Tensor, new, old):
    if new is None or old is None:  # stateless prediction network, or first step (old state is "zeros")
        return new
    if isinstance(new, tuple):
        return tuple(_select_hidden(mask, n, o) for n, o in zip(new, old))
    return torch.where(mask[None, :, None], new, old)

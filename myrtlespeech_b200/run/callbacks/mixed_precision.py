"""bf16 mixed precision for the RNN-T path without Apex.

The reference's ``MixedPrecision`` (``run/callbacks/mixed_precision.py:15-76``) is Apex AMP ``O1`` (fp16 + loss
scaling) and its ``ClipGradNorm`` reads ``amp.master_params`` (``run/callbacks/clip_grad_norm.py:34-39``); Apex is
not a dependency here.  bf16 keeps fp32's exponent range, so there is no loss scaling: the callback only moves the
batch to the GPU (what the reference's callback also does, ``:52-60``) and runs the model under
``torch.autocast(dtype=torch.bfloat16)``.  The fused joint + loss consumes bf16 operands and accumulates in fp32
regardless, so it needs no casting policy of its own.
"""
import contextlib
from typing import Dict, Optional, Union

import torch

from .callback import Callback, ModelCallback


def _to_cuda(x):
    if isinstance(x, torch.Tensor):
        return x.cuda()
    if isinstance(x, dict):
        return {k: _to_cuda(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        y = [_to_cuda(v) for v in x]
        return tuple(y) if isinstance(x, tuple) else y
    return x


class BF16MixedPrecision(Callback):
    """Drop-in for ``MixedPrecision`` in the callback list: ``on_batch_begin`` moves ``last_input`` to the GPU and
    opens a bf16 autocast region that ``on_loss_begin`` closes again (the loss runs in fp32)."""

    def __init__(self):
        super().__init__()
        if not torch.cuda.is_available():
            raise ValueError("cuda not available")
        self.stack = contextlib.ExitStack()

    def on_batch_begin(self, **kwargs) -> Dict:
        self.stack.close()
        self.stack.enter_context(torch.autocast(device_type="cuda", dtype=torch.bfloat16))
        return {"last_input": _to_cuda(kwargs["last_input"])}

    def on_loss_begin(self, **kwargs) -> None:
        self.stack.close()

    def on_batch_end(self, **kwargs) -> None:
        self.stack.close()


class ClipGradNorm(ModelCallback):
    """``ClipGradNorm`` over the model's own parameters (there are no Apex master copies with bf16 autocast).

    Args:
        model: the ``SeqToSeq`` / ``SpeechToText`` container or any module with ``parameters()`` (the reference takes
            the container too, ``run/callbacks/clip_grad_norm.py:23-32``).
        max_norm, norm_type: see :py:func:`torch.nn.utils.clip_grad_norm_`.
    """

    def __init__(self, model, max_norm: Union[float, int], norm_type: Union[float, int] = 2):
        super().__init__(model)
        self.max_norm = max_norm
        self.norm_type = norm_type
        self.last_norm: Optional[float] = None

    def on_backward_end(self, **kwargs) -> None:
        if not self.training:
            return
        self.last_norm = float(torch.nn.utils.clip_grad_norm_(
            parameters=self.model.parameters(), max_norm=self.max_norm, norm_type=self.norm_type))

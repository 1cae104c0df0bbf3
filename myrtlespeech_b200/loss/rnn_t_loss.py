"""RNN-T loss with the ``forward(inputs, targets)`` contract of ``loss/ctc_loss.py:51-101``."""
from typing import Tuple, Union

import torch

from ..functional import rnnt_joint_loss, rnnt_loss_from_logits
from ..model.rnn_t import JointHandle


class RNNTLoss(torch.nn.Module):
    """Transducer loss that owns the log-softmax, as the CTC wrapper does (``loss/ctc_loss.py:45,95``).

    Args:
        blank: Index of the blank label.

        reduction: Specifies the reduction to apply to the output:

            none:
                No reduction will be applied; a ``(batch,)`` tensor is returned.

            mean:
                The mean over the batch of the per-utterance losses.  (The reference does not define
                RNN-T ``mean``; its CTC ``mean`` divides by target length first,
                ``loss/ctc_loss.py:17-19``.  Batch mean is what ``torchaudio`` does.)

            sum:
                Sum all losses in a batch.

    The arithmetic runs in the CUDA library (``include/rnnt_b200.h``); there is no CPU fallback.
    """

    def __init__(self, blank: int = 0, reduction: str = "mean"):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(f"reduction={reduction} not supported")
        if blank < 0:
            raise ValueError(f"blank={blank} must be >= 0")
        self.blank = blank
        self.reduction = reduction
        self.use_cuda = torch.cuda.is_available()

    def forward(
        self,
        inputs: Tuple[Union[JointHandle, torch.Tensor], torch.Tensor],
        targets: Tuple[torch.Tensor, torch.Tensor],
    ) -> torch.Tensor:
        """Computes the RNN-T loss.

        Args:
            inputs: ``(x, x_lens)``.  ``x`` is either the :py:class:`JointHandle` returned by
                :py:class:`RNNTJoint` (fused path, no ``(B,T,U+1,V)`` tensor) or a dense
                unnormalised logits tensor of size ``(batch, max_seq_len, max_target_len + 1,
                features)``.  ``x_lens`` gives the number of valid frames per sequence.

            targets: ``(y, y_lens)``.  ``y`` has size ``(batch, max_target_len)`` and integer dtype;
                entries cannot be the blank index.  ``y_lens`` gives the target lengths.

        Raises:
            :py:class:`ValueError`: on mismatched batch sizes, lengths that exceed the padded sizes,
                non-integer length dtypes or a blank index outside the vocabulary.
        """
        x, x_lens = inputs
        y, y_lens = targets
        int_types = [torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64]
        for name, t in (("x_lens", x_lens), ("y_lens", y_lens), ("y", y)):
            if t.dtype not in int_types:
                raise ValueError(f"{name}.dtype={t.dtype} must be in {int_types}")
        B, T, U1, V = x.shape
        if len(x_lens) != B or len(y_lens) != B or y.size(0) != B:
            raise ValueError(f"batch size of x ({B}), x_lens ({len(x_lens)}), y ({y.size(0)}) and y_lens ({len(y_lens)}) must be equal")
        if y.dim() != 2 or y.size(1) != U1 - 1:
            raise ValueError(f"y must have size ({B}, {U1 - 1}), got {tuple(y.shape)}")
        if not (0 <= self.blank < V):
            raise ValueError(f"blank={self.blank} must be in [0, {V - 1}]")
        xl, yl = x_lens.detach().cpu(), y_lens.detach().cpu()
        if not bool((xl <= T).all()) or not bool((xl >= 1).all()):
            raise ValueError("x_lens values must be in [1, x seq_len]")
        if not bool((yl <= U1 - 1).all()):
            raise ValueError("y_lens values must be less than or equal to y seq_len")

        if isinstance(x, JointHandle):
            f, g = x.f, x.g
            if self.use_cuda:
                f, g = f.cuda(), g.cuda()
            loss = rnnt_joint_loss(f, g, x.weight, x.bias, y, xl, yl, self.blank)
        else:
            if self.use_cuda:
                x = x.cuda()
            loss = rnnt_loss_from_logits(x, y, xl, yl, self.blank)

        if self.reduction == "sum":
            return loss.sum()
        if self.reduction == "mean":
            return loss.mean()
        return loss

    def extra_repr(self) -> str:
        return f"blank={self.blank}, reduction={self.reduction}"

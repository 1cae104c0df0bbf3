"""B200-native RNN-T transducer head (joint + RNNTLoss + greedy decode) behind myrtlespeech's operator surface."""
from . import _lib  # noqa: F401
from .functional import greedy_joint_argmax, rnnt_joint_loss, rnnt_loss_from_logits  # noqa: F401

__all__ = ["rnnt_joint_loss", "rnnt_loss_from_logits", "greedy_joint_argmax"]

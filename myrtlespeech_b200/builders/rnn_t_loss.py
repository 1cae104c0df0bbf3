from ..loss.rnn_t_loss import RNNTLoss
from ..protos import rnn_t_loss_pb2


def build(rnn_t_loss_cfg) -> RNNTLoss:
    """Returns a :py:class:`.RNNTLoss` based on the config (cf. ``builders/ctc_loss.py:5-43``).

    Example:
        >>> from google.protobuf import text_format
        >>> cfg = text_format.Merge('''
        ... blank_index: 28;
        ... reduction: SUM;
        ... ''', rnn_t_loss_pb2.RNNTLoss())
        >>> build(cfg)
        RNNTLoss(blank=28, reduction=sum)
    """
    reduction_map = {
        rnn_t_loss_pb2.RNNTLoss.NONE: "none",
        rnn_t_loss_pb2.RNNTLoss.MEAN: "mean",
        rnn_t_loss_pb2.RNNTLoss.SUM: "sum",
    }
    try:
        reduction = reduction_map[rnn_t_loss_cfg.reduction]
    except KeyError:
        raise ValueError(f"reduction={rnn_t_loss_cfg.reduction} not supported")

    return RNNTLoss(blank=rnn_t_loss_cfg.blank_index, reduction=reduction)

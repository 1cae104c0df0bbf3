from typing import Tuple

from ..model.rnn_t import RNNT, RNNTJoint, RNNTPredictionNet, _LinearEncoder
from ..protos import rnn_t_pb2


def build(rnn_t_cfg, input_features: int, output_features: int, input_channels: int = 1) -> Tuple[RNNT, int]:
    """Returns an :py:class:`.RNNT` based on the config and its number of output features.

    ``input_features`` / ``input_channels`` are the sizes of the feature and channel axes of the model input
    ``(batch, channels, features, seq_len)``, derived from the pre-processing steps as in the reference
    (``builders/speech_to_text.py:249-272``; cf. ``builders/deep_speech_2.py`` taking the same arguments).

    Example:
        >>> from google.protobuf import text_format
        >>> cfg = text_format.Merge('''
        ... rnn_type: LSTM;
        ... encoder_hidden_size: 32; encoder_num_layers: 1;
        ... pred_embedding_size: 16; pred_hidden_size: 32; pred_num_layers: 1;
        ... joint_hidden_size: 64;
        ... ''', rnn_t_pb2.RNNT())
        >>> model, out = build(cfg, input_features=8, output_features=29)
        >>> out, model.joint.hidden_size
        (29, 64)
    """
    rnn_type_map = {rnn_t_pb2.RNNT.LSTM: "lstm", rnn_t_pb2.RNNT.GRU: "gru"}
    try:
        rnn_type = rnn_type_map[rnn_t_cfg.rnn_type]
    except KeyError:
        raise ValueError(f"rnn_type={rnn_t_cfg.rnn_type} not supported")
    H = rnn_t_cfg.joint_hidden_size
    if H < 1:
        raise ValueError(f"joint_hidden_size={H} must be >= 1")
    for name in ("encoder_hidden_size", "encoder_num_layers", "pred_embedding_size", "pred_hidden_size",
                 "pred_num_layers"):
        if getattr(rnn_t_cfg, name) < 1:
            raise ValueError(f"{name}={getattr(rnn_t_cfg, name)} must be >= 1")
    encoder = _LinearEncoder(input_features, rnn_t_cfg.encoder_hidden_size, rnn_t_cfg.encoder_num_layers, H,
                             rnn_type, input_channels=input_channels)
    prediction = RNNTPredictionNet(output_features, rnn_t_cfg.pred_embedding_size, rnn_t_cfg.pred_hidden_size,
                                   rnn_t_cfg.pred_num_layers, H, rnn_type)
    joint = RNNTJoint(H, output_features)
    return RNNT(encoder, prediction, joint), output_features

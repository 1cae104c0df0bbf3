"""``SpeechToText`` builder branch for the RNN-T members (cf. ``builders/speech_to_text.py:152-242``)."""
from typing import List

from . import rnn_t as rnn_t_builder
from . import rnn_t_greedy_decoder as decoder_builder
from . import rnn_t_loss as loss_builder
from ..protos import speech_to_text_pb2  # noqa: F401


class SpeechToText:
    """Container with the attribute names the reference's ``model/speech_to_text.py:9-33`` fixes."""

    def __init__(self, alphabet, model, loss, pre_process_steps, post_process):
        self.alphabet = alphabet
        self.model = model
        self.loss = loss
        self.pre_process_steps = pre_process_steps
        self.post_process = post_process


def build(stt_cfg) -> SpeechToText:
    """Builds the RNN-T flavour of ``SpeechToText``: model, loss and post-process by ``WhichOneof``.

    Keeps the reference's checks: every ``blank_index`` must lie in ``[0, len(alphabet) - 1]`` and
    all of them must match (``builders/speech_to_text.py:192-196,231-233``); unknown oneof members
    raise :py:class:`ValueError` (``:181-182,198-199,228-229``).

    Example:
        >>> from google.protobuf import text_format
        >>> cfg = text_format.Merge('''
        ... alphabet: "abc_";
        ... input_features: 8;
        ... rnn_t { encoder_hidden_size: 8; encoder_num_layers: 1; pred_embedding_size: 8;
        ...         pred_hidden_size: 8; pred_num_layers: 1; joint_hidden_size: 16; }
        ... rnn_t_loss { blank_index: 3; reduction: SUM; }
        ... rnn_t_greedy_decoder { blank_index: 3; max_symbols_per_step: 2; }
        ... ''', speech_to_text_pb2.SpeechToText())
        >>> stt = build(cfg)
        >>> stt.loss, stt.post_process
        (RNNTLoss(blank=3, reduction=sum), RNNTGreedyDecoder(blank_index=3, max_symbols_per_step=2))
    """
    alphabet = list(stt_cfg.alphabet)
    hi = max(0, len(alphabet) - 1)

    model_type = stt_cfg.WhichOneof("supported_models")
    if model_type == "rnn_t":
        model, _ = rnn_t_builder.build(
            stt_cfg.rnn_t, input_features=max(1, stt_cfg.input_features), output_features=len(alphabet)
        )
    else:
        raise ValueError(f"model={model_type} not supported")

    blank_indices: List[int] = []

    loss_type = stt_cfg.WhichOneof("supported_losses")
    if loss_type == "rnn_t_loss":
        blank_index = stt_cfg.rnn_t_loss.blank_index
        blank_indices.append(blank_index)
        if not (0 <= blank_index <= hi):
            raise ValueError(f"rnn_t_loss.blank_index={blank_index} must be in [0, {hi}]")
        loss = loss_builder.build(stt_cfg.rnn_t_loss)
    else:
        raise ValueError(f"loss={loss_type} not supported")

    post_process_type = stt_cfg.WhichOneof("supported_post_processes")
    if post_process_type == "rnn_t_greedy_decoder":
        blank_index = stt_cfg.rnn_t_greedy_decoder.blank_index
        blank_indices.append(blank_index)
        if not (0 <= blank_index <= hi):
            raise ValueError(f"rnn_t_greedy_decoder.blank_index={blank_index} must be in [0, {hi}]")
        post_process = decoder_builder.build(stt_cfg.rnn_t_greedy_decoder, model)
    else:
        raise ValueError(f"post_process={post_process_type} not supported")

    if blank_indices and not len(set(blank_indices)) == 1:
        raise ValueError("all blank_index values of RNN-T components must match")

    return SpeechToText(
        alphabet=alphabet, model=model, loss=loss, pre_process_steps=[], post_process=post_process
    )

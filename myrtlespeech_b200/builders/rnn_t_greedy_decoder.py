import torch

from ..post_process.rnn_t_greedy_decoder import RNNTGreedyDecoder
from ..protos import rnn_t_greedy_decoder_pb2  # noqa: F401


def build(cfg, model: torch.nn.Module) -> RNNTGreedyDecoder:
    """Returns a :py:class:`.RNNTGreedyDecoder` based on the config.

    ``max_symbols_per_step: 0`` (the proto3 default, i.e. unset) is rejected: the field has no
    meaningful zero value.
    """
    if cfg.max_symbols_per_step < 1:
        raise ValueError(f"max_symbols_per_step={cfg.max_symbols_per_step} must be >= 1")
    return RNNTGreedyDecoder(
        blank_index=cfg.blank_index, model=model, max_symbols_per_step=cfg.max_symbols_per_step
    )

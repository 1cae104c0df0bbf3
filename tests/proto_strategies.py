"""One hypothesis strategy per proto message (reference pattern: tests/protos/test_ctc_loss.py:15-38)
plus the ``all_fields_set`` guard of tests/protos/utils.py:6-46 so a new field cannot be forgotten."""
from typing import Dict, Iterable, Optional

import hypothesis.strategies as st

from myrtlespeech_b200.protos import rnn_t_greedy_decoder_pb2, rnn_t_loss_pb2, rnn_t_pb2


def all_fields_set(proto, kwargs: Dict, to_ignore: Optional[Iterable[str]] = None) -> None:
    expected = set(proto.DESCRIPTOR.fields_by_name.keys())
    if to_ignore:
        expected -= set(to_ignore)
    for oneof in proto.DESCRIPTOR.oneofs_by_name.values():
        names = set(f.name for f in oneof.fields)
        if len(names & set(kwargs)) != 1:
            raise ValueError(f"oneof field {oneof.name!r} not set correctly in {proto.DESCRIPTOR.name!r}")
        expected -= names
    if not (expected <= set(kwargs.keys())):
        raise ValueError(f"kwargs missing fields for {proto.DESCRIPTOR.name!r}: {expected - set(kwargs)}")


@st.composite
def rnn_t_losses(draw, return_kwargs: bool = False, alphabet_len: Optional[int] = None):
    kwargs = {}
    hi = 100 if alphabet_len is None else max(0, alphabet_len - 1)
    kwargs["blank_index"] = draw(st.integers(0, hi))
    kwargs["reduction"] = draw(st.sampled_from(rnn_t_loss_pb2.RNNTLoss.REDUCTION.values()))
    all_fields_set(rnn_t_loss_pb2.RNNTLoss, kwargs)
    cfg = rnn_t_loss_pb2.RNNTLoss(**kwargs)
    return (cfg, kwargs) if return_kwargs else cfg


@st.composite
def rnn_t_greedy_decoders(draw, return_kwargs: bool = False, blank_index: Optional[int] = None):
    kwargs = {}
    kwargs["blank_index"] = draw(st.integers(0, 100)) if blank_index is None else blank_index
    kwargs["max_symbols_per_step"] = draw(st.integers(1, 8))
    all_fields_set(rnn_t_greedy_decoder_pb2.RNNTGreedyDecoder, kwargs)
    cfg = rnn_t_greedy_decoder_pb2.RNNTGreedyDecoder(**kwargs)
    return (cfg, kwargs) if return_kwargs else cfg


@st.composite
def rnn_ts(draw, return_kwargs: bool = False):
    kwargs = {}
    kwargs["rnn_type"] = draw(st.sampled_from(rnn_t_pb2.RNNT.RNN_TYPE.values()))
    for name in ("encoder_hidden_size", "pred_embedding_size", "pred_hidden_size"):
        kwargs[name] = draw(st.integers(1, 16))
    for name in ("encoder_num_layers", "pred_num_layers"):
        kwargs[name] = draw(st.integers(1, 2))
    kwargs["joint_hidden_size"] = 8 * draw(st.integers(1, 4))
    all_fields_set(rnn_t_pb2.RNNT, kwargs)
    cfg = rnn_t_pb2.RNNT(**kwargs)
    return (cfg, kwargs) if return_kwargs else cfg

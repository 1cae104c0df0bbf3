"""``SpeechToText`` builder branch for the RNN-T members (cf. ``builders/speech_to_text.py:152-272``)."""
from typing import Callable, List, Tuple

from . import rnn_t as rnn_t_builder
from . import rnn_t_greedy_decoder as decoder_builder
from . import rnn_t_loss as loss_builder
from ..model.speech_to_text import SpeechToText, Stage
from ..protos import speech_to_text_pb2  # noqa: F401

try:  # pragma: no cover - depends on the environment
    from myrtlespeech.data.alphabet import Alphabet  # type: ignore
except Exception:

    class Alphabet:  # type: ignore[no-redef]
        """The part of ``data/alphabet.py:5-80`` this path uses: symbol <-> index maps."""

        def __init__(self, symbols: List[str]):
            if len(set(symbols)) != len(symbols):
                raise ValueError("Duplicate symbol in symbols")
            self.symbols = symbols
            self._index_map = dict([(s, i) for i, s in enumerate(symbols)])

        def __repr__(self) -> str:
            return f"Alphabet(symbols={self.symbols})"

        def __len__(self) -> int:
            return len(self.symbols)

        def __getitem__(self, index: int) -> str:
            return self.symbols[index]

        def get_symbols(self, indices: List[int]) -> List[str]:
            return [self.symbols[i] for i in indices if 0 <= i < len(self.symbols)]

        def get_indices(self, sentence: List[str]) -> List[int]:
            return [self._index_map[s] for s in sentence if s in self._index_map]


class _DeferredStep:
    """Stands in for a pre-processing callable when myrtlespeech's data pipeline is not importable.

    The transforms themselves (MFCC, SpecAugment, ...: ``data/preprocess.py``) belong to the reference's data
    pipeline, which this path leaves unchanged; only the *sizes* they imply are needed to build the model."""

    def __init__(self, name: str, cfg):
        self.name, self.cfg = name, cfg

    def __call__(self, x):
        raise RuntimeError(f"pre-processing step '{self.name}' needs myrtlespeech.builders.pre_process_step "
                           "(the reference's data pipeline), which is not importable here")

    def __repr__(self) -> str:
        return f"_DeferredStep({self.name})"


def _build_pre_process_steps(step_cfgs) -> Tuple[List[Tuple[Callable, Stage]], int, int]:
    """``(steps, input_features, input_channels)`` from the ``pre_process_step`` list, as
    ``builders/speech_to_text.py:249-272`` derives them: an MFCC step fixes the feature width, a context-frames step
    the channel count ``2 n_context + 1``; with no MFCC step the input is raw audio of width 1."""
    try:  # pragma: no cover - depends on the environment
        from myrtlespeech.builders.pre_process_step import build as build_step  # type: ignore
    except Exception:
        build_step = None
    input_features, input_channels = None, 1
    steps: List[Tuple[Callable, Stage]] = []
    for cfg in step_cfgs:
        kind = cfg.WhichOneof("pre_process_step")
        if kind == "mfcc":
            input_features = cfg.mfcc.n_mfcc
        elif kind == "context_frames":
            input_channels = 2 * cfg.context_frames.n_context + 1
        elif kind in ("standardize", "spec_augment"):
            pass
        else:
            raise ValueError(f"unknown pre_process_step '{kind}'")
        steps.append(build_step(cfg) if build_step is not None else (_DeferredStep(kind, cfg), Stage(cfg.stage)))
    return steps, input_features, input_channels


def build(stt_cfg) -> SpeechToText:
    """Builds the RNN-T flavour of ``SpeechToText``: model, loss and post-process by ``WhichOneof``.

    Keeps the reference's checks: every ``blank_index`` must lie in ``[0, len(alphabet) - 1]`` and
    all of them must match (``builders/speech_to_text.py:192-196,231-233``); unknown oneof members
    raise :py:class:`ValueError` (``:181-182,198-199,228-229``).  The result is an ``nn.Module`` with the
    ``model`` / ``loss`` / ``pre_process_steps`` / ``optim`` / ``alphabet`` / ``post_process`` attributes of
    ``model/speech_to_text.py:9-36``, so the reference's ``fit`` and ``Saver`` take it as it is.

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
        >>> sorted(k for k in stt.state_dict() if "joint" in k)
        ['model.joint.fc.bias', 'model.joint.fc.weight']
    """
    alphabet = Alphabet(list(stt_cfg.alphabet))
    hi = max(0, len(alphabet) - 1)

    pre_process_steps, input_features, input_channels = _build_pre_process_steps(stt_cfg.pre_process_step)
    if input_features is None:
        # no MFCC step: the reference falls back to raw audio of width 1 (builders/speech_to_text.py:266-270); the
        # `input_features` extension field covers features computed outside the config
        input_features = max(1, stt_cfg.input_features)

    model_type = stt_cfg.WhichOneof("supported_models")
    if model_type == "rnn_t":
        model, _ = rnn_t_builder.build(stt_cfg.rnn_t, input_features=input_features,
                                       output_features=len(alphabet), input_channels=input_channels)
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
        alphabet=alphabet, post_process=post_process, model=model, loss=loss, pre_process_steps=pre_process_steps
    )

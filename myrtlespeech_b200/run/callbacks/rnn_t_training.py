"""Callbacks that let an RNN-T model run under the reference's ``fit`` loop unchanged.

The loop passes only ``x`` to the model (``run/train.py:62-63``: ``out, _ = seq_to_seq.model(x)``) and
``(out, y)`` to the loss (``:66-67``), but an RNN-T model also needs the labels for its prediction network.
The reference's hook for rewriting the model input is ``Callback.on_batch_begin``: the handler stores ``x`` /
``y`` as ``last_input`` / ``last_target`` in its ``state_dict``, calls every callback with the state as keyword
arguments, merges any returned dict back, and hands the possibly modified values to the loop
(``run/callbacks/callback.py:221-254``).  The classes here derive from the reference's ``Callback`` (see
``callback.py`` in this package), so ``CallbackHandler.train(mode)`` (``run/callbacks/callback.py:483-491``,
called by ``fit`` at ``run/train.py:51``) reaches them like any other callback.
"""
from typing import Callable, Dict, List, Optional

from .callback import Callback


class RNNTTraining(Callback):
    """Packs the targets into the model input at the start of every batch.

    ``x = (feats, feat_lens)`` and ``y = (labels, label_lens)`` (the collate layout of
    ``data/batch.py:45-107``) become ``last_input = ((feats, labels), (feat_lens, label_lens))``, the input
    :py:class:`myrtlespeech_b200.model.RNNT` expects; ``last_target`` is left as it is for the loss.
    """

    def on_batch_begin(self, **kwargs) -> Optional[Dict]:
        (feats, feat_lens), (labels, label_lens) = kwargs["last_input"], kwargs["last_target"]
        return {"last_input": ((feats, labels), (feat_lens, label_lens))}


class ReportRNNTDecoder(Callback):
    """Decodes every evaluation batch and reports the error rate: the RNN-T counterpart of ``ReportCTCDecoder``
    (``run/run.py:50-109``), with the same report layout -- ``reports[<decoder class name>] = {"wer": float,
    "transcripts": [(hypothesis, reference), ...]}``, ``wer`` in percent, ``-1.0`` until an evaluation epoch ends.

    Args:
        rnnt_decoder: called as ``rnnt_decoder(*last_output)`` exactly like the reference calls its CTC decoder
            (``run/run.py:94``); ``last_output`` is ``(JointHandle, frame_lens)`` and
            :py:class:`RNNTGreedyDecoder` takes the encoder output out of the handle.
        alphabet: converts sequences of indices to sequences of symbols (``data/alphabet.py:48``).
        word_segmentor: groups symbols into words (``run/run.py:29-47``); :py:data:`None` scores symbols.
    """

    def __init__(self, rnnt_decoder, alphabet, word_segmentor: Optional[Callable] = None):
        super().__init__()
        self.rnnt_decoder = rnnt_decoder
        self.alphabet = alphabet
        self.word_segmentor = word_segmentor
        self.distances: List[int] = []
        self.lengths: List[int] = []

    @property
    def _name(self) -> str:
        return self.rnnt_decoder.__class__.__name__

    def _reset(self, **kwargs) -> None:
        kwargs["reports"][self._name] = {"wer": -1.0, "transcripts": []}
        self.distances = []
        self.lengths = []

    def on_train_begin(self, **kwargs) -> None:
        self._reset(**kwargs)

    def on_epoch_begin(self, **kwargs) -> None:
        self._reset(**kwargs)

    def _process(self, sentence: List[int]) -> List[str]:
        symbols = self.alphabet.get_symbols(sentence)
        return symbols if self.word_segmentor is None else self.word_segmentor(symbols)

    def on_batch_end(self, **kwargs) -> None:
        if self.training:
            return
        transcripts = kwargs["reports"][self._name]["transcripts"]
        targets, target_lens = kwargs["last_target"][0], kwargs["last_target"][1]
        acts = self.rnnt_decoder(*kwargs["last_output"])
        for act, target, target_len in zip(acts, targets, target_lens):
            act = self._process(act)
            exp = self._process([int(e) for e in target[: int(target_len)]])
            transcripts.append((act, exp))
            self.distances.append(_levenshtein(act, exp))
            self.lengths.append(len(exp))

    def on_epoch_end(self, **kwargs) -> None:
        if self.training:
            return
        wer = float(sum(self.distances)) / max(1, sum(self.lengths)) * 100
        kwargs["reports"][self._name]["wer"] = wer


def _levenshtein(a, b) -> int:
    """Edit distance between two sequences (``post_process/utils.py:4``)."""
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]

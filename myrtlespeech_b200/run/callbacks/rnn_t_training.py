"""Callbacks that let an RNN-T model run under the reference's ``fit`` loop unchanged.

The loop passes only ``x`` to the model (``run/train.py:62-63``: ``out, _ = seq_to_seq.model(x)``) and
``(out, y)`` to the loss (``:66-67``), but an RNN-T model also needs the labels for its prediction network.
The reference's hook for rewriting the model input is ``Callback.on_batch_begin``: the handler stores ``x`` /
``y`` as ``last_input`` / ``last_target`` in its ``state_dict``, calls every callback with the state as keyword
arguments, merges any returned dict back, and hands the possibly modified values to the loop
(``run/callbacks/callback.py:221-254``).  The classes here follow that protocol by duck typing (same method names,
``**kwargs`` in, ``Optional[dict]`` out), so they can be passed in ``fit(..., callbacks=[...])`` as they are or
subclass the reference's ``Callback`` when it is importable.
"""
from typing import Dict, List, Optional


class RNNTTraining:
    """Packs the targets into the model input at the start of every batch.

    ``x = (feats, feat_lens)`` and ``y = (labels, label_lens)`` (the collate layout of
    ``data/batch.py:45-107``) become ``last_input = ((feats, labels), (feat_lens, label_lens))``, the input
    :py:class:`myrtlespeech_b200.model.RNNT` expects; ``last_target`` is left as it is for the loss.
    """

    def on_batch_begin(self, **kwargs) -> Optional[Dict]:
        (feats, feat_lens), (labels, label_lens) = kwargs["last_input"], kwargs["last_target"]
        return {"last_input": ((feats, labels), (feat_lens, label_lens))}

    def __getattr__(self, name):
        # every other hook of the reference's Callback protocol is a no-op
        if name.startswith("on_"):
            return lambda **kwargs: None
        raise AttributeError(name)


class ReportRNNTDecoder:
    """Decodes every evaluation batch with an :py:class:`RNNTGreedyDecoder` and keeps transcripts and word errors,
    the RNN-T counterpart of ``ReportCTCDecoder`` (``run/run.py:50-109``, decoder call at ``:94``).

    Args:
        decoder: ``decoder(feats, feat_lens) -> List[List[int]]``.
        alphabet: maps ids to symbols with ``get_symbols`` (``data/alphabet.py:5``); ``None`` keeps ids.
    """

    def __init__(self, decoder, alphabet=None):
        self.decoder = decoder
        self.alphabet = alphabet
        self.training = True
        self.hypotheses: List = []
        self.references: List = []

    def train(self, mode: bool = True):
        self.training = mode

    def on_epoch_begin(self, **kwargs) -> None:
        self.hypotheses, self.references = [], []

    def on_batch_end(self, **kwargs) -> None:
        if self.training:
            return
        (feats, _labels), (feat_lens, _label_lens) = kwargs["last_input"]
        labels, label_lens = kwargs["last_target"]
        hyps = self.decoder(feats, feat_lens)
        refs = [labels[b, : int(label_lens[b])].tolist() for b in range(len(hyps))]
        if self.alphabet is not None:
            hyps = [self.alphabet.get_symbols(h) for h in hyps]
            refs = [self.alphabet.get_symbols(r) for r in refs]
        self.hypotheses.extend(hyps)
        self.references.extend(refs)

    def on_epoch_end(self, **kwargs) -> Optional[Dict]:
        if self.training or not self.references:
            return None
        errs = sum(_levenshtein(h, r) for h, r in zip(self.hypotheses, self.references))
        total = max(1, sum(len(r) for r in self.references))
        reports = dict(kwargs.get("reports", {}))
        reports[self.decoder.__class__.__name__ + "/error_rate"] = errs / total
        return {"reports": reports}

    def __getattr__(self, name):
        if name.startswith("on_"):
            return lambda **kwargs: None
        raise AttributeError(name)


def _levenshtein(a, b) -> int:
    """Edit distance between two sequences (``post_process/utils.py:4``)."""
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]

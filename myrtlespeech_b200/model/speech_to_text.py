"""``SeqToSeq`` / ``SpeechToText`` containers with the surface the reference's loop and ``Saver`` use.

``run/train.py:41-91`` needs ``seq_to_seq.train(mode)``, ``.model``, ``.loss``, ``.optim`` and ``.lr_scheduler``;
``Saver`` (``run/run.py:172-185``) stores ``seq_to_seq.state_dict()``, whose keys therefore start with ``model.``
(``model.joint.fc.weight`` ...); ``builders/task_config.py:69-100`` builds the optimiser from
``seq_to_seq.parameters()`` and assigns ``.optim`` / ``.lr_scheduler`` afterwards.  When myrtlespeech is importable
the classes below derive from its ``SeqToSeq`` (``model/seq_to_seq.py:10-61``) so ``isinstance`` checks hold; otherwise
an ``nn.Module`` with the same attributes, CUDA placement and ``pre_process`` property is used.
"""
import enum
from typing import Callable, Optional, Sequence, Tuple

import torch

try:  # pragma: no cover - depends on the environment
    from myrtlespeech.run.stage import Stage as Stage  # type: ignore
except Exception:
    from ..protos import stage_pb2 as _stage_pb2

    class Stage(enum.Enum):  # type: ignore[no-redef]
        """``run/stage.py:6-9`` over this package's run-time ``Stage`` enum descriptor."""

        TRAIN = _stage_pb2.TRAIN
        EVAL = _stage_pb2.EVAL
        TRAIN_AND_EVAL = _stage_pb2.TRAIN_AND_EVAL

try:  # pragma: no cover - depends on the environment
    from myrtlespeech.model.seq_to_seq import SeqToSeq as _RefSeqToSeq  # type: ignore
except Exception:
    _RefSeqToSeq = None


if _RefSeqToSeq is not None:  # pragma: no cover - depends on the environment
    SeqToSeq = _RefSeqToSeq
else:

    class SeqToSeq(torch.nn.Module):  # type: ignore[no-redef]
        """A generic sequence-to-sequence model: ``model``, ``loss``, ``pre_process_steps`` and ``optim``.

        ``pre_process_steps`` is a sequence of ``(callable, Stage)``; :py:attr:`pre_process` applies the callables
        whose stage matches ``self.training``.  ``model`` moves to the GPU at construction when CUDA is available.
        """

        def __init__(self, model: torch.nn.Module, loss: torch.nn.Module,
                     pre_process_steps: Sequence[Tuple[Callable, Stage]],
                     optim: Optional[torch.optim.Optimizer] = None):
            super().__init__()
            self.model = model
            self.loss = loss
            self.pre_process_steps = pre_process_steps
            self.optim = optim

            self.use_cuda = torch.cuda.is_available()
            if self.use_cuda:
                self.model = self.model.cuda()

        @property
        def pre_process(self) -> Callable:
            def process(x):
                for step, stage in self.pre_process_steps:
                    if stage is Stage.TRAIN and not self.training:
                        continue
                    if stage is Stage.EVAL and self.training:
                        continue
                    x = step(x)
                return x

            return process


class SpeechToText(SeqToSeq):
    """A :py:class:`SeqToSeq` for speech recognition: adds ``alphabet`` and ``post_process``
    (``model/speech_to_text.py:9-36``).  ``lr_scheduler`` starts as :py:data:`None` -- the reference assigns it in
    ``builders/task_config.py:98`` and ``fit`` reads it at ``run/train.py:85``."""

    def __init__(self, alphabet, post_process, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alphabet = alphabet
        self.post_process = post_process
        self.lr_scheduler = None

    def extra_repr(self) -> str:
        return f"(alphabet): {self.alphabet}"

"""The callback base class the RNN-T callbacks derive from.

When myrtlespeech itself is importable the base **is** the reference's ``Callback``
(``run/callbacks/callback.py:10-71``), so the classes here are genuine subclasses and anything in the reference
that checks ``isinstance(cb, Callback)`` or calls ``cb.train(mode)`` (``CallbackHandler.train``,
``run/callbacks/callback.py:483-491``, called by ``fit`` at ``run/train.py:51``) works unchanged.  When it is not
(this repo used on its own, the GPU test box), a stand-in with the identical protocol is used: a ``training``
attribute, ``train(mode=True) -> self`` and one do-nothing method per hook of the training loop.
"""
from typing import Dict, Optional

try:  # pragma: no cover - depends on the environment
    from myrtlespeech.run.callbacks.callback import Callback as Callback  # type: ignore
    from myrtlespeech.run.callbacks.callback import ModelCallback as ModelCallback  # type: ignore

    REFERENCE_CALLBACK = True
except Exception:  # myrtlespeech is not installed
    REFERENCE_CALLBACK = False

    class Callback:  # type: ignore[no-redef]
        """Same surface as the reference's ``Callback``: all hooks do nothing, ``train`` flips ``training``."""

        def __init__(self, training: bool = True):
            self.training = training

        def on_train_begin(self, **kwargs) -> Optional[Dict]:
            ...

        def on_epoch_begin(self, **kwargs) -> Optional[Dict]:
            ...

        def on_batch_begin(self, **kwargs) -> Optional[Dict]:
            ...

        def on_loss_begin(self, **kwargs) -> Optional[Dict]:
            ...

        def on_backward_begin(self, **kwargs) -> Optional[Dict]:
            ...

        def on_backward_end(self, **kwargs) -> Optional[Dict]:
            ...

        def on_step_end(self, **kwargs) -> Optional[Dict]:
            ...

        def on_batch_end(self, **kwargs) -> Optional[Dict]:
            ...

        def on_epoch_end(self, **kwargs) -> Optional[Dict]:
            ...

        def on_train_end(self, **kwargs) -> Optional[Dict]:
            ...

        def train(self, mode=True):
            self.training = mode
            return self

    class ModelCallback(Callback):  # type: ignore[no-redef]
        """Callback with access to the model (``run/callbacks/callback.py:74-88``)."""

        def __init__(self, model, training: bool = True):
            super().__init__(training=training)
            self.model = model

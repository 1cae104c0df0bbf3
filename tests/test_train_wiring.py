"""an RNN-T ``SpeechToText`` runs under the reference's REAL training loop.

Where the reference tree exists (this container) the tests import ``fit`` and ``CallbackHandler`` from
``/root/reference/src`` (``run/train.py:13-91``, ``run/callbacks/callback.py:90-491``; see ``tests/conftest.py`` for the
one generated module they need) and the callbacks of this package are genuine subclasses of the reference's
``Callback``.  On the GPU box the reference does not exist, so ``_fit`` / ``_Handler`` below restate the loop -- and
``test_restated_loop_makes_the_same_calls_as_the_reference`` proves here, with a recording callback, that the
restatement makes exactly the reference's sequence of callback calls, ``train(mode)`` included.

The CPU tests cannot run the product loss or decoder (CUDA only, no fallback), so they wire in test doubles for those
two; the ``-m gpu`` tests run the real ``RNNTLoss`` / ``RNNTGreedyDecoder``.
"""
import io
import os
import subprocess
import sys
from contextlib import ExitStack

import pytest
import torch
from google.protobuf import text_format

from myrtlespeech_b200.builders import speech_to_text as stt_builder
from myrtlespeech_b200.model.rnn_t import JointHandle
from myrtlespeech_b200.protos import speech_to_text_pb2
from myrtlespeech_b200.run.callbacks import BF16MixedPrecision, ClipGradNorm, RNNTTraining, ReportRNNTDecoder
from myrtlespeech_b200.run.callbacks.callback import REFERENCE_CALLBACK, Callback, ModelCallback
from myrtlespeech_b200.run.callbacks.rnn_t_training import _levenshtein
from tests.conftest import HAVE_REFERENCE_LOOP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------------------------
# the loop: the reference's own when importable, otherwise a restatement
# ---------------------------------------------------------------------------------------------------------------------
class _Handler:
    """Restates ``CallbackHandler`` (``run/callbacks/callback.py:90-491``): state dict in, dict updates merged (unknown
    keys are an error), the documented keys set before each stage, counters advanced after it, ``train`` forwarded."""

    def __init__(self, callbacks=None, training=True):
        self.callbacks = callbacks if callbacks is not None else []
        self.state_dict = {}
        self.training = training

    def __call__(self, name):
        for cb in self.callbacks:
            new = getattr(cb, name)(**self.state_dict)
            if new is None:
                continue
            for k, v in new.items():
                if k not in self.state_dict:
                    raise Exception(f"{k} is not a valid key in CallbackHandler state.")
                self.state_dict[k] = v

    def on_train_begin(self, epochs):
        self.state_dict.update(dict(epoch=0, epochs=epochs, total_train_batches=0, epoch_batches=0, reports={}))
        self("on_train_begin")

    def on_epoch_begin(self):
        self.state_dict["epoch_batches"] = 0
        self("on_epoch_begin")

    def on_batch_begin(self, x, y):
        self.state_dict["last_input"], self.state_dict["last_target"] = x, y
        self("on_batch_begin")
        return self.state_dict["last_input"], self.state_dict["last_target"]

    def on_loss_begin(self, out, y):
        self.state_dict["last_output"], self.state_dict["last_target"] = out, y
        self.state_dict["loss"] = {"last_output": out, "last_target": y}
        self("on_loss_begin")
        return self.state_dict["loss"]["last_output"], self.state_dict["loss"]["last_target"]

    def on_backward_begin(self, loss):
        self.state_dict["skip_bwd"], self.state_dict["last_loss"] = False, loss
        self("on_backward_begin")
        return self.state_dict["last_loss"], self.state_dict["skip_bwd"]

    def on_backward_end(self):
        self.state_dict["skip_step"] = False
        self("on_backward_end")
        return self.state_dict["skip_step"]

    def on_step_end(self):
        self.state_dict["skip_zero"] = False
        self("on_step_end")
        return self.state_dict["skip_zero"]

    def on_batch_end(self):
        self.state_dict["stop_epoch"] = False
        self("on_batch_end")
        self.state_dict["epoch_batches"] += 1
        if self.training:
            self.state_dict["total_train_batches"] += 1
        return self.state_dict["stop_epoch"]

    def on_epoch_end(self):
        self.state_dict["stop_training"] = False
        self("on_epoch_end")
        if self.training:
            self.state_dict["epoch"] += 1
        return self.state_dict["stop_training"]

    def on_train_end(self):
        self("on_train_end")

    def train(self, mode=True):
        self.training = mode
        for cb in self.callbacks:
            cb.train(mode=mode)
        return self


def _fit(seq_to_seq, epochs, train_loader, eval_loader=None, callbacks=None, handler_cls=_Handler):
    """Restates ``fit`` (``run/train.py:41-91``) statement by statement."""
    cb_handler = handler_cls(callbacks)
    cb_handler.on_train_begin(epochs)
    for epoch in range(epochs):
        stages = ["train"]
        if eval_loader is not None:
            stages.append("eval")
            if epoch == 0:
                stages.insert(0, "eval")
        for stage in stages:
            is_training = stage == "train"
            seq_to_seq.train(mode=is_training)
            cb_handler.train(mode=is_training)
            cb_handler.on_epoch_begin()
            with ExitStack() as stack:
                if not is_training:
                    stack.enter_context(torch.no_grad())
                loader = train_loader if is_training else eval_loader
                for x, y in loader:
                    x, y = cb_handler.on_batch_begin(x, y)
                    out, _ = seq_to_seq.model(x)
                    loss_out, loss_y = cb_handler.on_loss_begin(out, y)
                    loss = seq_to_seq.loss(loss_out, loss_y)
                    loss, skip_bwd = cb_handler.on_backward_begin(loss)
                    if is_training:
                        if not skip_bwd:
                            loss.backward()
                        if seq_to_seq.optim is not None:
                            if not cb_handler.on_backward_end():
                                seq_to_seq.optim.step()
                            if not cb_handler.on_step_end():
                                seq_to_seq.optim.zero_grad()
                    if cb_handler.on_batch_end():
                        break
                if is_training and seq_to_seq.lr_scheduler is not None:
                    seq_to_seq.lr_scheduler.step()
            if cb_handler.on_epoch_end():
                break
    cb_handler.on_train_end()


def _loop():
    """(fit, CallbackHandler): the reference's own objects when its tree is importable."""
    if HAVE_REFERENCE_LOOP:
        from myrtlespeech.run.callbacks.callback import CallbackHandler
        from myrtlespeech.run.train import fit
        return fit, CallbackHandler
    return _fit, _Handler


# ---------------------------------------------------------------------------------------------------------------------
# fixtures: config, batches in the reference's collate layout, CPU test doubles
# ---------------------------------------------------------------------------------------------------------------------
CFG = """
alphabet: "abcdefg_";
pre_process_step { stage: TRAIN_AND_EVAL; mfcc { n_mfcc: 6; win_length: 400; hop_length: 160; } }
pre_process_step { stage: TRAIN_AND_EVAL; context_frames { n_context: 1; } }
rnn_t { encoder_hidden_size: 16; encoder_num_layers: 1; pred_embedding_size: 8;
        pred_hidden_size: 16; pred_num_layers: 1; joint_hidden_size: 32; }
rnn_t_loss { blank_index: 7; reduction: SUM; }
rnn_t_greedy_decoder { blank_index: 7; max_symbols_per_step: 2; }
"""
N_MFCC, N_CH = 6, 3   # input_features, input_channels = 2 * n_context + 1 implied by CFG


def _build():
    return stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))


def _batches(n, seed=0):
    """``((inputs (B, C, F, T), in_lens), (targets (B, U), target_lens))`` as ``seq_to_seq_collate_fn`` returns them
    (``data/batch.py:45-107``): sequence axis last, sorted by length, int32 targets, int64 lengths."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        B, T, U = 3, 11, 4
        x = torch.randn(B, N_CH, N_MFCC, T, generator=g)
        y = torch.randint(0, 7, (B, U), generator=g, dtype=torch.int32)
        out.append(((x, torch.tensor([11, 9, 7])), (y, torch.tensor([4, 3, 2]))))
    return out


class _TorchaudioLoss(torch.nn.Module):
    """CPU test double for ``RNNTLoss`` with its ``forward(inputs, targets)`` contract: materialises the lazy joint and
    calls ``torchaudio.functional.rnnt_loss`` (the independent implementation the oracle is pinned to)."""

    def __init__(self, blank):
        super().__init__()
        self.blank = blank

    def forward(self, inputs, targets):
        import torchaudio
        (x, x_lens), (y, y_lens) = inputs, targets
        assert isinstance(x, JointHandle)
        return torchaudio.functional.rnnt_loss(x.materialize().float(), y.int(), x_lens.int(), y_lens.int(),
                                               blank=self.blank, reduction="sum")


class _StubDecoder:
    """CPU test double for the decoder: called as ``decoder(*last_output)`` (``run/run.py:94``)."""

    def __init__(self):
        self.calls = []

    def __call__(self, x, lengths):
        assert isinstance(x, JointHandle) and x.f.dim() == 3
        self.calls.append((tuple(x.shape), lengths.tolist()))
        return [[0, 1, 2][: int(n) % 4] for n in lengths]


class _Saver(ModelCallback):
    """The reference's ``Saver`` (``run/run.py:172-185``; the module it lives in needs generated protos and cannot be
    imported): ``torch.save(self.model.state_dict(), ...)`` at the end of every training epoch -- into memory here."""

    def __init__(self, model):
        super().__init__(model)
        self.saved = []

    def on_epoch_end(self, **kwargs):
        if not self.training:
            return
        buf = io.BytesIO()
        torch.save(self.model.state_dict(), buf)
        self.saved.append(buf.getvalue())


class _Recorder(Callback):
    """Records every call the handler makes, ``train`` included."""

    def __init__(self):
        super().__init__()
        self.calls = []

    def train(self, mode=True):
        self.calls.append(("train", mode))
        return super().train(mode)


for _hook in ("on_train_begin", "on_epoch_begin", "on_batch_begin", "on_loss_begin", "on_backward_begin", "on_backward_end",
              "on_step_end", "on_batch_end", "on_epoch_end", "on_train_end"):
    def _make(hook):
        def method(self, **kwargs):
            self.calls.append((hook, sorted(kwargs), kwargs.get("epoch"), kwargs.get("epoch_batches"),
                               kwargs.get("total_train_batches")))
        return method
    setattr(_Recorder, _hook, _make(_hook))


def _cpu_stt():
    torch.manual_seed(0)
    stt = _build()
    assert not next(stt.model.parameters()).is_cuda or torch.cuda.is_available()
    stt.loss = _TorchaudioLoss(7)            # test double: the product loss is CUDA-only
    stt.optim = torch.optim.SGD(stt.parameters(), lr=1e-2)      # builders/task_config.py:69-95
    stt.lr_scheduler = torch.optim.lr_scheduler.StepLR(stt.optim, step_size=1, gamma=0.5)  # :98
    return stt


needs_cpu_model = pytest.mark.skipif(torch.cuda.is_available(), reason="the CPU doubles assume the model stays on the CPU")


# ---------------------------------------------------------------------------------------------------------------------
# tests
# ---------------------------------------------------------------------------------------------------------------------
def test_callbacks_are_reference_callbacks_when_the_reference_is_importable():
    assert REFERENCE_CALLBACK == HAVE_REFERENCE_LOOP
    for cb in (RNNTTraining(), ReportRNNTDecoder(_StubDecoder(), None), ClipGradNorm(torch.nn.Linear(1, 1), 1.0)):
        assert isinstance(cb, Callback)
        assert cb.training is True and cb.train(False) is cb and cb.training is False
    if HAVE_REFERENCE_LOOP:
        from myrtlespeech.run.callbacks.callback import Callback as RefCallback
        assert Callback is RefCallback


def test_container_is_a_seq_to_seq_module_with_the_saver_key_layout():
    """``Saver`` stores ``seq_to_seq.state_dict()`` (``run/run.py:172-185``): keys are ``model.<...>``."""
    stt = _build()
    assert isinstance(stt, torch.nn.Module)
    for name in ("model", "loss", "pre_process_steps", "optim", "alphabet", "post_process", "lr_scheduler", "pre_process"):
        assert hasattr(stt, name), name
    assert stt.optim is None and stt.lr_scheduler is None
    sd = stt.state_dict()
    assert "model.joint.fc.weight" in sd and "model.joint.fc.bias" in sd
    assert all(k.startswith("model.") for k in sd), [k for k in sd if not k.startswith("model.")]
    assert len(list(stt.parameters())) == len(list(stt.model.parameters()))   # the decoder does not re-register the model
    assert stt.train(False) is stt and not stt.model.training and stt.train(True).model.training
    if HAVE_REFERENCE_LOOP:
        from myrtlespeech.model.seq_to_seq import SeqToSeq
        assert isinstance(stt, SeqToSeq)
    # sizes derived from the pre-processing steps, as builders/speech_to_text.py:249-272
    assert stt.model.encoder.input_features == N_MFCC and stt.model.encoder.input_channels == N_CH
    assert len(stt.pre_process_steps) == 2


def test_rnnt_training_packs_labels_into_the_model_input():
    _, Handler = _loop()
    feats, feat_lens = torch.zeros(2, 1, 3, 5), torch.tensor([5, 4])
    labels, label_lens = torch.ones(2, 3, dtype=torch.int32), torch.tensor([3, 2])
    h = Handler([RNNTTraining()])
    x, y = h.on_batch_begin((feats, feat_lens), (labels, label_lens))
    (f2, l2), (fl2, ll2) = x
    assert f2 is feats and l2 is labels and fl2 is feat_lens and ll2 is label_lens
    assert y[0] is labels and y[1] is label_lens


def test_levenshtein():
    assert _levenshtein([1, 2, 3], [1, 3]) == 1
    assert _levenshtein([], [4, 5]) == 2
    assert _levenshtein([7], [7]) == 0


@needs_cpu_model
def test_fit_runs_train_and_eval_stages_with_the_rnnt_callbacks():
    """Two epochs of ``fit`` -- EVAL, TRAIN, EVAL, then TRAIN, EVAL (``run/train.py:43-48``) -- with the label-packing,
    report, clipping and saver callbacks; under the reference's own ``fit`` and ``CallbackHandler`` in this container."""
    fit, _ = _loop()
    stt = _cpu_stt()
    before = {k: v.clone() for k, v in stt.state_dict().items()}
    decoder = _StubDecoder()

    class WordSegmentor:           # run/run.py:29-47, with "_"... any separator: here symbols are scored as words of one
        def __call__(self, sentence):
            return ["".join(sentence)] if sentence else []

    report = ReportRNNTDecoder(decoder, stt.alphabet, WordSegmentor())
    clip = ClipGradNorm(stt, max_norm=5.0)
    saver = _Saver(stt)
    rec = _Recorder()
    fit(stt, 2, _batches(2), _batches(1, seed=9), callbacks=[RNNTTraining(), report, clip, saver, rec])

    # every stage switched the callbacks' mode through CallbackHandler.train (run/train.py:51)
    assert [c[1] for c in rec.calls if c[0] == "train"] == [False, True, False, True, False]
    # evaluation batches were decoded through decoder(*last_output), training batches were not
    assert len(decoder.calls) == 3 and all(shape[0] == 3 for shape, _ in decoder.calls)
    assert clip.last_norm is not None and clip.last_norm > 0
    # parameters moved, and the Saver snapshot of epoch 2 reloads strictly into a freshly built container
    after = stt.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)
    assert len(saver.saved) == 2
    fresh = _build()
    fresh.load_state_dict(torch.load(io.BytesIO(saver.saved[-1])), strict=True)
    for k, v in fresh.state_dict().items():
        assert torch.equal(v.cpu(), after[k].cpu()), k
    # StepLR stepped once per training epoch (run/train.py:85-86)
    assert abs(stt.optim.param_groups[0]["lr"] - 1e-2 * 0.25) < 1e-12
    # every parameter received a finite gradient during training (tests/model/test_deep_speech_1.py:85-112 pattern)
    stt.train(True)
    (x, y) = _batches(1)[0]
    out, _ = stt.model(((x[0], y[0]), (x[1], y[1])))
    stt.loss(out, y).backward()
    assert not [n for n, p in stt.model.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]


@needs_cpu_model
def test_report_layout_matches_report_ctc_decoder():
    """``reports[<decoder class>] = {"wer": percent, "transcripts": [(hyp, ref), ...]}`` (``run/run.py:66-109``)."""
    fit, Handler = _loop()
    stt = _cpu_stt()
    seen = {}

    class Peek(Callback):
        def on_epoch_end(self, **kwargs):
            if not self.training:
                seen.update({k: dict(v) for k, v in kwargs["reports"].items()})

    fit(stt, 1, _batches(1), _batches(2, seed=3), callbacks=[RNNTTraining(), ReportRNNTDecoder(_StubDecoder(), stt.alphabet), Peek()])
    rep = seen["_StubDecoder"]
    assert set(rep) == {"wer", "transcripts"} and len(rep["transcripts"]) == 6 and rep["wer"] >= 0.0
    hyp, ref = rep["transcripts"][0]
    assert all(isinstance(s, str) for s in hyp + ref)


@pytest.mark.skipif(not HAVE_REFERENCE_LOOP, reason="needs the reference tree")
@needs_cpu_model
def test_restated_loop_makes_the_same_calls_as_the_reference():
    """The restated ``_fit`` / ``_Handler`` (what the GPU box runs, where the reference tree does not exist) makes
    exactly the same sequence of callback calls, with the same state keys and counters, as the reference's ``fit``."""
    from myrtlespeech.run.train import fit as ref_fit
    traces = []
    for loop in (ref_fit, _fit):
        stt = _cpu_stt()
        rec = _Recorder()
        loop(stt, 2, _batches(2), _batches(1, seed=9), callbacks=[RNNTTraining(), rec])
        traces.append(rec.calls)
    assert traces[0] == traces[1]
    assert len(traces[0]) > 40


def test_stand_in_base_classes_without_the_reference():
    """The package used on its own (no myrtlespeech on the path, as on the GPU box): the stand-in ``Callback`` /
    ``SeqToSeq`` have the same surface.  Fresh interpreter, reference hidden."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from myrtlespeech_b200.run.callbacks.callback import REFERENCE_CALLBACK, Callback\n"
        "from myrtlespeech_b200.run.callbacks import RNNTTraining, ClipGradNorm\n"
        "from myrtlespeech_b200.model.speech_to_text import SeqToSeq, SpeechToText, Stage\n"
        "import torch\n"
        "assert not REFERENCE_CALLBACK and SeqToSeq.__module__.startswith('myrtlespeech_b200')\n"
        "cb = RNNTTraining(); assert isinstance(cb, Callback) and cb.train(False) is cb and not cb.training\n"
        "assert cb.on_epoch_end(reports={}) is None\n"
        "m = SpeechToText(alphabet=None, post_process=None, model=torch.nn.Linear(2, 2), loss=torch.nn.MSELoss(),\n"
        "                 pre_process_steps=[(lambda x: x + 1, Stage.TRAIN), (lambda x: x * 2, Stage.TRAIN_AND_EVAL)])\n"
        "assert m.pre_process(1) == 4 and m.train(False).pre_process(1) == 2 and m.optim is None and m.lr_scheduler is None\n"
        "assert sorted(m.state_dict()) == ['model.bias', 'model.weight']\n"
        "print('ok')\n" % ROOT
    )
    env = dict(os.environ, PYTHONPATH="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]


def test_clip_grad_norm_clips_the_models_own_parameters():
    lin = torch.nn.Linear(4, 4)
    lin(torch.ones(2, 4)).sum().mul(100).backward()
    clip = ClipGradNorm(lin, max_norm=1.0)
    clip.on_backward_end()
    total = torch.sqrt(sum((p.grad ** 2).sum() for p in lin.parameters()))
    assert clip.last_norm > 1.0 and float(total) <= 1.0 + 1e-4
    assert clip.on_batch_begin(last_input=None) is None  # other hooks are no-ops


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the real loss and decoder under the loop
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("mixed", [False, True], ids=["fp32", "bf16_autocast"])
def test_fit_on_gpu_with_the_product_loss_and_decoder(mixed):
    """``fit`` (the reference's when importable, else the restatement proven equal above) with the CUDA ``RNNTLoss`` and
    the one-launch ``RNNTGreedyDecoder``: two epochs with evaluation stages, a Saver-style round trip at the end."""
    fit, _ = _loop()
    torch.manual_seed(0)
    stt = _build()
    assert next(stt.model.parameters()).is_cuda
    stt.optim = torch.optim.SGD(stt.parameters(), lr=1e-2)
    stt.lr_scheduler = torch.optim.lr_scheduler.StepLR(stt.optim, step_size=1, gamma=0.5)
    before = {k: v.clone() for k, v in stt.state_dict().items()}
    report = ReportRNNTDecoder(stt.post_process, stt.alphabet)
    clip = ClipGradNorm(stt, max_norm=5.0)
    saver = _Saver(stt)
    rec = _Recorder()
    seen = {}

    class Peek(Callback):
        def on_backward_begin(self, **kwargs):
            seen.setdefault("losses", []).append(float(kwargs["last_loss"]))

        def on_epoch_end(self, **kwargs):
            if not self.training:
                seen["reports"] = {k: dict(v) for k, v in kwargs["reports"].items()}

    cbs = ([BF16MixedPrecision()] if mixed else []) + [RNNTTraining(), report, clip, saver, rec, Peek()]
    fit(stt, 2, _batches(2), _batches(1, seed=9), callbacks=cbs)
    assert [c[1] for c in rec.calls if c[0] == "train"] == [False, True, False, True, False]
    assert all(l > 0 and l == l for l in seen["losses"]) and len(seen["losses"]) == 7
    rep = seen["reports"]["RNNTGreedyDecoder"]
    assert len(rep["transcripts"]) == 3 and rep["wer"] >= 0.0
    after = stt.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)
    assert clip.last_norm is not None and clip.last_norm > 0
    fresh = _build()
    fresh.load_state_dict(torch.load(io.BytesIO(saver.saved[-1])), strict=True)
    assert "model.joint.fc.weight" in fresh.state_dict()
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, after[k]), k

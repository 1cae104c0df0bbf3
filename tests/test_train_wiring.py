"""an RNN-T SpeechToText runs under the reference's loop protocol (run/train.py:60-82) with
the label-packing callback; every parameter receives a gradient (the reference's model-test pattern,
tests/model/test_deep_speech_1.py:85-112)."""
import pytest
import torch
from google.protobuf import text_format

from myrtlespeech_b200.builders import speech_to_text as stt_builder
from myrtlespeech_b200.protos import speech_to_text_pb2
from myrtlespeech_b200.run.callbacks import BF16MixedPrecision, ClipGradNorm, RNNTTraining, ReportRNNTDecoder
from myrtlespeech_b200.run.callbacks.rnn_t_training import _levenshtein


class _Handler:
    """The part of run/callbacks/callback.py:90-254 the loop uses: state dict, kwargs in, dict updates merged."""

    def __init__(self, callbacks):
        self.callbacks = callbacks
        self.state_dict = {}

    def __call__(self, name):
        for cb in self.callbacks:
            upd = getattr(cb, name)(**self.state_dict)
            if upd:
                self.state_dict.update(upd)

    def on_batch_begin(self, x, y):
        self.state_dict["last_input"], self.state_dict["last_target"] = x, y
        self("on_batch_begin")
        return self.state_dict["last_input"], self.state_dict["last_target"]


def test_rnnt_training_packs_labels_into_the_model_input():
    feats, feat_lens = torch.zeros(2, 5, 3), torch.tensor([5, 4])
    labels, label_lens = torch.ones(2, 3, dtype=torch.int32), torch.tensor([3, 2])
    h = _Handler([RNNTTraining()])
    x, y = h.on_batch_begin((feats, feat_lens), (labels, label_lens))
    (f2, l2), (fl2, ll2) = x
    assert f2 is feats and l2 is labels and fl2 is feat_lens and ll2 is label_lens
    assert y[0] is labels and y[1] is label_lens
    h("on_epoch_end")  # unknown hooks are no-ops


def test_levenshtein():
    assert _levenshtein([1, 2, 3], [1, 3]) == 1
    assert _levenshtein([], [4, 5]) == 2
    assert _levenshtein([7], [7]) == 0


CFG = """
alphabet: "abcdefg_";
input_features: 6;
rnn_t { encoder_hidden_size: 16; encoder_num_layers: 1; pred_embedding_size: 8;
        pred_hidden_size: 16; pred_num_layers: 1; joint_hidden_size: 32; }
rnn_t_loss { blank_index: 7; reduction: SUM; }
rnn_t_greedy_decoder { blank_index: 7; max_symbols_per_step: 2; }
"""


@pytest.mark.gpu
def test_one_training_step_and_eval_decode_under_the_loop_protocol():
    torch.manual_seed(0)
    stt = stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))
    report = ReportRNNTDecoder(stt.post_process)
    handler = _Handler([RNNTTraining(), report])
    B, T, U = 3, 11, 4
    x = (torch.randn(B, T, 6), torch.tensor([11, 9, 7]))
    y = (torch.randint(0, 7, (B, U), dtype=torch.int32), torch.tensor([4, 3, 2]))

    # run/train.py:60-73
    xi, yi = handler.on_batch_begin(x, y)
    out, _ = stt.model(xi)
    loss = stt.loss(out, yi)
    loss.backward()
    assert torch.isfinite(loss) and float(loss) > 0
    missing = [n for n, p in stt.model.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not missing, missing
    assert any(float(p.grad.abs().sum()) > 0 for p in stt.model.joint.parameters())

    # evaluation stage: run/run.py:84-109
    report.train(False)
    handler("on_epoch_begin")
    with torch.no_grad():
        handler("on_batch_end")
    handler("on_epoch_end")
    assert len(report.hypotheses) == B and all(isinstance(h, list) for h in report.hypotheses)
    rate = handler.state_dict["reports"]["RNNTGreedyDecoder/error_rate"]
    assert 0.0 <= rate


def test_clip_grad_norm_clips_the_models_own_parameters():
    lin = torch.nn.Linear(4, 4)
    lin(torch.ones(2, 4)).sum().mul(100).backward()
    clip = ClipGradNorm(lin, max_norm=1.0)
    clip.on_backward_end()
    total = torch.sqrt(sum((p.grad ** 2).sum() for p in lin.parameters()))
    assert clip.last_norm > 1.0 and float(total) <= 1.0 + 1e-4
    clip.on_batch_begin(last_input=None)  # other hooks are no-ops


def test_checkpoint_keys_round_trip():
    """Saver stores the container's state_dict (run/run.py:172-185); the joint's parameters are
    ordinary ``fc.weight`` / ``fc.bias`` entries and reload with strict=True."""
    torch.manual_seed(1)
    a = stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))
    b = stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))
    sd = a.model.state_dict()
    assert "joint.fc.weight" in sd and "joint.fc.bias" in sd
    b.model.load_state_dict(sd, strict=True)
    for (ka, va), (kb, vb) in zip(a.model.state_dict().items(), b.model.state_dict().items()):
        assert ka == kb and torch.equal(va.cpu(), vb.cpu())


@pytest.mark.gpu
def test_bf16_autocast_training_step():
    torch.manual_seed(0)
    stt = stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))
    handler = _Handler([BF16MixedPrecision(), RNNTTraining(), ClipGradNorm(stt, max_norm=5.0)])
    B, T, U = 2, 9, 3
    x = (torch.randn(B, T, 6), torch.tensor([9, 7]))
    y = (torch.randint(0, 7, (B, U), dtype=torch.int32), torch.tensor([3, 2]))
    xi, yi = handler.on_batch_begin(x, y)
    assert xi[0][0].is_cuda
    out, _ = stt.model(xi)
    handler("on_loss_begin")
    loss = stt.loss(out, yi)
    loss.backward()
    handler("on_backward_end")
    assert torch.isfinite(loss) and float(loss) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in stt.model.parameters())
    assert handler.callbacks[2].last_norm is not None
    handler("on_batch_end")

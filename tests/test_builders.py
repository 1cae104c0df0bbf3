"""Builder round-trips and invalid-enum ValueErrors (reference pattern: tests/builders/test_ctc_loss.py:15-60,
tests/builders/test_speech_to_text.py:15-23)."""
import hypothesis.strategies as st
import pytest
import torch
from google.protobuf import text_format
from hypothesis import assume, given

from myrtlespeech_b200.builders import rnn_t, rnn_t_greedy_decoder, rnn_t_loss, speech_to_text
from myrtlespeech_b200.loss import RNNTLoss
from myrtlespeech_b200.model import RNNT, RNNTJoint
from myrtlespeech_b200.post_process import RNNTGreedyDecoder
from myrtlespeech_b200.protos import rnn_t_loss_pb2, rnn_t_pb2, speech_to_text_pb2
from tests.proto_strategies import rnn_t_greedy_decoders, rnn_t_losses, rnn_ts


@given(cfg=rnn_t_losses())
def test_build_returns_correct_rnn_t_loss(cfg):
    loss = rnn_t_loss.build(cfg)
    assert isinstance(loss, RNNTLoss)
    assert loss.blank == cfg.blank_index
    expected = {rnn_t_loss_pb2.RNNTLoss.NONE: "none", rnn_t_loss_pb2.RNNTLoss.MEAN: "mean",
                rnn_t_loss_pb2.RNNTLoss.SUM: "sum"}[cfg.reduction]
    assert loss.reduction == expected


@given(cfg=rnn_t_losses(), invalid_reduction=st.integers(0, 128))
def test_unknown_reduction_raises_value_error(cfg, invalid_reduction):
    assume(invalid_reduction not in rnn_t_loss_pb2.RNNTLoss.REDUCTION.values())
    cfg.reduction = invalid_reduction
    with pytest.raises(ValueError):
        rnn_t_loss.build(cfg)


@given(cfg=rnn_ts(), features=st.integers(1, 8), vocab=st.integers(2, 12))
def test_build_rnn_t_matches_cfg(cfg, features, vocab):
    model, out = rnn_t.build(cfg, input_features=features, output_features=vocab)
    assert isinstance(model, RNNT) and isinstance(model.joint, RNNTJoint)
    assert out == vocab
    assert model.joint.hidden_size == cfg.joint_hidden_size
    assert model.joint.fc.out_features == vocab
    assert model.prediction.rnn.num_layers == cfg.pred_num_layers
    assert model.prediction.rnn.hidden_size == cfg.pred_hidden_size
    assert model.encoder.rnn.num_layers == cfg.encoder_num_layers
    want = torch.nn.LSTM if cfg.rnn_type == rnn_t_pb2.RNNT.LSTM else torch.nn.GRU
    assert isinstance(model.prediction.rnn, want) and isinstance(model.encoder.rnn, want)


@given(cfg=rnn_ts(), invalid=st.integers(2, 64))
def test_unknown_rnn_type_raises_value_error(cfg, invalid):
    assume(invalid not in rnn_t_pb2.RNNT.RNN_TYPE.values())
    cfg.rnn_type = invalid
    with pytest.raises(ValueError):
        rnn_t.build(cfg, 4, 5)


@given(cfg=rnn_t_greedy_decoders())
def test_build_greedy_decoder_matches_cfg(cfg):
    model = torch.nn.Identity()
    dec = rnn_t_greedy_decoder.build(cfg, model)
    assert isinstance(dec, RNNTGreedyDecoder)
    assert dec.blank_index == cfg.blank_index
    assert dec.max_symbols_per_step == cfg.max_symbols_per_step
    assert dec.model is model
    assert list(dec.parameters()) == []  # the decoder does not own the model


def test_greedy_decoder_zero_symbols_raises():
    cfg = rnn_t_greedy_decoders().example()
    cfg.max_symbols_per_step = 0
    with pytest.raises(ValueError):
        rnn_t_greedy_decoder.build(cfg, torch.nn.Identity())


STT = '''
alphabet: "abcdefghijklmnopqrstuvwxyz '_";
input_features: 8;
rnn_t { encoder_hidden_size: 8; encoder_num_layers: 1; pred_embedding_size: 8;
        pred_hidden_size: 8; pred_num_layers: 1; joint_hidden_size: 16; }
rnn_t_loss { blank_index: %d; reduction: SUM; }
rnn_t_greedy_decoder { blank_index: %d; max_symbols_per_step: 4; }
'''


def test_speech_to_text_builds():
    stt = speech_to_text.build(text_format.Merge(STT % (28, 28), speech_to_text_pb2.SpeechToText()))
    assert len(stt.alphabet) == 29
    assert stt.model.joint.fc.out_features == 29
    assert stt.loss.blank == 28 and stt.loss.reduction == "sum"
    assert stt.post_process.blank_index == 28 and stt.post_process.model is stt.model
    assert stt.pre_process_steps == []
    assert (stt.model.encoder.input_features, stt.model.encoder.input_channels) == (8, 1)


def test_speech_to_text_blank_mismatch_raises():
    with pytest.raises(ValueError, match="must match"):
        speech_to_text.build(text_format.Merge(STT % (28, 27), speech_to_text_pb2.SpeechToText()))


def test_speech_to_text_blank_out_of_range_raises():
    with pytest.raises(ValueError, match="must be in"):
        speech_to_text.build(text_format.Merge(STT % (29, 29), speech_to_text_pb2.SpeechToText()))


@pytest.mark.parametrize("missing", ["rnn_t {", "rnn_t_loss {", "rnn_t_greedy_decoder {"])
def test_speech_to_text_missing_member_raises(missing):
    text = STT % (28, 28)
    start = text.index(missing)
    end = text.index("}", start) + 1
    cfg = text_format.Merge(text[:start] + text[end:], speech_to_text_pb2.SpeechToText())
    with pytest.raises(ValueError, match="not supported"):
        speech_to_text.build(cfg)


STT_STEPS = '''
alphabet: "abc_";
pre_process_step { stage: TRAIN_AND_EVAL; mfcc { n_mfcc: 13; win_length: 400; hop_length: 160; } }
pre_process_step { stage: TRAIN; spec_augment { feature_mask: 2; time_mask: 3; n_feature_masks: 1; n_time_masks: 1; } }
pre_process_step { stage: TRAIN_AND_EVAL; standardize { } }
pre_process_step { stage: TRAIN_AND_EVAL; context_frames { n_context: 2; } }
input_features: 99;
rnn_t { encoder_hidden_size: 8; encoder_num_layers: 1; pred_embedding_size: 8;
        pred_hidden_size: 8; pred_num_layers: 1; joint_hidden_size: 16; }
rnn_t_loss { blank_index: 3; reduction: SUM; }
rnn_t_greedy_decoder { blank_index: 3; max_symbols_per_step: 4; }
'''


def test_input_sizes_are_derived_from_the_pre_process_steps():
    """``builders/speech_to_text.py:249-272``: n_mfcc fixes the feature width (the ``input_features`` extension field is
    then ignored), context frames the channel count; the encoder takes the collate layout (B, C, F, T)."""
    stt = speech_to_text.build(text_format.Merge(STT_STEPS, speech_to_text_pb2.SpeechToText()))
    enc = stt.model.encoder
    assert (enc.input_features, enc.input_channels) == (13, 5)
    assert [str(stage).split(".")[-1] for _, stage in stt.pre_process_steps] == ["TRAIN_AND_EVAL", "TRAIN", "TRAIN_AND_EVAL",
                                                                                 "TRAIN_AND_EVAL"]
    x = torch.randn(2, 5, 13, 7, device=next(enc.parameters()).device)
    (out, lens), _ = enc((x, torch.tensor([7, 5])))
    assert tuple(out.shape) == (2, 7, 16) and lens.tolist() == [7, 5]
    with pytest.raises(ValueError, match="encoder input must have size"):
        enc((torch.randn(2, 7, 13 * 5), torch.tensor([7, 5])))


def test_proto_patch_applies_to_the_reference_schema(tmp_path):
    """``protos/speech_to_text_rnn_t.proto.patch`` is an applicable unified diff against the reference's
    ``speech_to_text.proto``, and the run-time descriptors agree with the patched text (names and field numbers)."""
    import os
    import re
    import shutil
    import subprocess
    ref = "/root/reference/src/myrtlespeech/protos/speech_to_text.proto"
    if not os.path.exists(ref) or shutil.which("patch") is None:
        pytest.skip("needs the reference tree and patch(1)")
    dst = tmp_path / "src" / "myrtlespeech" / "protos"
    dst.mkdir(parents=True)
    shutil.copy(ref, dst / "speech_to_text.proto")
    patch = os.path.join(os.path.dirname(speech_to_text.__file__), "..", "protos", "speech_to_text_rnn_t.proto.patch")
    r = subprocess.run(["patch", "-p1", "-i", os.path.abspath(patch)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    text = (dst / "speech_to_text.proto").read_text()
    fields = speech_to_text_pb2.SpeechToText.DESCRIPTOR.fields_by_name
    for name in ("alphabet", "pre_process_step", "rnn_t", "rnn_t_loss", "rnn_t_greedy_decoder", "input_features"):
        m = re.search(r"\b%s = (\d+);" % name, text)
        assert m and int(m.group(1)) == fields[name].number, name
    for name, oneof in (("rnn_t", "supported_models"), ("rnn_t_loss", "supported_losses"),
                        ("rnn_t_greedy_decoder", "supported_post_processes")):
        assert fields[name].containing_oneof.name == oneof

"""Protobuf schema for the RNN-T components, built at import time.

The reference compiles ``src/myrtlespeech/protos/*.proto`` with ``protoc`` at
install (``Dockerfile:28``).  ``protoc`` is not available here, so the same
proto3 messages are declared as ``FileDescriptorProto``s and registered in a
private ``DescriptorPool``; the resulting classes behave exactly like generated
``*_pb2`` classes (``text_format.Merge``, ``WhichOneof``, nested enums,
``DESCRIPTOR.fields_by_name`` -- everything the reference's builders and tests
use, e.g. ``tests/protos/utils.py:19``, ``tests/builders/test_ctc_loss.py:57``).

Messages (package ``myrtlespeech.protos``), each modelled on its CTC analog:

* ``RNNTLoss``            <- ``protos/ctc_loss.proto:6-24``
* ``RNNTGreedyDecoder``   <- ``protos/ctc_greedy_decoder.proto:5-8``
* ``RNNT``                <- ``protos/deep_speech_2.proto`` / ``rnn.proto`` style model message
* ``Stage``, ``PreProcessStep`` (+ ``MFCC``, ``SpecAugment``, ``Standardize``, ``ContextFrames``)
                          <- ``protos/stage.proto:6-10``, ``protos/pre_process_step.proto:9-63``, same field numbers
* ``SpeechToText``        <- ``protos/speech_to_text.proto:13-35``: ``alphabet = 1`` and ``pre_process_step = 2`` as in
                          the reference, plus the three new oneof members (8, 9, 10).  The reference's own members
                          (``deep_speech_1/2``, ``ctc_*``: numbers 3-7) are not redeclared here -- they belong to
                          models outside this path -- so a config that uses them is rejected by the parser, and a
                          serialized RNN-T config is wire-compatible with the patched reference message
                          (``speech_to_text_rnn_t.proto.patch``).

The equivalent ``.proto`` text is kept next to this file (``*.proto``) for a
maintainer who wants to drop it into the reference tree.
"""
from types import SimpleNamespace

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_POOL = descriptor_pool.DescriptorPool()
_PKG = "myrtlespeech.protos"


def _field(msg, name, number, ftype, type_name=None, oneof_index=None, label=_F.LABEL_OPTIONAL):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name is not None:
        f.type_name = type_name
    if oneof_index is not None:
        f.oneof_index = oneof_index
    return f


def _file(name, deps=()):
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = f"myrtlespeech/protos/{name}.proto"
    fd.package = _PKG
    fd.syntax = "proto3"
    for d in deps:
        fd.dependency.append(f"myrtlespeech/protos/{d}.proto")
    return fd


def _build():
    # ---- rnn_t_loss.proto --------------------------------------------------------------
    fd = _file("rnn_t_loss")
    m = fd.message_type.add()
    m.name = "RNNTLoss"
    e = m.enum_type.add()
    e.name = "REDUCTION"
    for i, n in enumerate(("NONE", "MEAN", "SUM")):
        v = e.value.add()
        v.name, v.number = n, i
    _field(m, "blank_index", 1, _F.TYPE_UINT32)
    _field(m, "reduction", 2, _F.TYPE_ENUM, f".{_PKG}.RNNTLoss.REDUCTION")
    _POOL.Add(fd)

    # ---- rnn_t_greedy_decoder.proto ----------------------------------------------------
    fd = _file("rnn_t_greedy_decoder")
    m = fd.message_type.add()
    m.name = "RNNTGreedyDecoder"
    _field(m, "blank_index", 1, _F.TYPE_UINT32)
    _field(m, "max_symbols_per_step", 2, _F.TYPE_UINT32)
    _POOL.Add(fd)

    # ---- rnn_t.proto -------------------------------------------------------------------
    fd = _file("rnn_t")
    m = fd.message_type.add()
    m.name = "RNNT"
    e = m.enum_type.add()
    e.name = "RNN_TYPE"
    for i, n in enumerate(("LSTM", "GRU")):
        v = e.value.add()
        v.name, v.number = n, i
    _field(m, "rnn_type", 1, _F.TYPE_ENUM, f".{_PKG}.RNNT.RNN_TYPE")
    _field(m, "encoder_hidden_size", 2, _F.TYPE_UINT32)
    _field(m, "encoder_num_layers", 3, _F.TYPE_UINT32)
    _field(m, "pred_embedding_size", 4, _F.TYPE_UINT32)
    _field(m, "pred_hidden_size", 5, _F.TYPE_UINT32)
    _field(m, "pred_num_layers", 6, _F.TYPE_UINT32)
    _field(m, "joint_hidden_size", 7, _F.TYPE_UINT32)
    _POOL.Add(fd)

    # ---- stage.proto / pre_process_step.proto (schema identical to the reference's) ------
    fd = _file("stage")
    e = fd.enum_type.add()
    e.name = "Stage"
    for i, n in enumerate(("TRAIN", "EVAL", "TRAIN_AND_EVAL")):
        v = e.value.add()
        v.name, v.number = n, i
    _POOL.Add(fd)

    fd = _file("pre_process_step", deps=("stage",))
    m = fd.message_type.add()
    m.name = "PreProcessStep"
    m.oneof_decl.add().name = "pre_process_step"
    _field(m, "stage", 1, _F.TYPE_ENUM, f".{_PKG}.Stage")
    _field(m, "mfcc", 2, _F.TYPE_MESSAGE, f".{_PKG}.MFCC", oneof_index=0)
    _field(m, "standardize", 3, _F.TYPE_MESSAGE, f".{_PKG}.Standardize", oneof_index=0)
    _field(m, "context_frames", 4, _F.TYPE_MESSAGE, f".{_PKG}.ContextFrames", oneof_index=0)
    _field(m, "spec_augment", 5, _F.TYPE_MESSAGE, f".{_PKG}.SpecAugment", oneof_index=0)
    m = fd.message_type.add()
    m.name = "MFCC"
    _field(m, "n_mfcc", 1, _F.TYPE_UINT32)
    _field(m, "win_length", 2, _F.TYPE_UINT32)
    _field(m, "hop_length", 3, _F.TYPE_UINT32)
    _field(m, "legacy", 4, _F.TYPE_BOOL)
    m = fd.message_type.add()
    m.name = "SpecAugment"
    for i, n in enumerate(("feature_mask", "time_mask", "n_feature_masks", "n_time_masks"), 1):
        _field(m, n, i, _F.TYPE_UINT32)
    fd.message_type.add().name = "Standardize"
    m = fd.message_type.add()
    m.name = "ContextFrames"
    _field(m, "n_context", 1, _F.TYPE_UINT32)
    _POOL.Add(fd)

    # ---- speech_to_text.proto: the reference's message with the RNN-T oneof members added ----
    fd = _file("speech_to_text", deps=("pre_process_step", "rnn_t", "rnn_t_loss", "rnn_t_greedy_decoder"))
    m = fd.message_type.add()
    m.name = "SpeechToText"
    for n in ("supported_models", "supported_losses", "supported_post_processes"):
        m.oneof_decl.add().name = n
    _field(m, "alphabet", 1, _F.TYPE_STRING)
    _field(m, "pre_process_step", 2, _F.TYPE_MESSAGE, f".{_PKG}.PreProcessStep", label=_F.LABEL_REPEATED)
    # numbers 3-7 are the reference's deep_speech_1/2, ctc_loss and ctc decoders
    _field(m, "rnn_t", 8, _F.TYPE_MESSAGE, f".{_PKG}.RNNT", oneof_index=0)
    _field(m, "rnn_t_loss", 9, _F.TYPE_MESSAGE, f".{_PKG}.RNNTLoss", oneof_index=1)
    _field(m, "rnn_t_greedy_decoder", 10, _F.TYPE_MESSAGE, f".{_PKG}.RNNTGreedyDecoder", oneof_index=2)
    # Extension for configs without an MFCC step (features computed outside the config, e.g. the synthetic batches of
    # the tests): width of the feature axis.  Ignored when a pre_process_step fixes it.
    _field(m, "input_features", 11, _F.TYPE_UINT32)
    _POOL.Add(fd)

    def cls(name):
        return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(f"{_PKG}.{name}"))

    stage = _POOL.FindEnumTypeByName(f"{_PKG}.Stage")
    return (cls("RNNTLoss"), cls("RNNTGreedyDecoder"), cls("RNNT"), cls("SpeechToText"), cls("PreProcessStep"),
            SimpleNamespace(**{v.name: v.number for v in stage.values}, DESCRIPTOR=stage))


RNNTLoss, RNNTGreedyDecoder, RNNT, SpeechToText, PreProcessStep, _Stage = _build()

# ``from myrtlespeech_b200.protos import rnn_t_loss_pb2`` mirrors the reference's generated modules
rnn_t_loss_pb2 = SimpleNamespace(RNNTLoss=RNNTLoss)
rnn_t_greedy_decoder_pb2 = SimpleNamespace(RNNTGreedyDecoder=RNNTGreedyDecoder)
rnn_t_pb2 = SimpleNamespace(RNNT=RNNT)
speech_to_text_pb2 = SimpleNamespace(SpeechToText=SpeechToText)
pre_process_step_pb2 = SimpleNamespace(PreProcessStep=PreProcessStep)
stage_pb2 = SimpleNamespace(Stage=_Stage, TRAIN=_Stage.TRAIN, EVAL=_Stage.EVAL, TRAIN_AND_EVAL=_Stage.TRAIN_AND_EVAL)

__all__ = ["rnn_t_loss_pb2", "rnn_t_greedy_decoder_pb2", "rnn_t_pb2", "speech_to_text_pb2", "pre_process_step_pb2",
           "stage_pb2"]

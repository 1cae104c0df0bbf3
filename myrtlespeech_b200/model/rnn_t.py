"""RNN-T model pieces: the joint network (the hot path) and thin torch containers around it.

Follows the reference's module conventions: modules take and return ``(tensor,
lens)`` tuples (``model/fully_connected.py:133-166``), a model returns
``((out, out_lens), hidden)`` (``model/deep_speech_2.py:143-172``) so that the
train loop's ``out, _ = seq_to_seq.model(x)`` (``run/train.py:63``) works, and
parameters move to the GPU at construction when CUDA is available
(``model/fully_connected.py:103-105``).
"""
from typing import Optional, Tuple

import torch


class JointHandle:
    """Lazy joint output: everything needed to evaluate ``W . tanh(f_t + g_u) + b`` on demand.

    ``RNNTJoint.forward`` returns this in place of the ``(B, T, U+1, V)`` logits so that
    :py:class:`myrtlespeech_b200.loss.RNNTLoss` can run the fused CUDA path that never materialises
    them.  ``materialize()`` builds the dense tensor with plain torch ops (small shapes / debugging).
    """

    def __init__(self, f: torch.Tensor, g: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        if f.dim() != 3 or g.dim() != 3 or f.size(0) != g.size(0) or f.size(2) != g.size(2):
            raise ValueError(f"f {tuple(f.shape)} and g {tuple(g.shape)} must be (B,T,H) and (B,U+1,H)")
        if weight.dim() != 2 or weight.size(1) != f.size(2):
            raise ValueError(f"weight {tuple(weight.shape)} must be (V, {f.size(2)})")
        self.f, self.g, self.weight, self.bias = f, g, weight, bias

    @property
    def shape(self) -> Tuple[int, int, int, int]:
        return (self.f.size(0), self.f.size(1), self.g.size(1), self.weight.size(0))

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self) -> torch.device:
        return self.f.device

    def materialize(self) -> torch.Tensor:
        h = torch.tanh(self.f.unsqueeze(2) + self.g.unsqueeze(1))
        return torch.nn.functional.linear(h, self.weight, self.bias)

    def __repr__(self) -> str:
        return f"JointHandle(shape={self.shape}, dtype={self.f.dtype}, device={self.f.device})"


class RNNTJoint(torch.nn.Module):
    r"""Additive joint network :math:`z_{t,u} = W \tanh(f_t + g_u) + b`.

    Args:
        hidden_size: joint width ``H`` (size of ``f_t`` and ``g_u``).
        out_features: vocabulary size ``V`` including the blank.
        bias: whether the output projection has a bias.
        lazy: if :py:data:`True` (default) ``forward`` returns a :py:class:`JointHandle` for the
            fused loss; if :py:data:`False` it returns the dense ``(B, T, U+1, V)`` logits.
    """

    def __init__(self, hidden_size: int, out_features: int, bias: bool = True, lazy: bool = True):
        super().__init__()
        self.hidden_size = hidden_size
        self.out_features = out_features
        self.lazy = lazy
        self.fc = torch.nn.Linear(hidden_size, out_features, bias=bias)
        self.use_cuda = torch.cuda.is_available()
        if self.use_cuda:
            self.fc = self.fc.cuda()

    def forward(
        self, f: Tuple[torch.Tensor, torch.Tensor], g: Tuple[torch.Tensor, torch.Tensor]
    ) -> Tuple[object, torch.Tensor]:
        """``f = (B,T,H) encoder output, lens``; ``g = (B,U+1,H) prediction output, lens``.

        Returns ``(handle_or_logits, f_lens)``; the second element is the length of the *time* axis,
        as every reference module returns the lengths of its output sequence.
        """
        (fx, f_lens), (gx, _g_lens) = f, g
        if self.use_cuda:
            fx, gx = fx.cuda(), gx.cuda()
        handle = JointHandle(fx, gx, self.fc.weight, self.fc.bias)
        return (handle if self.lazy else handle.materialize()), f_lens

    def extra_repr(self) -> str:
        return f"lazy={self.lazy}"


class RNNTPredictionNet(torch.nn.Module):
    """Embedding + RNN over the label history, projected to the joint width.

    Input row 0 is the start-of-sequence step (a learned zero-initialised embedding index ``V``),
    so ``U`` labels produce ``U + 1`` outputs.
    """

    def __init__(self, vocab_size: int, embedding_size: int, hidden_size: int, num_layers: int,
                 joint_hidden_size: int, rnn_type: str = "lstm"):
        super().__init__()
        self.vocab_size = vocab_size
        self.embedding = torch.nn.Embedding(vocab_size + 1, embedding_size)
        rnn_cls = {"lstm": torch.nn.LSTM, "gru": torch.nn.GRU}[rnn_type]
        self.rnn = rnn_cls(embedding_size, hidden_size, num_layers=num_layers, batch_first=True)
        self.proj = torch.nn.Linear(hidden_size, joint_hidden_size)
        self.use_cuda = torch.cuda.is_available()
        if self.use_cuda:
            self.cuda()

    def forward(self, y: Tuple[torch.Tensor, torch.Tensor], hx=None):
        labels, lens = y
        if self.use_cuda:
            labels = labels.cuda()
        sos = torch.full((labels.size(0), 1), self.vocab_size, dtype=torch.long, device=labels.device)
        inp = torch.cat([sos, labels.long()], dim=1)
        out, hid = self.rnn(self.embedding(inp), hx)
        return (self.proj(out), lens + 1), hid

    def step(self, label: Optional[torch.Tensor], hx, batch: int, device):
        """One decoding step.  ``label`` (B,) long or None for start-of-sequence."""
        if label is None:
            label = torch.full((batch,), self.vocab_size, dtype=torch.long, device=device)
        out, hid = self.rnn(self.embedding(label.long().unsqueeze(1)), hx)
        return self.proj(out[:, 0]), hid


class RNNT(torch.nn.Module):
    """Encoder + prediction network + joint.

    ``forward(x)`` takes ``x = ((audio_feats, labels), (audio_lens, label_lens))`` -- the labels are
    packed into the model input because the reference's loop passes only ``x`` to the model
    (``run/train.py:62-63``) -- and returns ``((joint_out, out_lens), hidden)``.

    ``audio_feats`` has the layout the reference's collate function produces, ``(batch, channels, features,
    seq_len)`` with the sequence axis last (``data/batch.py:45-107``, the input contract of
    ``model/deep_speech_2.py:143-172``); the encoder owns the transposition to its own layout.
    """

    def __init__(self, encoder: torch.nn.Module, prediction: RNNTPredictionNet, joint: RNNTJoint):
        super().__init__()
        self.encoder = encoder
        self.prediction = prediction
        self.joint = joint

    def encode(self, feats: torch.Tensor, lens: torch.Tensor):
        out = self.encoder((feats, lens))
        # reference encoders return ((out, lens), hidden); plain modules may return (out, lens)
        if isinstance(out[0], tuple):
            return out[0]
        return out

    def forward(self, x):
        (feats, labels), (feat_lens, label_lens) = x
        f = self.encode(feats, feat_lens)
        g, hid = self.prediction((labels, label_lens))
        return self.joint(f, g), hid


class _LinearEncoder(torch.nn.Module):
    """Minimal (features -> RNN -> joint width) encoder used by the builder.

    Input: ``(x, lens)`` with ``x`` of size ``(batch, channels, features, seq_len)`` -- the reference's model input
    layout (``data/batch.py:45-107``; ``model/deep_speech_1.py`` / ``deep_speech_2.py`` take the same).  Channels and
    features are flattened to one axis of width ``input_channels * input_features`` and the sequence axis is moved
    in front of it for the batch-first RNN.  Output: ``((batch, seq_len, joint_hidden_size), lens)``.
    """

    def __init__(self, input_features: int, hidden_size: int, num_layers: int, joint_hidden_size: int,
                 rnn_type: str = "lstm", input_channels: int = 1):
        super().__init__()
        rnn_cls = {"lstm": torch.nn.LSTM, "gru": torch.nn.GRU}[rnn_type]
        self.input_features = input_features
        self.input_channels = input_channels
        self.rnn = rnn_cls(input_channels * input_features, hidden_size, num_layers=num_layers, batch_first=True)
        self.proj = torch.nn.Linear(hidden_size, joint_hidden_size)
        self.use_cuda = torch.cuda.is_available()
        if self.use_cuda:
            self.cuda()

    def forward(self, x):
        feats, lens = x
        if feats.dim() != 4 or feats.size(1) != self.input_channels or feats.size(2) != self.input_features:
            raise ValueError(f"encoder input must have size (batch, {self.input_channels}, {self.input_features}, "
                             f"seq_len), got {tuple(feats.shape)}")
        if self.use_cuda:
            feats = feats.cuda()
        b, c, n, t = feats.shape
        out, hid = self.rnn(feats.reshape(b, c * n, t).transpose(1, 2))
        return (self.proj(out), lens), hid

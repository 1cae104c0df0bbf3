"""Host-side logic that runs without a GPU: argument validation, the lazy joint handle, the C-ABI
library's symbol table and its pure-host entry points."""
import ctypes
import os
import re

import hypothesis.strategies as st
import pytest
import torch
from hypothesis import assume, given

import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from myrtlespeech_b200.loss import RNNTLoss
from myrtlespeech_b200.model import JointHandle, RNNTJoint
from myrtlespeech_b200.post_process import RNNTGreedyDecoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ---------------------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rnnt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rnnt_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), "ctypes signatures out of sync with include/rnnt_b200.h"


def test_abi_version_and_workspace_arithmetic():
    lib = _lib.load()
    assert lib.rnnt_abi_version() == 3
    small = lib.rnnt_fused_workspace_bytes(4, 200, 50, 29, 512)
    big = lib.rnnt_fused_workspace_bytes(32, 500, 100, 1024, 1024)
    assert 0 < small < big < 1 << 30
    # unsupported shapes are reported, not silently accepted
    assert lib.rnnt_fused_workspace_bytes(4, 200, 50, 8192, 512) > 0            # word-piece vocabularies up to 8192
    assert lib.rnnt_fused_workspace_bytes(4, 200, 50, 8193, 512) == 0
    assert b"V=8193" in lib.rnnt_last_error()
    assert lib.rnnt_fused_workspace_bytes(4, 200, 4095, 29, 512) > 0            # up to 4096 lattice columns
    assert lib.rnnt_fused_workspace_bytes(4, 200, 4096, 29, 512) == 0
    assert lib.rnnt_fused_workspace_bytes(4, 200, 50, 29, 20) == 0
    assert lib.rnnt_lattice_workspace_bytes(4, 200, 50) > 0


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    lib = _lib.load()
    rc = lib.rnnt_fused_forward(None, None, None, None, None, None, None, 1, 4, 2, 5, 8, 0, None, None, 0, None)
    assert rc == 1 and b"NULL" in lib.rnnt_last_error()
    rc = lib.rnnt_fused_forward_keep(None, None, None, None, None, None, None, 1, 4, 2, 5, 8, 0, None, None, 0, None, 0, None)
    assert rc == 1 and b"NULL" in lib.rnnt_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_decode_abi_workspace_arithmetic_and_argument_checks():
    """The one-launch decode entry points: workspace sizes are pure host arithmetic, unsupported shapes report 0 and
    bad arguments are rejected before any CUDA call."""
    lib = _lib.load()
    one = lib.rnnt_greedy_decode_workspace_bytes(128, 1024, 1024, 512)
    two = lib.rnnt_greedy_decode_stack_workspace_bytes(128, 1024, 1024, 512, 2)
    assert 0 < one < two < 1 << 28
    assert lib.rnnt_greedy_decode_stack_workspace_bytes(128, 1024, 1024, 512, 1) == one
    assert lib.rnnt_greedy_decode_workspace_bytes(128, 1024, 1024, 500) == 0        # Hp % 8
    assert lib.rnnt_greedy_decode_workspace_bytes(128, 1024, 1020, 512) == 0        # H % 8
    assert lib.rnnt_greedy_decode_stack_workspace_bytes(128, 1024, 1024, 512, 4) == 0   # more than three layers
    rc = lib.rnnt_greedy_decode_lstm(None, None, None, None, None, None, None, None, 4, 10, 29, 64, 64, 28, 2, None, 20,
                                     None, None, 0, None)
    assert rc == 1 and b"NULL" in lib.rnnt_last_error()
    dummy = ctypes.c_void_p(16)   # never dereferenced: the range checks come first
    rc = lib.rnnt_greedy_decode_lstm_stack(dummy, dummy, dummy, None, dummy, dummy, 1, None, None, dummy, None, 4, 10, 29,
                                           64, 64, 29, 2, dummy, 20, dummy, dummy, 1 << 20, None)
    assert rc == 1 and b"blank=29" in lib.rnnt_last_error()


def test_product_path_fails_loudly_on_cpu_tensors():
    f = torch.zeros(1, 2, 8); g = torch.zeros(1, 3, 8); W = torch.zeros(5, 8)
    with pytest.raises(_lib.RNNTLibraryError, match="no CPU fallback"):
        M.rnnt_joint_loss(f, g, W, None, torch.zeros(1, 2, dtype=torch.int32), [2], [2], 0)
    with pytest.raises(_lib.RNNTLibraryError, match="no CPU fallback"):
        M.rnnt_loss_from_logits(torch.zeros(1, 2, 3, 5), torch.zeros(1, 2, dtype=torch.int32), [2], [2], 0)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "myrtlespeech_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, fn)).read(), os.path.join(dirpath, fn)


# ---- joint handle --------------------------------------------------------------------------------
@given(B=st.integers(1, 3), T=st.integers(1, 5), U=st.integers(0, 4), V=st.integers(2, 6), H=st.integers(1, 8))
def test_joint_handle_shape_and_materialize(B, T, U, V, H):
    joint = RNNTJoint(H, V)
    joint.use_cuda = False
    joint = joint.cpu()
    f = torch.randn(B, T, H); g = torch.randn(B, U + 1, H)
    (handle, lens) = joint((f, torch.full((B,), T)), (g, torch.full((B,), U + 1)))
    assert isinstance(handle, JointHandle)
    assert tuple(handle.shape) == (B, T, U + 1, V)
    assert lens.tolist() == [T] * B
    z = handle.materialize()
    want = torch.tanh(f[:, :, None] + g[:, None]) @ joint.fc.weight.T + joint.fc.bias
    assert torch.allclose(z, want, atol=1e-6)
    joint.lazy = False
    dense, _ = joint((f, lens), (g, lens))
    assert torch.equal(dense, z)


def test_joint_handle_rejects_mismatched_shapes():
    with pytest.raises(ValueError):
        JointHandle(torch.zeros(2, 3, 8), torch.zeros(1, 3, 8), torch.zeros(5, 8), None)
    with pytest.raises(ValueError):
        JointHandle(torch.zeros(2, 3, 8), torch.zeros(2, 3, 8), torch.zeros(5, 4), None)


def test_every_joint_parameter_is_reachable_from_the_handle():
    joint = RNNTJoint(8, 5)
    joint.use_cuda = False
    joint = joint.cpu()
    h, _ = joint((torch.randn(1, 2, 8), torch.tensor([2])), (torch.randn(1, 3, 8), torch.tensor([3])))
    h.materialize().sum().backward()
    assert all(p.grad is not None for p in joint.parameters())


# ---- loss argument validation (reference pattern: tests/post_process/test_ctc_greedy_decoder.py:105-152) ----
def _loss_inputs(B=2, T=4, U=3, V=5):
    x = torch.zeros(B, T, U + 1, V)
    return (x, torch.full((B,), T)), (torch.ones(B, U, dtype=torch.int32), torch.full((B,), U))


def test_loss_unknown_reduction_raises():
    with pytest.raises(ValueError):
        RNNTLoss(blank=0, reduction="median")


@pytest.mark.parametrize("dtype", [torch.half, torch.float, torch.double])
def test_loss_float_lengths_raise(dtype):
    inputs, targets = _loss_inputs()
    with pytest.raises(ValueError):
        RNNTLoss(0, "sum")((inputs[0], inputs[1].to(dtype)), targets)
    with pytest.raises(ValueError):
        RNNTLoss(0, "sum")(inputs, (targets[0], targets[1].to(dtype)))


@given(xb=st.integers(1, 6), lb=st.integers(1, 6))
def test_loss_batch_mismatch_raises(xb, lb):
    assume(xb != lb)
    inputs, targets = _loss_inputs(B=xb)
    with pytest.raises(ValueError):
        RNNTLoss(0, "sum")((inputs[0], torch.ones(lb, dtype=torch.int64)), targets)


def test_loss_length_overflow_and_blank_range_raise():
    inputs, targets = _loss_inputs()
    with pytest.raises(ValueError):
        RNNTLoss(0, "sum")((inputs[0], torch.tensor([4, 5])), targets)
    with pytest.raises(ValueError):
        RNNTLoss(0, "sum")(inputs, (targets[0], torch.tensor([3, 4])))
    with pytest.raises(ValueError):
        RNNTLoss(5, "sum")(inputs, targets)
    with pytest.raises(ValueError):
        RNNTLoss(0, "sum")(inputs, (torch.ones(2, 2, dtype=torch.int32), targets[1]))


# ---- decoder argument validation -----------------------------------------------------------------
class _FakeModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.joint = RNNTJoint(8, 5)


@pytest.mark.parametrize("dtype", [torch.half, torch.float, torch.double])
def test_decoder_float_lengths_raise(dtype):
    dec = RNNTGreedyDecoder(0, _FakeModel())
    with pytest.raises(ValueError):
        dec(torch.zeros(2, 3, 8), torch.tensor([3, 3]).to(dtype))


@given(xb=st.integers(1, 8), lb=st.integers(1, 8))
def test_decoder_batch_mismatch_raises(xb, lb):
    assume(xb != lb)
    dec = RNNTGreedyDecoder(0, _FakeModel())
    with pytest.raises(ValueError):
        dec(torch.zeros(xb, 3, 8), torch.ones(lb, dtype=torch.int16))


def test_decoder_length_overflow_raises_and_repr():
    dec = RNNTGreedyDecoder(3, _FakeModel(), max_symbols_per_step=2)
    with pytest.raises(ValueError):
        dec(torch.zeros(2, 3, 8), torch.tensor([3, 4]))
    assert "blank_index=3" in repr(dec) and "max_symbols_per_step=2" in repr(dec)
    with pytest.raises(ValueError):
        RNNTGreedyDecoder(0, _FakeModel(), max_symbols_per_step=0)


def test_decoder_tells_inputs_apart_by_type_and_rank():
    """JointHandle / (B, T, H) encoder output / (B, C, F, T) model input in the reference's collate layout
    (data/batch.py:45-107); never by comparing a size with the joint width."""
    dec = RNNTGreedyDecoder(3, _FakeModel())
    # rank 3 whose last axis is not the joint width is an error, not "raw features"
    with pytest.raises(ValueError, match=r"encoder output must have size \(batch, seq_len, 8\)"):
        dec(torch.zeros(2, 3, 5), torch.tensor([3, 3]))
    with pytest.raises(ValueError, match="x must be"):
        dec(torch.zeros(2, 3), torch.tensor([3, 3]))
    # rank 4: lengths are checked against the LAST axis, and the model's encoder is applied
    with pytest.raises(ValueError, match="less than or equal to x seq_len"):
        dec(torch.zeros(2, 1, 8, 3), torch.tensor([3, 4]))
    with pytest.raises(AttributeError):      # _FakeModel has no encoder: proves the rank-4 branch calls model.encode
        dec(torch.zeros(2, 1, 8, 3), torch.tensor([3, 3]))
    # a JointHandle carries the encoder output; the length check runs against its time axis
    h = JointHandle(torch.zeros(2, 3, 8), torch.zeros(2, 2, 8), torch.zeros(5, 8), None)
    with pytest.raises(ValueError, match="less than or equal to x seq_len"):
        dec(h, torch.tensor([3, 4]))


def test_custom_op_is_registered_with_a_fake_implementation():
    """``torch.ops.rnnt_b200.fused_joint_loss`` / ``..._backward`` exist, and their fake (meta) implementations infer the
    output shapes without a GPU -- what ``torch.compile`` / ``torch.export`` need to trace through the ctypes call."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from myrtlespeech_b200 import functional as F
    lib = _lib.load()
    for dims in [(32, 500, 100, 1024, 1024), (3, 37, 11, 300, 128), (1, 1, 0, 3, 8), (2, 6, 1100, 12, 8), (4, 200, 50, 29, 512)]:
        assert lib.rnnt_fused_state_bytes(*dims) == F._state_bytes(*dims), dims
        assert 0 < lib.rnnt_fused_state_bytes(*dims) < lib.rnnt_fused_workspace_bytes(*dims)
        assert lib.rnnt_fused_kept_bytes(*dims) == F._kept_bytes(*dims) > 0, dims
    assert lib.rnnt_fused_kept_bytes(32, 500, 100, 1024, 1024) == 2 * 13312 * 128 * (1024 + 1024)     # 6.98 GB at the target shape
    assert lib.rnnt_fused_kept_bytes(4, 200, 50, 5000, 512) == F._kept_bytes(4, 200, 50, 5000, 512) == 0
    F.set_keep_activations(False)
    try:
        assert F._kept_bytes(32, 500, 100, 1024, 1024) == 0
    finally:
        F.set_keep_activations(True)
    cap = F._keep_cap
    F.set_keep_activations(True, max_bytes=1 << 30)          # a budget per live graph: beyond it the activations are recomputed
    try:
        assert F._kept_bytes(32, 500, 100, 1024, 1024) == 0 and F._kept_bytes(4, 200, 50, 29, 512) > 0
    finally:
        F.set_keep_activations(True, max_bytes=cap)
    with FakeTensorMode():
        f = torch.empty(3, 37, 128, device="cuda"); g = torch.empty(3, 12, 128, device="cuda")
        W = torch.empty(300, 128, device="cuda"); b = torch.empty(300, device="cuda")
        y = torch.empty(3, 11, dtype=torch.int32, device="cuda")
        fl = torch.empty(3, dtype=torch.int64); yl = torch.empty(3, dtype=torch.int64)
        loss, state, logits = torch.ops.rnnt_b200.fused_joint_loss(f, g, W, b, y, fl, yl, 299)
        assert tuple(loss.shape) == (3,) and loss.dtype == torch.float32 and loss.device.type == "cuda"
        assert state.dtype == torch.uint8 and state.numel() == F._state_bytes(3, 37, 11, 300, 128)
        assert logits.dtype == torch.uint8 and logits.numel() == F._kept_bytes(3, 37, 11, 300, 128)
        grads = torch.ops.rnnt_b200.fused_joint_loss_backward(loss, f, g, W, None, y, fl, yl, state, logits, 299)
        assert [tuple(t.shape) for t in grads] == [(3, 37, 128), (3, 12, 128), (300, 128), (300,)]

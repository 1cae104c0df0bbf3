"""GPU parity tests: the CUDA path (through the C-ABI library) against the CPU oracle.

Tolerance: BASELINE.json's ``north_star`` asks for loss and gradients within 1e-3 relative with fp32
accumulation.  ``rel`` below is ||got - ref||_F / ||ref||_F; ``relmax`` is max|got - ref| / max|ref|.
The oracle runs in its bf16-faithful mode (rounds h and dz where the kernels do), so the comparison
measures the kernels and not bf16 itself; a looser check against the exact fp64 oracle bounds the
total error including bf16.
"""
import json
import os

import numpy as np
import pytest
import torch

import myrtlespeech_b200 as M
from myrtlespeech_b200.loss import RNNTLoss
from myrtlespeech_b200.model import RNNTJoint
from oracle import rnnt_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["persistent", "recompute", "slab"], autouse=True)
def kernel_path(request):
    """Every test runs on the three schedules of the C-ABI library: the persistent kernels with the joint's activations
    kept for the backward pass (one forward launch, one backward mega-kernel; the default), the same kernels recomputing
    the activations in the backward pass, and the per-slab kernels (the general fallback)."""
    from myrtlespeech_b200 import _lib, functional as F
    lib = _lib.load()
    lib.rnnt_debug_set(b"path", 0 if request.param == "slab" else 1)
    F.set_keep_activations(request.param == "persistent")
    yield request.param
    lib.rnnt_debug_set(b"path", 1)
    F.set_keep_activations(True)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def bf16(x):
    return torch.tensor(np.asarray(x), dtype=torch.float32).bfloat16().float()


def make(seed, B, T, U, V, H, blank, ragged):
    rng = np.random.default_rng(seed)
    f = bf16(rng.normal(size=(B, T, H)))
    g = bf16(rng.normal(size=(B, U + 1, H)))
    W = bf16(rng.uniform(-1, 1, size=(V, H)) / np.sqrt(H))
    bias = torch.tensor(rng.uniform(-1, 1, size=V) / np.sqrt(H), dtype=torch.float32)
    labels = np.array([k for k in range(V) if k != blank])
    y = torch.tensor(rng.choice(labels, size=(B, max(U, 1)))[:, :U].reshape(B, U), dtype=torch.int32)
    fl = np.full(B, T); yl = np.full(B, U)
    if ragged and B > 1:
        fl = np.sort(rng.integers(max(1, T // 2), T + 1, size=B))[::-1].copy(); fl[0] = T
        yl = rng.integers(U // 2, U + 1, size=B); yl[0] = U
    return f, g, W, bias, y, fl.astype(np.int64), yl.astype(np.int64)


def run_cuda(f, g, W, bias, y, fl, yl, blank, grad_loss=None):
    fd = f.cuda().requires_grad_(True); gd = g.cuda().requires_grad_(True)
    Wd = W.cuda().requires_grad_(True)
    bd = None if bias is None else bias.cuda().requires_grad_(True)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), torch.tensor(fl), torch.tensor(yl), blank)
    gl = torch.ones_like(loss) if grad_loss is None else torch.tensor(grad_loss, dtype=torch.float32, device="cuda")
    loss.backward(gl)
    torch.cuda.synchronize()
    out = dict(loss=loss.detach().cpu().numpy(), df=fd.grad.cpu().numpy(), dg=gd.grad.cpu().numpy(),
               dW=Wd.grad.cpu().numpy())
    if bd is not None:
        out["db"] = bd.grad.cpu().numpy()
    return out


CASES = [
    # seed, B, T, U, V, H, blank, ragged
    (1, 1, 2, 2, 5, 8, 0, False),
    (2, 2, 5, 3, 6, 8, 5, False),
    (3, 3, 9, 4, 7, 16, 0, True),
    (4, 2, 12, 6, 29, 24, 28, True),
    (5, 2, 20, 9, 40, 72, 39, True),
    (6, 2, 37, 11, 300, 128, 299, True),       # two V chunks, second partly out of range
    (7, 1, 1, 0, 3, 8, 1, False),              # empty target, single frame
    (8, 3, 33, 0, 9, 16, 8, False),            # U = 0 for the whole batch
    (9, 2, 17, 15, 16, 8, 7, True),            # blank in the middle
    (10, 2, 6, 1100, 12, 8, 11, True),         # more than 1024 lattice columns (two columns per lattice thread)
]


@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "B%d_T%d_U%d_V%d_H%d" % c[1:6])
def test_fused_matches_oracle(cfg):
    f, g, W, bias, y, fl, yl = make(*cfg)
    blank = cfg[6]
    got = run_cuda(f, g, W, bias, y, fl, yl, blank)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, blank, faithful=True)
    exact = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, blank)
    assert rel(got["loss"], ref["loss"]) < TOL
    assert rel(got["loss"], exact["loss"]) < TOL
    for k in ("df", "dg", "dW", "db"):
        assert rel(got[k], ref[k]) < TOL, (k, rel(got[k], ref[k]))
        assert rel(got[k], exact[k]) < 2e-2, (k, rel(got[k], exact[k]))  # bf16 h / dz included
    # padded lattice rows get exactly zero gradient
    for b in range(cfg[1]):
        assert np.all(got["df"][b, fl[b]:] == 0)
        assert np.all(got["dg"][b, yl[b] + 1:] == 0)


def test_baseline_config_c1_matches_oracle():
    """BASELINE.json configs[0]: B=4 T=200 U=50 V=29 H=512 (the CPU-oracle-sized case), ragged."""
    cfg = (7, 4, 200, 50, 29, 512, 28, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    got = run_cuda(f, g, W, bias, y, fl, yl, 28)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, 28, faithful=True)
    assert rel(got["loss"], ref["loss"]) < 1e-5
    for k in ("df", "dg", "dW", "db"):
        assert rel(got[k], ref[k]) < TOL, (k, rel(got[k], ref[k]))


def test_subword_shape_multi_slab_matches_oracle():
    """V = H = 1024 (configs[2] widths) over more than one slab of tiles."""
    cfg = (8, 2, 60, 20, 1024, 1024, 1023, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    got = run_cuda(f, g, W, bias, y, fl, yl, 1023)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, 1023, faithful=True)
    assert rel(got["loss"], ref["loss"]) < 1e-5
    for k in ("df", "dg", "dW", "db"):
        assert rel(got[k], ref[k]) < TOL, (k, rel(got[k], ref[k]))


def test_ring_wraparound_many_tiles():
    """208 tiles = 104 pair-tiles: more than producers x ring slots (50 x 2) of the backward mega-kernel, so every
    producer reuses its slots and the consumers' `done` counters gate the reuse; ragged, V and H off the tile sizes."""
    cfg = (11, 2, 200, 60, 130, 136, 129, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    got = run_cuda(f, g, W, bias, y, fl, yl, 129)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, 129, faithful=True)
    assert rel(got["loss"], ref["loss"]) < 1e-5
    for k in ("df", "dg", "dW", "db"):
        assert rel(got[k], ref[k]) < TOL, (k, rel(got[k], ref[k]))


@pytest.mark.parametrize("shape", [
    # B, T, U, V, H: role splits of the backward mega-kernel other than the 8-block x 3-group one
    (2, 20, 5, 2048, 64),     # 8 V-blocks x 1 H-block (the widest vocabulary the backward mega-kernel takes)
    (2, 20, 5, 4096, 64),     # 16 V-blocks x 1 H-block: the widest vocabulary the backward mega-kernel takes (one K-group)
    (1, 24, 6, 3000, 512),    # 12 dW blocks, 12 V chunks with a partly filled last one, in the mega-kernel
    (2, 12, 4, 5000, 64),     # word-piece vocabulary beyond 4096: 20 V chunks; per-slab backward
    (1, 9, 3, 8192, 128),     # the largest supported vocabulary
    (2, 20, 5, 64, 1536),     # 1 V-block x 3 H-blocks, odd number of dW blocks (no consumer sharing)
    (1, 9, 3, 2048, 3072),    # 48 dW blocks: more consumers than the split allows -> per-slab kernels
], ids=lambda s: "V%d_H%d" % (s[3], s[4]))
def test_wide_shapes(shape):
    B, T, U, V, H = shape
    f, g, W, bias, y, fl, yl = make(21, B, T, U, V, H, V - 1, True)
    got = run_cuda(f, g, W, bias, y, fl, yl, V - 1)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, V - 1, faithful=True)
    assert rel(got["loss"], ref["loss"]) < 1e-5
    for k in ("df", "dg", "dW", "db"):
        assert rel(got[k], ref[k]) < TOL, (k, rel(got[k], ref[k]))


def test_repeated_calls_are_deterministic_in_loss_and_stable_in_grads():
    """Back-to-back steps on the same workspace size: loss bit-identical, gradients equal up to atomics order."""
    cfg = (12, 3, 50, 12, 70, 64, 69, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    a = run_cuda(f, g, W, bias, y, fl, yl, 69)
    b = run_cuda(f, g, W, bias, y, fl, yl, 69)
    assert np.array_equal(a["loss"], b["loss"])
    for k in ("df", "dg", "dW", "db"):
        assert rel(a[k], b[k]) < 1e-5, k


def test_golden_fixtures():
    """tests/golden/rnnt_small.json (torchaudio-checked).  Inputs there are not bf16-representable, so the
    CUDA path sees bf16-rounded copies; compare against the oracle on the same rounded inputs and,
    loosely, against the stored fp64 answers."""
    with open(os.path.join(GOLDEN, "rnnt_small.json")) as fh:
        G = json.load(fh)
    for case in G["cases"]:
        a = case["inputs"]
        H = len(a["f"][0][0])
        pad = (-H) % 8
        f = torch.nn.functional.pad(bf16(a["f"]), (0, pad)); g = torch.nn.functional.pad(bf16(a["g"]), (0, pad))
        W = torch.nn.functional.pad(bf16(a["W"]), (0, pad)); bias = torch.tensor(a["bias"], dtype=torch.float32)
        y = torch.tensor(a["y"], dtype=torch.int32)
        fl = np.array(a["f_lens"]); yl = np.array(a["y_lens"])
        got = run_cuda(f, g, W, bias, y, fl, yl, case["blank"])
        ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, case["blank"],
                                faithful=True)
        assert rel(got["loss"], ref["loss"]) < 1e-5
        assert rel(got["loss"], case["loss"]) < 5e-3          # bf16 input rounding only
        for k in ("df", "dg", "dW", "db"):
            assert rel(got[k], ref[k]) < TOL, k


def test_known_answer_vector_through_lattice_entry():
    z = np.array([.1, .6, .1, .1, .1, .1, .1, .6, .1, .1, .1, .1, .2, .8, .1,
                  .1, .6, .1, .1, .1, .1, .1, .2, .1, .1, .7, .1, .2, .1, .1]).reshape(1, 2, 3, 5)
    zt = torch.tensor(z, dtype=torch.float32, device="cuda", requires_grad=True)
    loss = M.rnnt_loss_from_logits(zt, torch.tensor([[1, 2]], dtype=torch.int32), torch.tensor([2]), torch.tensor([2]), 0)
    loss.sum().backward()
    assert abs(loss.item() - 4.495666) < 1e-5
    _, dz = O.rnnt_loss_from_logits(z, np.array([[1, 2]]), [2], [2], 0)
    assert np.allclose(zt.grad.cpu().numpy(), dz, atol=1e-6)


@pytest.mark.parametrize("shape", [(3, 7, 4, 6, 5), (4, 40, 17, 9, 8), (2, 33, 0, 4, 0), (2, 300, 120, 5, 4),
                                   (3, 130, 127, 4, 3), (2, 90, 128, 4, 0), (5, 1, 9, 3, 1),
                                   (2, 24, 1500, 4, 1), (2, 12, 2600, 3, 2)])   # > 1024 columns: 2 / 4 columns per thread
def test_lattice_entry_matches_oracle(shape):
    B, T, U, V, blank = shape
    rng = np.random.default_rng(B * 100 + T)
    z = rng.normal(size=(B, T, U + 1, V))
    labels = np.array([k for k in range(V) if k != blank])
    y = rng.choice(labels, size=(B, max(U, 1)))[:, :U].reshape(B, U)
    fl = rng.integers(max(1, T // 2), T + 1, size=B); fl[0] = T
    yl = rng.integers(U // 2, U + 1, size=B); yl[0] = U
    loss, dz = O.rnnt_loss_from_logits(z, y, fl, yl, blank)
    zt = torch.tensor(z, dtype=torch.float32, device="cuda", requires_grad=True)
    l = M.rnnt_loss_from_logits(zt, torch.tensor(y, dtype=torch.int32), torch.tensor(fl), torch.tensor(yl), blank)
    l.sum().backward()
    assert rel(l.detach().cpu().numpy(), loss) < 1e-5
    assert rel(zt.grad.cpu().numpy(), dz) < 1e-3


def test_linearity_in_grad_loss():
    """Size-independent property: the backward pass is linear in the upstream gradient."""
    cfg = (21, 3, 14, 5, 29, 64, 28, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    a = run_cuda(f, g, W, bias, y, fl, yl, 28, grad_loss=[1.0, 0.0, 0.0])
    b = run_cuda(f, g, W, bias, y, fl, yl, 28, grad_loss=[0.0, 2.0, -0.5])
    c = run_cuda(f, g, W, bias, y, fl, yl, 28, grad_loss=[1.0, 2.0, -0.5])
    for k in ("dW", "db", "df", "dg"):
        assert rel(a[k] + b[k], c[k]) < 2e-3, k
    assert np.all(a["df"][1:] == 0) and np.all(a["dg"][1:] == 0)  # utterances are independent


def test_module_surface_and_reductions():
    """RNNTJoint -> RNNTLoss.forward(inputs, targets), both the lazy handle and a dense logits tensor."""
    cfg = (22, 3, 11, 4, 12, 32, 11, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    joint = RNNTJoint(32, 12)
    with torch.no_grad():
        joint.fc.weight.copy_(W); joint.fc.bias.copy_(bias)
    f_lens = torch.tensor(fl); y_lens = torch.tensor(yl)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, 11, faithful=True)
    for reduction, want in (("none", ref["loss"]), ("sum", ref["loss"].sum()), ("mean", ref["loss"].mean())):
        loss_mod = RNNTLoss(blank=11, reduction=reduction)
        out = joint((f, f_lens), (g, y_lens + 1))
        val = loss_mod(out, (y, y_lens))
        assert rel(val.detach().cpu().numpy(), want) < 1e-5
        joint.lazy = False
        dense = joint((f, f_lens), (g, y_lens + 1))
        val2 = loss_mod(dense, (y, y_lens))
        assert rel(val2.detach().cpu().numpy(), want) < 1e-3
        joint.lazy = True
    # every joint parameter receives a gradient (reference pattern: tests/model/test_deep_speech_1.py:85-112)
    joint.zero_grad()
    RNNTLoss(11, "sum")(joint((f, f_lens), (g, y_lens + 1)), (y, y_lens)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in joint.parameters())
    assert rel(joint.fc.weight.grad.cpu().numpy(), ref["dW"]) < TOL


def test_full_size_properties():
    """BASELINE target shape (B=32 T=500 U=100 V=H=1024): too big for the CPU oracle, so check
    size-independent invariants: each softmax row's dz sums to zero => db sums to ~0 relative to its
    scale; loss is finite and positive; one utterance recomputed alone gives the same loss and grads."""
    B, T, U, V, H = 32, 500, 100, 1024, 1024
    f, g, W, bias, y, fl, yl = make(31, B, T, U, V, H, V - 1, False)
    got = run_cuda(f, g, W, bias, y, fl, yl, V - 1)
    assert np.all(np.isfinite(got["loss"])) and np.all(got["loss"] > 0)
    assert abs(got["db"].sum()) < 1e-3 * np.abs(got["db"]).sum()
    one = run_cuda(f[5:6], g[5:6], W, bias, y[5:6], fl[5:6], yl[5:6], V - 1)
    assert rel(one["loss"], got["loss"][5:6]) < 1e-6
    assert rel(one["df"], got["df"][5:6]) < 1e-5
    assert rel(one["dg"], got["dg"][5:6]) < 1e-5


def test_subword_width_against_torchaudio():
    """Independent implementation at V = H = 1024 (beyond what the numpy oracle checks cheaply): torch joint
    (Linear(tanh(f + g))) + ``torchaudio.functional.rnnt_loss`` with autograd, fp32, on the same bf16-representable
    inputs.  The torch path does not round h or dz to bf16, so the bound is the looser "includes bf16" one."""
    torchaudio = pytest.importorskip("torchaudio")
    B, T, U, V, H = 3, 48, 17, 1024, 1024
    f, g, W, bias, y, fl, yl = make(31, B, T, U, V, H, V - 1, True)
    got = run_cuda(f, g, W, bias, y, fl, yl, V - 1)
    dev = "cuda"
    ft, gt, Wt, bt = (x.clone().to(dev).requires_grad_(True) for x in (f, g, W, bias))
    logits = torch.nn.functional.linear(torch.tanh(ft.unsqueeze(2) + gt.unsqueeze(1)), Wt, bt)
    try:
        loss = torchaudio.functional.rnnt_loss(logits, y.to(dev), torch.tensor(fl, dtype=torch.int32, device=dev),
                                               torch.tensor(yl, dtype=torch.int32, device=dev), blank=V - 1,
                                               reduction="none")
        loss.sum().backward()
    except RuntimeError:   # torchaudio built without its CUDA transducer: use its CPU kernel
        dev = "cpu"
        ft, gt, Wt, bt = (x.clone().requires_grad_(True) for x in (f, g, W, bias))
        logits = torch.nn.functional.linear(torch.tanh(ft.unsqueeze(2) + gt.unsqueeze(1)), Wt, bt)
        loss = torchaudio.functional.rnnt_loss(logits, y, torch.tensor(fl, dtype=torch.int32),
                                               torch.tensor(yl, dtype=torch.int32), blank=V - 1, reduction="none")
        loss.sum().backward()
    assert rel(got["loss"], loss.detach().cpu().numpy()) < 5e-3
    for name, t in (("df", ft), ("dg", gt), ("dW", Wt), ("db", bt)):
        assert rel(got[name], t.grad.cpu().numpy()) < 2e-2, name


def test_custom_op_passes_opcheck(kernel_path):
    """``torch.library.opcheck`` on the registered operator: schema (no undeclared mutation or aliasing), autograd
    registration, fake-tensor implementation against the real one, and eager vs AOT-dispatched outputs and gradients with
    dynamic shapes.  Gradients are accumulated with fp32 atomics whose order differs run to run, hence the tolerances.
    opcheck compares every output, the opaque ``kept`` buffer included; its bytes are defined only where the forward
    pass writes (the tiles of the batch, the vocabulary columns that exist), so the ragged cases run with the activations
    recomputed (``kept`` is empty) and a dense batch with a 64-column vocabulary checks the kept schedule."""
    from myrtlespeech_b200 import functional as F
    cfg = (23, 3, 14, 5, 29, 64, 28, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    F.set_keep_activations(False)
    try:
        args = (f.cuda().requires_grad_(True), g.cuda().requires_grad_(True), W.cuda().requires_grad_(True),
                bias.cuda().requires_grad_(True), y.cuda(), torch.tensor(fl), torch.tensor(yl), 28)
        res = torch.library.opcheck(torch.ops.rnnt_b200.fused_joint_loss, args, atol=1e-5, rtol=1e-3)
        assert all(v == "SUCCESS" for v in res.values()), res
        # no bias; bf16 leaves (the gradients come back in the leaves' dtype)
        args = (f.bfloat16().cuda().requires_grad_(True), g.bfloat16().cuda().requires_grad_(True),
                W.bfloat16().cuda().requires_grad_(True), None, y.cuda(), torch.tensor(fl), torch.tensor(yl), 28)
        res = torch.library.opcheck(torch.ops.rnnt_b200.fused_joint_loss, args, atol=1e-2, rtol=2e-2)
        assert all(v == "SUCCESS" for v in res.values()), res
    finally:
        F.set_keep_activations(kernel_path == "persistent")
    if kernel_path == "persistent":
        f, g, W, bias, y, fl, yl = make(24, 2, 16, 7, 64, 64, 63, False)     # every tile and every column of `kept` is written
        args = (f.cuda().requires_grad_(True), g.cuda().requires_grad_(True), W.cuda().requires_grad_(True),
                bias.cuda().requires_grad_(True), y.cuda(), torch.tensor(fl), torch.tensor(yl), 63)
        res = torch.library.opcheck(torch.ops.rnnt_b200.fused_joint_loss, args, atol=1e-5, rtol=1e-3)
        assert all(v == "SUCCESS" for v in res.values()), res


def test_fused_op_under_torch_compile_matches_eager():
    """The operator is one opaque node for the compiler: a compiled function that wraps it gives the eager results."""
    cfg = (24, 2, 12, 4, 17, 32, 16, True)
    f, g, W, bias, y, fl, yl = make(*cfg)
    fd, gd, Wd, bd, yd = f.cuda(), g.cuda(), W.cuda(), bias.cuda(), y.cuda()
    flt, ylt = torch.tensor(fl), torch.tensor(yl)

    def fn(f_, g_, W_, b_):
        return M.rnnt_joint_loss(f_ * 1.0, g_, W_, b_, yd, flt, ylt, 16).sum()

    want = fn(fd, gd, Wd, bd)
    got = torch.compile(fn, backend="aot_eager")(fd, gd, Wd, bd)
    assert rel(got.detach().cpu().numpy(), want.detach().cpu().numpy()) < 1e-6


def test_two_live_graphs_share_one_scratch_workspace():
    """Two forward passes are alive at once (gradient accumulation over two graphs): each keeps only its ~state prefix,
    the scratch workspace is shared, and both backward passes give the right gradients."""
    a = make(31, 2, 20, 6, 29, 64, 28, True)
    b = make(32, 3, 15, 4, 29, 64, 28, True)
    outs = []
    leaves = []
    for f, g, W, bias, y, fl, yl in (a, b):
        fd = f.cuda().requires_grad_(True); gd = g.cuda().requires_grad_(True)
        Wd = W.cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
        outs.append(M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), torch.tensor(fl), torch.tensor(yl), 28).sum())
        leaves.append((fd, gd, Wd, bd))
    outs[0].backward()      # the second forward pass ran in between and overwrote the scratch
    outs[1].backward()
    for (f, g, W, bias, y, fl, yl), (fd, gd, Wd, bd) in zip((a, b), leaves):
        ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, 28, faithful=True)
        for k, t in (("df", fd), ("dg", gd), ("dW", Wd), ("db", bd)):
            assert rel(t.grad.cpu().numpy(), ref[k]) < TOL, k


def test_backward_walks_only_tiles_with_occupancy_and_loses_nothing(kernel_path):
    """The backward mega-kernel walks the compacted list of lattice tiles whose cells have non-zero arc occupancy; a tile
    whose occupancies all underflow to zero contributes exact zeros to every gradient.  Same gradients with the list and
    with every tile (up to the order of the fp32 atomics), fewer tiles walked, and an utterance whose upstream gradient
    is zero costs no tiles at all."""
    import ctypes
    from myrtlespeech_b200 import _lib, functional as F
    if kernel_path == "slab":
        pytest.skip("the tile list belongs to the persistent backward kernel")
    lib = _lib.load()
    B, T, U, V, H = 3, 300, 60, 64, 64
    f, g, W, bias, y, fl, yl = make(41, B, T, U, V, H, V - 1, False)
    gl = [1.0, 0.0, 0.5]
    out = {}
    counts = {}
    for mode in (0, -1):
        lib.rnnt_debug_set(b"prune", mode)
        try:
            out[mode] = run_cuda(f, g, W, bias, y, fl, yl, V - 1, grad_loss=gl)
            ws = next(iter(F._ws_pool.values()))
            n = (ctypes.c_int * 2)()
            if mode == 0:
                _lib.check(lib.rnnt_debug_read_active_tiles(ws.data_ptr(), B, T, U, V, H, n))
                counts = dict(active=n[0], total=n[1])
        finally:
            lib.rnnt_debug_set(b"prune", 0)
    per_utt = ((T + 15) // 16) * ((U + 1 + 7) // 8)
    assert counts["total"] == B * per_utt
    assert 0 < counts["active"] < 2 * per_utt, counts          # utterance 1 (grad 0) costs nothing; the others lose their corners
    for k in ("df", "dg", "dW", "db"):
        assert rel(out[0][k], out[-1][k]) < 1e-6, k
    assert np.all(out[0]["df"][1] == 0) and np.all(out[0]["dg"][1] == 0)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, V - 1, grad_loss=np.array(gl),
                            faithful=True)
    for k in ("df", "dg", "dW", "db"):
        assert rel(out[0][k], ref[k]) < TOL, k


@pytest.mark.parametrize("shape", [
    (3, 37, 11, 300, 128),      # vocabulary ends inside a 64-column box and inside a 256-column chunk
    (5, 70, 23, 1000, 520),     # H not a multiple of 64
    (1, 16, 7, 29, 512),        # a single tile: the pair-tile's second half is empty
    (3, 50, 20, 29, 512),       # narrow vocabulary: one logit box per tile, the box ring spans several tiles
    (2, 130, 40, 2048, 256),    # 32 boxes per tile
    (2, 40, 9, 4096, 64),       # widest vocabulary that is kept
])
def test_kept_activations_match_recompute(shape, kernel_path):
    """Keeping logits (fp16) and h (bf16) for the backward pass against recomputing them there: same loss bit for bit (the
    forward arithmetic is the same), gradients equal up to the fp16 rounding of the kept logits -- far inside the bf16
    rounding of dz both schedules share -- and both within tolerance of the oracle."""
    from myrtlespeech_b200 import _lib, functional as F
    if kernel_path != "persistent":
        pytest.skip("compares the two persistent schedules")
    B, T, U, V, H = shape
    f, g, W, bias, y, fl, yl = make(77, B, T, U, V, H, V - 1, True)
    gl = list(np.linspace(0.5, 1.5, B))
    lib = _lib.load()
    assert lib.rnnt_fused_kept_bytes(B, T, U, V, H) > 0
    out = {}
    for keep in (True, False):
        F.set_keep_activations(keep)
        try:
            out[keep] = run_cuda(f, g, W, bias, y, fl, yl, V - 1, grad_loss=gl)
        finally:
            F.set_keep_activations(True)
    assert np.array_equal(out[True]["loss"], out[False]["loss"])
    for k in ("df", "dg", "dW", "db"):
        assert rel(out[True][k], out[False][k]) < 2e-4, (k, rel(out[True][k], out[False][k]))
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, V - 1, grad_loss=np.array(gl),
                            faithful=True)
    for k in ("df", "dg", "dW", "db"):
        assert rel(out[True][k], ref[k]) < TOL, (k, rel(out[True][k], ref[k]))


def test_kept_buffer_is_optional_at_the_c_abi(kernel_path):
    """``rnnt_fused_forward_keep`` / ``rnnt_fused_backward_kept`` with a NULL buffer are ``rnnt_fused_forward`` /
    ``rnnt_fused_backward``; with a buffer the backward call reads what the forward call kept, and a negative upstream
    gradient flips every sign (the kept schedule carries the row scale in an exponent and the sign separately)."""
    from myrtlespeech_b200 import _lib
    if kernel_path != "persistent":
        pytest.skip("C-ABI check of the kept-activation entry points")
    lib = _lib.load()
    B, T, U, V, H = 2, 40, 12, 70, 64
    f, g, W, bias, y, fl, yl = make(5, B, T, U, V, H, V - 1, True)
    dev = torch.device("cuda")
    fb, gb, Wb = f.to(dev).bfloat16(), g.to(dev).bfloat16(), W.to(dev).bfloat16()
    bf_, yi = bias.to(dev), y.to(dev).int()
    fli, yli = torch.tensor(fl, dtype=torch.int32), torch.tensor(yl, dtype=torch.int32)
    nbytes = lib.rnnt_fused_workspace_bytes(B, T, U, V, H)
    kbytes = lib.rnnt_fused_kept_bytes(B, T, U, V, H)
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for name, kept, sign in (("null", None, 1.0), ("kept", torch.empty(kbytes, dtype=torch.uint8, device=dev), 1.0),
                             ("negative", torch.empty(kbytes, dtype=torch.uint8, device=dev), -1.0)):
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        loss = torch.empty(B, device=dev)
        kp, kn = (None, 0) if kept is None else (kept.data_ptr(), kbytes)
        _lib.check(lib.rnnt_fused_forward_keep(fb.data_ptr(), gb.data_ptr(), Wb.data_ptr(), bf_.data_ptr(), yi.data_ptr(),
                                               fli.data_ptr(), yli.data_ptr(), B, T, U, V, H, V - 1, loss.data_ptr(),
                                               ws.data_ptr(), nbytes, kp, kn, st))
        gl = torch.full((B,), sign, device=dev)
        df = torch.empty(B, T, H, device=dev); dg = torch.empty(B, U + 1, H, device=dev)
        dW = torch.empty(V, H, device=dev); db = torch.empty(V, device=dev)
        _lib.check(lib.rnnt_fused_backward_kept(fb.data_ptr(), gb.data_ptr(), Wb.data_ptr(), bf_.data_ptr(), yi.data_ptr(),
                                                fli.data_ptr(), yli.data_ptr(), B, T, U, V, H, V - 1, gl.data_ptr(),
                                                df.data_ptr(), dg.data_ptr(), dW.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                                nbytes, kp, kn, st))
        torch.cuda.synchronize()
        res[name] = [t.cpu().numpy() for t in (loss, df, dg, dW, db)]
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), fl, yl, V - 1, faithful=True)
    for name in ("null", "kept"):
        assert rel(res[name][0], ref["loss"]) < TOL
        for got, k in zip(res[name][1:], ("df", "dg", "dW", "db")):
            assert rel(got, ref[k]) < TOL, (name, k)
    for a, b in zip(res["kept"][1:], res["negative"][1:]):
        assert rel(-b, a) < 1e-6

"""Property-based differential test in the style of the reference's tests/loss/test_ctc_loss.py:75-108: random
shapes, blank index and ragged lengths; the CUDA path (both kernel schedules) must match the CPU oracle."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import myrtlespeech_b200 as M
from myrtlespeech_b200 import _lib
from oracle import rnnt_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@st.composite
def problems(draw):
    B = draw(st.integers(1, 4))
    T = draw(st.integers(1, 40))
    U = draw(st.integers(0, 12))
    V = draw(st.integers(2, 300))
    H = 8 * draw(st.integers(1, 40))
    blank = draw(st.integers(0, V - 1))
    seed = draw(st.integers(0, 2 ** 16))
    path = draw(st.sampled_from([0, 1]))
    rng = np.random.default_rng(seed)
    fl = np.sort(rng.integers(1, T + 1, size=B))[::-1].copy(); fl[0] = T
    yl = rng.integers(0, U + 1, size=B); yl[0] = U
    return dict(B=B, T=T, U=U, V=V, H=H, blank=blank, seed=seed, path=path, fl=fl.astype(np.int64), yl=yl.astype(np.int64))


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck))
@given(problems())
def test_random_problems_match_the_oracle(pr):
    B, T, U, V, H, blank = pr["B"], pr["T"], pr["U"], pr["V"], pr["H"], pr["blank"]
    rng = np.random.default_rng(pr["seed"])

    def bf(x):
        return torch.tensor(np.asarray(x), dtype=torch.float32).bfloat16().float()

    f, g = bf(rng.normal(size=(B, T, H))), bf(rng.normal(size=(B, U + 1, H)))
    W = bf(rng.uniform(-1, 1, size=(V, H)) / np.sqrt(H))
    bias = torch.tensor(rng.uniform(-1, 1, size=V) / np.sqrt(H), dtype=torch.float32)
    labels = np.array([k for k in range(V) if k != blank])
    y = torch.tensor(rng.choice(labels, size=(B, max(U, 1)))[:, :U].reshape(B, U), dtype=torch.int32)
    lib = _lib.load()
    lib.rnnt_debug_set(b"path", pr["path"])
    try:
        fd = f.cuda().requires_grad_(True); gd = g.cuda().requires_grad_(True)
        Wd = W.cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), torch.tensor(pr["fl"]), torch.tensor(pr["yl"]), blank)
        loss.sum().backward()
        torch.cuda.synchronize()
    finally:
        lib.rnnt_debug_set(b"path", 1)
    ref = O.rnnt_joint_loss(f.numpy(), g.numpy(), W.numpy(), bias.numpy(), y.numpy(), pr["fl"], pr["yl"], blank, faithful=True)
    assert rel(loss.detach().cpu().numpy(), ref["loss"]) < 1e-3
    for name, got in (("df", fd.grad), ("dg", gd.grad), ("dW", Wd.grad), ("db", bd.grad)):
        assert rel(got.cpu().numpy(), ref[name]) < 1e-3, (name, pr)


# --------------------------------------------------------------------------------------------------------------------
# one-launch greedy decode: random shapes, cell types, depths and both schedules against the oracle
# --------------------------------------------------------------------------------------------------------------------
@st.composite
def decode_problems(draw):
    cell = draw(st.sampled_from(["lstm", "gru"]))
    layers = draw(st.integers(1, 3))
    # the grid-barrier schedule covers one LSTM layer only
    variant = draw(st.sampled_from([0, 1])) if (cell == "lstm" and layers == 1) else 1
    return dict(B=draw(st.integers(1, 20)), T=draw(st.integers(1, 10)), V=draw(st.integers(2, 90)),
                H=8 * draw(st.integers(1, 20)), Hp=8 * draw(st.integers(1, 20)), E=draw(st.integers(1, 12)),
                S=draw(st.integers(1, 3)), blank=None, seed=draw(st.integers(0, 2 ** 16)), cell=cell, layers=layers,
                variant=variant)


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck))
@given(decode_problems())
def test_random_decode_problems_match_the_oracle(pr):
    from myrtlespeech_b200.model import RNNT, RNNTJoint, RNNTPredictionNet
    from myrtlespeech_b200.post_process import RNNTGreedyDecoder
    import myrtlespeech_b200.post_process.rnn_t_greedy_decoder as D

    B, T, V, H, Hp, E, S = pr["B"], pr["T"], pr["V"], pr["H"], pr["Hp"], pr["E"], pr["S"]
    torch.manual_seed(pr["seed"])
    blank = int(torch.randint(0, V, (1,)))          # any blank index, not only the last
    joint = RNNTJoint(H, V)
    pred = RNNTPredictionNet(V, E, Hp, pr["layers"], H, rnn_type=pr["cell"])
    with torch.no_grad():
        joint.fc.weight.mul_(4.0); pred.proj.weight.mul_(3.0); joint.fc.bias[blank] += 2.0
        for prm in list(joint.parameters()) + list(pred.parameters()):
            prm.copy_(prm.bfloat16().float())
    f = (torch.randn(B, T, H) * 1.5).bfloat16()
    lens = torch.randint(0, T + 1, (B,), dtype=torch.int32)
    n = lambda t: None if t is None else t.detach().cpu().float().numpy()  # noqa: E731
    r, L = pred.rnn, range(pr["layers"])
    make_step = O.lstm_pred_step if pr["cell"] == "lstm" else O.gru_pred_step
    step = make_step(n(pred.embedding.weight), [n(getattr(r, f"weight_ih_l{l}")) for l in L],
                     [n(getattr(r, f"weight_hh_l{l}")) for l in L], [n(getattr(r, f"bias_ih_l{l}")) for l in L],
                     [n(getattr(r, f"bias_hh_l{l}")) for l in L], n(pred.proj.weight), n(pred.proj.bias), faithful=True)
    want, margins = O.greedy_decode(f.float().numpy(), lens.numpy(), n(joint.fc.weight), n(joint.fc.bias), step, blank, S,
                                    faithful=True, per_utterance_margin=True)
    lib = _lib.load()
    lib.rnnt_debug_set(b"decode_variant", pr["variant"])
    calls, orig = [], D.greedy_decode_lstm
    D.greedy_decode_lstm = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        got = RNNTGreedyDecoder(blank, RNNT(torch.nn.Identity(), pred, joint).cuda(), max_symbols_per_step=S)(f.cuda(), lens)
    finally:
        D.greedy_decode_lstm = orig
        lib.rnnt_debug_set(b"decode_variant", 1)
    assert calls, ("the one-launch decode was not taken", pr)
    for g, w, m in zip(got, want, margins):
        if m > 5e-3:   # utterances whose smallest top-2 logit margin is within bf16 / tanh.approx noise are excused
            assert g == w, pr

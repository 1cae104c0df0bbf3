"""Pins the CPU oracle: known-answer vector, brute force, torchaudio, finite differences.

The reference has no RNN-T test to mirror; the style follows its
one loss test, a differential check with allclose (tests/loss/test_ctc_loss.py:75-108).
"""
import json
import os

import numpy as np
import pytest

from oracle import rnnt_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

KAT_LOGITS = np.array(
    [.1, .6, .1, .1, .1, .1, .1, .6, .1, .1, .1, .1, .2, .8, .1,
     .1, .6, .1, .1, .1, .1, .1, .2, .1, .1, .7, .1, .2, .1, .1]
).reshape(1, 2, 3, 5)
KAT_GRAD = np.array(
    [-.1312, -.3999, .1770, .1770, .1770, -.1857, .1225, -.1817, .1225, .1225,
     -.3209, .0627, .0693, .1262, .0627, .0546, -.2182, .0546, .0546, .0546,
     .1207, .1207, -.4830, .1207, .1207, -.6926, .1687, .1865, .1687, .1687]
).reshape(1, 2, 3, 5)


def test_known_answer_vector():
    loss, dz = O.rnnt_loss_from_logits(KAT_LOGITS, np.array([[1, 2]]), [2], [2], blank=0)
    assert abs(loss[0] - 4.495666) < 1e-5
    assert np.allclose(dz, KAT_GRAD, atol=2e-4)


@pytest.mark.parametrize("T", [1, 2, 3, 4])
@pytest.mark.parametrize("U", [0, 1, 2, 3])
@pytest.mark.parametrize("blank", [0, 4])
def test_brute_force(T, U, blank):
    rng = np.random.default_rng(100 * T + 10 * U + blank)
    V = 5
    z = rng.normal(size=(1, T, U + 1, V))
    labels = [k for k in range(V) if k != blank]
    y = rng.choice(labels, size=(1, max(U, 1)))
    loss, _ = O.rnnt_loss_from_logits(z, y, [T], [U], blank)
    assert abs(loss[0] - O.brute_force_loss(z[0], y[0][:U], blank)) < 1e-9


def test_alpha_beta_consistency_and_coef_identity():
    rng = np.random.default_rng(7)
    T, U, V = 9, 5, 6
    z = rng.normal(size=(T, U + 1, V))
    lp = z - O.logsumexp(z)[..., None]
    y = rng.integers(0, V - 1, size=U)
    lpb = lp[..., V - 1]
    lpl = np.full((T, U + 1), O.NEG)
    for u in range(U):
        lpl[:, u] = lp[:, u, y[u]]
    a, b, lnP = O.lattice_alpha_beta(lpb, lpl, T, U)
    assert abs(lnP - b[0, 0]) < 1e-10
    c1, c2 = O.lattice_coefs(a, b, lpb, lpl, lnP, T, U)
    assert np.allclose(c1 + c2, np.exp(a + b - lnP), atol=1e-12)
    # every anti-diagonal carries total occupancy 1
    for d in range(T + U):
        s = sum(np.exp(a[t, d - t] + b[t, d - t] - lnP) for t in range(T) if 0 <= d - t <= U)
        assert abs(s - 1.0) < 1e-10


def test_finite_differences_joint():
    rng = np.random.default_rng(3)
    B, T, U, V, H = 2, 4, 3, 5, 6
    f = rng.normal(size=(B, T, H)); g = rng.normal(size=(B, U + 1, H))
    W = rng.normal(size=(V, H)) * 0.5; bias = rng.normal(size=V) * 0.1
    y = rng.integers(0, V - 1, size=(B, U)); fl = [4, 3]; yl = [3, 2]
    r = O.rnnt_joint_loss(f, g, W, bias, y, fl, yl, blank=V - 1)

    def total(f_, g_, W_, b_):
        return O.rnnt_joint_loss(f_, g_, W_, b_, y, fl, yl, blank=V - 1)["loss"].sum()

    eps = 1e-6
    for name, arr, grad in (("f", f, r["df"]), ("g", g, r["dg"]), ("W", W, r["dW"]), ("b", bias, r["db"])):
        it = np.nditer(arr, flags=["multi_index"])
        n = 0
        for _ in it:
            idx = it.multi_index
            if n % 3:  # a third of the entries is plenty
                n += 1
                continue
            n += 1
            p = arr.copy(); p[idx] += eps
            m = arr.copy(); m[idx] -= eps
            args = dict(f=f, g=g, W=W, b=bias); args[name] = p
            up = total(args["f"], args["g"], args["W"], args["b"])
            args[name] = m
            dn = total(args["f"], args["g"], args["W"], args["b"])
            assert abs((up - dn) / (2 * eps) - grad[idx]) < 1e-6, (name, idx)
    # padded rows get exactly zero
    assert np.all(r["df"][1, 3] == 0) and np.all(r["dg"][1, 3] == 0)


@pytest.mark.parametrize("blank_last", [True, False])
def test_against_torchaudio(blank_last):
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(11)
    B, T, U, V = 3, 7, 4, 6
    blank = V - 1 if blank_last else 0
    z = rng.normal(size=(B, T, U + 1, V))
    labels = [k for k in range(V) if k != blank]
    y = rng.choice(labels, size=(B, U))
    fl = [7, 5, 6]; yl = [4, 2, 3]
    loss, dz = O.rnnt_loss_from_logits(z, y, fl, yl, blank)
    zt = torch.tensor(z, dtype=torch.float32, requires_grad=True)
    lt = ta.functional.rnnt_loss(
        zt, torch.tensor(y, dtype=torch.int32), torch.tensor(fl, dtype=torch.int32),
        torch.tensor(yl, dtype=torch.int32), blank=blank, reduction="none", fused_log_softmax=True)
    lt.sum().backward()
    assert np.allclose(lt.detach().numpy(), loss, rtol=1e-5)
    assert np.allclose(zt.grad.numpy(), dz, atol=2e-5)


def test_golden_fixture_matches_oracle():
    """tests/golden/rnnt_small.json was produced by tests/golden/make_golden.py
    (torchaudio-checked); the oracle must reproduce it."""
    path = os.path.join(GOLDEN, "rnnt_small.json")
    with open(path) as fh:
        G = json.load(fh)
    for case in G["cases"]:
        a = {k: np.array(v) for k, v in case["inputs"].items()}
        r = O.rnnt_joint_loss(a["f"], a["g"], a["W"], a["bias"], a["y"].astype(int),
                              a["f_lens"].astype(int), a["y_lens"].astype(int), case["blank"])
        assert np.allclose(r["loss"], case["loss"], rtol=1e-9)
        for k in ("df", "dg", "dW", "db"):
            assert np.allclose(r[k], np.array(case[k]), atol=1e-9), k


def test_bf16_round():
    torch = pytest.importorskip("torch")
    x = np.random.default_rng(0).normal(size=1000) * 3
    ref = torch.tensor(x, dtype=torch.float32).bfloat16().double().numpy()
    assert np.array_equal(O.bf16_round(x), ref)


def test_greedy_decode_constructed():
    """Constructed argmax path, as tests/post_process/test_ctc_greedy_decoder.py:16-102 does for CTC."""
    H = V = 6
    blank = V - 1
    W = np.eye(V) * 10.0
    # frame t wants symbol seq[t]; once a symbol is emitted the pred net vetoes everything but blank
    want = [2, blank, 0, 3]
    f = np.zeros((1, len(want), H))
    for t, k in enumerate(want):
        f[0, t, k] = 1.0

    def pred(label, state):
        g = np.zeros(H)
        if label is not None:
            g[:] = -5.0
            g[blank] = 5.0
        return g, None

    hyp, margin = O.greedy_decode(f, [4], W, None, pred, blank, max_symbols_per_step=3)
    # after the first emission the veto stays on (pred ignores history), so only the first symbol appears
    assert hyp == [[2]]
    assert margin > 0


def test_lstm_pred_step_matches_torch_lstm():
    """The oracle's prediction-network step (embedding + LSTM cell + projection) against torch.nn.LSTM -- the module
    the reference's RNN wrapper builds (src/myrtlespeech/model/rnn.py:133-205) -- fed one label at a time."""
    import torch
    from myrtlespeech_b200.model.rnn_t import RNNTPredictionNet

    torch.manual_seed(0)
    V, E, Hp, H = 11, 6, 8, 10
    pred = RNNTPredictionNet(V, E, Hp, 1, H).cpu().double()
    n = lambda t: t.detach().numpy()  # noqa: E731
    step = O.lstm_pred_step(n(pred.embedding.weight), n(pred.rnn.weight_ih_l0), n(pred.rnn.weight_hh_l0),
                            n(pred.rnn.bias_ih_l0), n(pred.rnn.bias_hh_l0), n(pred.proj.weight), n(pred.proj.bias))
    labels = [None, 3, 0, 10, 7, 7]
    state, hx = None, None
    for lab in labels:
        got, state = step(lab, state)
        tl = None if lab is None else torch.tensor([lab])
        want, hx = pred.step(tl, hx, 1, torch.device("cpu"))
        np.testing.assert_allclose(got, want[0].detach().numpy(), rtol=1e-10, atol=1e-12)
    # bf16-faithful variant stays close to the exact one (it only moves rounding points)
    fstep = O.lstm_pred_step(n(pred.embedding.weight), n(pred.rnn.weight_ih_l0), n(pred.rnn.weight_hh_l0),
                             n(pred.rnn.bias_ih_l0), n(pred.rnn.bias_hh_l0), n(pred.proj.weight), n(pred.proj.bias),
                             faithful=True)
    s1 = s2 = None
    for lab in labels:
        a, s1 = step(lab, s1)
        b, s2 = fstep(lab, s2)
        assert np.max(np.abs(a - b)) < 3e-2


def test_lstm_pred_step_two_layers_matches_torch_lstm():
    import torch
    from myrtlespeech_b200.model.rnn_t import RNNTPredictionNet

    torch.manual_seed(1)
    V, E, Hp, H = 9, 5, 8, 6
    pred = RNNTPredictionNet(V, E, Hp, 2, H).cpu().double()
    n = lambda t: t.detach().numpy()  # noqa: E731
    r = pred.rnn
    step = O.lstm_pred_step(n(pred.embedding.weight), [n(r.weight_ih_l0), n(r.weight_ih_l1)],
                            [n(r.weight_hh_l0), n(r.weight_hh_l1)], [n(r.bias_ih_l0), n(r.bias_ih_l1)],
                            [n(r.bias_hh_l0), n(r.bias_hh_l1)], n(pred.proj.weight), n(pred.proj.bias))
    state, hx = None, None
    for lab in [None, 2, 8, 0, 0, 5]:
        got, state = step(lab, state)
        want, hx = pred.step(None if lab is None else torch.tensor([lab]), hx, 1, torch.device("cpu"))
        np.testing.assert_allclose(got, want[0].detach().numpy(), rtol=1e-10, atol=1e-12)


def test_gru_pred_step_matches_torch_gru():
    import torch
    from myrtlespeech_b200.model.rnn_t import RNNTPredictionNet

    torch.manual_seed(2)
    V, E, Hp, H = 9, 5, 8, 6
    pred = RNNTPredictionNet(V, E, Hp, 2, H, rnn_type="gru").cpu().double()
    n = lambda t: t.detach().numpy()  # noqa: E731
    r = pred.rnn
    step = O.gru_pred_step(n(pred.embedding.weight), [n(r.weight_ih_l0), n(r.weight_ih_l1)],
                           [n(r.weight_hh_l0), n(r.weight_hh_l1)], [n(r.bias_ih_l0), n(r.bias_ih_l1)],
                           [n(r.bias_hh_l0), n(r.bias_hh_l1)], n(pred.proj.weight), n(pred.proj.bias))
    state, hx = None, None
    for lab in [None, 2, 8, 0, 0, 5]:
        got, state = step(lab, state)
        want, hx = pred.step(None if lab is None else torch.tensor([lab]), hx, 1, torch.device("cpu"))
        np.testing.assert_allclose(got, want[0].detach().numpy(), rtol=1e-10, atol=1e-12)


def test_verify_greedy_transcript_accepts_the_greedy_path_and_rejects_others():
    """The decision-by-decision checker used for long GPU decodes: the oracle's own transcript is accepted with zero
    regret; a changed, a truncated and an extended transcript are rejected."""
    rng = np.random.default_rng(0)
    V, H, Hp, E, T, S = 12, 16, 8, 4, 25, 3
    step = O.lstm_pred_step(rng.normal(size=(V + 1, E)), [rng.normal(size=(4 * Hp, E))], [rng.normal(size=(4 * Hp, Hp))],
                            [None], [None], rng.normal(size=(H, Hp)), None, faithful=True)
    f = rng.normal(size=(1, T, H)); W = rng.normal(size=(V, H))
    hyps, _ = O.greedy_decode(f, [T], W, None, step, V - 1, S, faithful=True)
    hyp = hyps[0]
    assert len(hyp) > 10
    ok, regret, ties = O.verify_greedy_transcript(f[0], T, W, None, step, V - 1, S, hyp, 1e-9, faithful=True)
    assert ok and regret == 0.0 and ties == 0
    bad = list(hyp); bad[3] = (bad[3] + 1) % (V - 1)
    for other in (bad, hyp[:-1], hyp + [1]):
        ok, regret, _ = O.verify_greedy_transcript(f[0], T, W, None, step, V - 1, S, other, 1e-2, faithful=True)
        assert not ok and regret == float("inf")
    # a huge eps accepts anything that fits the lattice, and reports how far from greedy it was
    ok, regret, _ = O.verify_greedy_transcript(f[0], T, W, None, step, V - 1, S, bad, 1e3, faithful=True)
    assert ok and regret > 1e-2

"""Pins ``tests/torch_reference.py`` (the chunked fp64 checker the full-size GPU parity tests use) to the numpy oracle:
same loss and gradients on small ragged cases, in both the exact and the bf16-faithful modes."""
import numpy as np
import pytest

from oracle import rnnt_oracle as O
from tests import torch_reference as R


def _case(seed, B, T, U, V, H, blank):
    rng = np.random.default_rng(seed)
    f = O.bf16_round(rng.normal(size=(B, T, H)))
    g = O.bf16_round(rng.normal(size=(B, U + 1, H)))
    W = O.bf16_round(rng.uniform(-1, 1, size=(V, H)) / np.sqrt(H))
    bias = rng.uniform(-1, 1, size=V) / np.sqrt(H)
    labels = np.array([k for k in range(V) if k != blank])
    y = rng.choice(labels, size=(B, max(U, 1)))[:, :U].reshape(B, U)
    fl = rng.integers(max(1, T // 2), T + 1, size=B); fl[0] = T
    yl = rng.integers(U // 2, U + 1, size=B); yl[0] = U
    return f, g, W, bias, y, fl, yl


CASES = [(1, 1, 2, 2, 5, 8, 0), (2, 3, 9, 4, 7, 16, 6), (3, 2, 13, 6, 29, 24, 28), (4, 2, 7, 0, 5, 8, 2),
         (5, 3, 21, 9, 40, 32, 17), (6, 1, 1, 3, 6, 8, 5)]


@pytest.mark.parametrize("faithful", [False, True])
@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "B%d_T%d_U%d_V%d_H%d" % c[1:6])
def test_torch_reference_equals_numpy_oracle(cfg, faithful):
    f, g, W, bias, y, fl, yl = _case(*cfg)
    blank = cfg[6]
    gl = np.linspace(0.5, 1.5, cfg[1])
    want = O.rnnt_joint_loss(f, g, W, bias, y, fl, yl, blank, grad_loss=gl, faithful=faithful)
    got = R.rnnt_joint_loss(f, g, W, bias, y, fl, yl, blank, grad_loss=gl, faithful=faithful)
    assert np.allclose(got["loss"], want["loss"], rtol=1e-12, atol=1e-12)
    for k in ("df", "dg", "dW", "db"):
        scale = np.abs(want[k]).max() + 1e-300
        # bf16 rounding of dz can flip on a last-bit difference between numpy's and torch's exp/log; one flip moves an
        # element by 2^-9 relative, far below the 1e-3 the GPU tests assert
        tol = 1e-10 if not faithful else 1e-4
        assert np.abs(got[k] - want[k]).max() / scale < tol, k


def test_known_answer_vector():
    """The public KAT through the torch reference's lattice: cost 4.495666."""
    import torch
    z = np.array([.1, .6, .1, .1, .1, .1, .1, .6, .1, .1, .1, .1, .2, .8, .1,
                  .1, .6, .1, .1, .1, .1, .1, .2, .1, .1, .7, .1, .2, .1, .1]).reshape(2, 3, 5)
    lp = torch.log_softmax(torch.tensor(z), -1)
    lpb = lp[:, :, 0].contiguous()
    lpl = torch.full((2, 3), R.NEG, dtype=torch.float64)
    lpl[:, 0] = lp[:, 0, 1]; lpl[:, 1] = lp[:, 1, 2]
    alpha, beta, lnp = R._lattice(lpb, lpl, 2, 2)
    assert abs(-float(lnp) - 4.495666) < 1e-6
    assert abs(float(beta[0, 0]) - float(lnp)) < 1e-12

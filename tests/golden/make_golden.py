"""Generates tests/golden/rnnt_small.json.

The reference has no RNN-T implementation to import, so the
fixture is produced by the fp64 oracle and cross-checked here against
torchaudio.functional.rnnt_loss + torch autograd through the eager joint before
it is written.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch
import torchaudio

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import rnnt_oracle as O  # noqa: E402


def case(seed, B, T, U, V, H, blank, ragged):
    rng = np.random.default_rng(seed)
    f = rng.normal(size=(B, T, H)).round(4)
    g = rng.normal(size=(B, U + 1, H)).round(4)
    W = (rng.uniform(-1, 1, size=(V, H)) / np.sqrt(H)).round(4)
    bias = (rng.uniform(-1, 1, size=V) / np.sqrt(H)).round(4)
    labels = [k for k in range(V) if k != blank]
    y = rng.choice(labels, size=(B, U))
    fl = np.full(B, T); yl = np.full(B, U)
    if ragged:
        fl = np.sort(rng.integers(max(1, T // 2), T + 1, size=B))[::-1].copy(); fl[0] = T
        yl = rng.integers(U // 2, U + 1, size=B); yl[0] = U
    r = O.rnnt_joint_loss(f, g, W, bias, y, fl, yl, blank)
    # cross-check with torch eager + torchaudio (fp64 joint, fp32 loss)
    tf = torch.tensor(f, requires_grad=True); tg = torch.tensor(g, requires_grad=True)
    tW = torch.tensor(W, requires_grad=True); tb = torch.tensor(bias, requires_grad=True)
    z = torch.tanh(tf[:, :, None] + tg[:, None]) @ tW.T + tb
    l = torchaudio.functional.rnnt_loss(z.float(), torch.tensor(y, dtype=torch.int32),
                                        torch.tensor(fl, dtype=torch.int32), torch.tensor(yl, dtype=torch.int32),
                                        blank=blank, reduction="none")
    l.sum().backward()
    assert np.allclose(l.detach().numpy(), r["loss"], rtol=1e-5), (l, r["loss"])
    for k, t in (("df", tf), ("dg", tg), ("dW", tW), ("db", tb)):
        assert np.allclose(t.grad.numpy(), r[k], atol=5e-5), k
    return dict(
        blank=blank,
        inputs=dict(f=f.tolist(), g=g.tolist(), W=W.tolist(), bias=bias.tolist(), y=y.tolist(),
                    f_lens=fl.tolist(), y_lens=yl.tolist()),
        loss=r["loss"].tolist(), df=r["df"].tolist(), dg=r["dg"].tolist(),
        dW=r["dW"].tolist(), db=r["db"].tolist(),
    )


if __name__ == "__main__":
    cases = [
        case(1, 2, 5, 3, 6, 8, 5, False),
        case(2, 3, 9, 4, 7, 16, 0, True),
        case(3, 2, 12, 6, 29, 24, 28, True),
    ]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rnnt_small.json")
    with open(out, "w") as fh:
        json.dump(dict(generator="tests/golden/make_golden.py", checked_against="torchaudio 2.11 rnnt_loss + torch autograd",
                       cases=cases), fh)
    print("wrote", out, os.path.getsize(out), "bytes")

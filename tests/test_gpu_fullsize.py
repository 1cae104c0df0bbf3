"""Full-size parity: the CUDA path against the fp64 checker at every BASELINE.json training configuration.

configs[1] (B=32 T=500 U=100 V=29 H=512), configs[2] (B=32 T=400 U=150 V=H=1024) and the north-star target shape
(B=32 T=500 U=100 V=H=1024; configs[3] is this shape per GPU) are 1.6-1.9 M lattice rows: too large for the numpy
oracle, so the checker is ``tests/torch_reference.py`` -- the oracle's function evaluated utterance by utterance in
fp64 on the GPU, pinned to the numpy oracle by ``tests/test_torch_reference.py``.  loss, df, dg, dW and db are
compared on BOTH kernel schedules (persistent kernels and per-slab kernels).

Tolerances (``north_star``: "loss and gradients within 1e-3 relative in fp32-accumulate"):

* ``rel``    = ||got - ref||_F / ||ref||_F   < 1e-3 for every output;
* ``relmax`` = max|got - ref| / max|ref|     < the per-output bound in ``RELMAX``.

Measured (round 2, both schedules): loss 1e-7, db 1e-6 .. 1e-5, dW 2e-5 (per-slab) / 3e-5 .. 4e-5 (persistent: the fp32 TMEM
accumulators of a dW block are flushed every 256 pair-tiles), dg 2e-5 / relmax 8e-5, df 1.9e-4 / relmax 3.2e-3 .. 4.0e-3.

Why ``relmax`` of df is allowed above 1e-3 -- the one bound that is not met element-wise.  The kernels evaluate
``h = bf16(tanh.approx.f32(f + g))``; the checker evaluates ``bf16(tanh(f + g))`` with a correctly rounded tanh.
``scripts/micro/tanh_flip.cu`` measures how often the two disagree on this input distribution: 3.4e-5 of all ``h``
(1.2e-6 with an ``ex2`` + ``rcp`` tanh at half the MUFU rate, 7e-8 with libm ``tanhf``) -- always by exactly one bf16
ulp.  At the target shape that is ~5e4 flipped values among 1.6e9.  A flipped ``h`` moves the factor ``1 - h^2`` of its
``dpre = dh (1 - h^2)`` by ``2 |h| 2^-8 <= 7.8e-3``, i.e. by about 1.5 % of a typical term, and ``df[b,t,:]`` is a sum
over only U+1 ~ 100 such terms dominated by the few lattice cells the alignment passes through, so the worst of the
5e4 flips shows up as 3e-3 .. 4e-3 of max|df|; the 99.7 % of df elements without a flipped term agree to 1e-5.  dg
(500 terms per element), dW and db (1.6e6 rows per element) average the flips away.  No implementation of tanh other
than the checker's own removes the flips altogether (any last-bit difference flips some roundings), and against exact
arithmetic the error of either side is dominated by the bf16 rounding of ``h`` itself (2^-9 on EVERY element), so the
bound is a property of comparing two bf16-faithful evaluations, not a loss of accuracy.
"""
import json
import os

import numpy as np
import pytest
import torch

import myrtlespeech_b200 as M
from tests import torch_reference as R

pytestmark = pytest.mark.gpu

TOL = 1e-3
#: element-wise bounds, max|got - ref| / max|ref| (see the module docstring for df / dg)
RELMAX = {"loss": 1e-5, "df": 6e-3, "dg": 5e-4, "dW": 3e-4, "db": 1e-4}

CONFIGS = {
    # name: (B, T, U, V, H)
    "configs1_chars": (32, 500, 100, 29, 512),
    "configs2_subword": (32, 400, 150, 1024, 1024),
    "target": (32, 500, 100, 1024, 1024),
}
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPORT = os.path.join(_ROOT, "gpurun_out", "parity_fullsize.json")


def _inputs(name, ragged):
    """SURVEY.md 8(d) synthetic inputs: f, g ~ N(0,1), W, bias ~ U(-1/sqrt(H), 1/sqrt(H)), bf16-rounded once."""
    B, T, U, V, H = CONFIGS[name]
    gen = torch.Generator().manual_seed(1234)
    f = torch.randn(B, T, H, generator=gen).bfloat16()
    g = torch.randn(B, U + 1, H, generator=gen).bfloat16()
    W = ((torch.rand(V, H, generator=gen) * 2 - 1) / H ** 0.5).bfloat16()
    bias = (torch.rand(V, generator=gen) * 2 - 1) / H ** 0.5
    y = torch.randint(0, V - 1, (B, U), generator=gen, dtype=torch.int32)
    fl = torch.full((B,), T, dtype=torch.int64)
    yl = torch.full((B,), U, dtype=torch.int64)
    if ragged:  # T_b ~ U[T/2, T], U_b ~ U[U/2, U], sorted by T_b descending as the reference collate does
        fl = torch.randint(T // 2, T + 1, (B,), generator=gen).sort(descending=True).values
        yl = torch.randint(U // 2, U + 1, (B,), generator=gen)
        fl[0] = T
        yl[0] = U
    gl = torch.linspace(0.5, 1.5, B)      # a non-trivial upstream gradient per utterance
    return f, g, W, bias, y, fl, yl, gl


_ref_cache = {}


def _reference(name, ragged):
    key = (name, ragged)
    if key not in _ref_cache:
        f, g, W, bias, y, fl, yl, gl = _inputs(name, ragged)
        V = CONFIGS[name][3]
        _ref_cache[key] = R.rnnt_joint_loss(f.float().numpy(), g.float().numpy(), W.float().numpy(), bias.numpy(),
                                            y.numpy(), fl.numpy(), yl.numpy(), V - 1, grad_loss=gl.numpy(),
                                            faithful=True, device="cuda")
        torch.cuda.empty_cache()
    return _ref_cache[key]


def _cuda(name, ragged):
    f, g, W, bias, y, fl, yl, gl = _inputs(name, ragged)
    V = CONFIGS[name][3]
    fd = f.float().cuda().requires_grad_(True); gd = g.float().cuda().requires_grad_(True)   # fp32 leaves: fp32 gradients
    Wd = W.float().cuda().requires_grad_(True); bd = bias.cuda().requires_grad_(True)
    loss = M.rnnt_joint_loss(fd, gd, Wd, bd, y.cuda(), fl, yl, V - 1)
    loss.backward(gl.cuda())
    torch.cuda.synchronize()
    return dict(loss=loss.detach().cpu().numpy(), df=fd.grad.float().cpu().numpy(), dg=gd.grad.float().cpu().numpy(),
                dW=Wd.grad.float().cpu().numpy(), db=bd.grad.cpu().numpy())


def _errors(got, ref):
    out = {}
    for k in ("loss", "df", "dg", "dW", "db"):
        a = np.asarray(got[k], dtype=np.float64); b = np.asarray(ref[k], dtype=np.float64)
        out[k] = dict(rel=float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)),
                      relmax=float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300)))
    return out


def _record(key, errs):
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        rep = {}
        if os.path.exists(_REPORT):
            with open(_REPORT) as fh:
                rep = json.load(fh)
        rep[key] = errs
        with open(_REPORT, "w") as fh:
            json.dump(rep, fh, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.mark.parametrize("schedule", ["persistent", "recompute", "slab"])
@pytest.mark.parametrize("name,ragged", [("configs1_chars", False), ("configs2_subword", False), ("target", False),
                                         ("target", True)])
def test_full_size_matches_fp64_reference(name, ragged, schedule):
    from myrtlespeech_b200 import _lib, functional as F
    lib = _lib.load()
    lib.rnnt_debug_set(b"path", 0 if schedule == "slab" else 1)
    F.set_keep_activations(schedule == "persistent")     # "persistent": activations kept (the default); "recompute": not
    try:
        got = _cuda(name, ragged)
    finally:
        lib.rnnt_debug_set(b"path", 1)
        F.set_keep_activations(True)
    ref = _reference(name, ragged)
    errs = _errors(got, ref)
    _record(f"{name}{'_ragged' if ragged else ''}/{schedule}", errs)
    bad = {k: e for k, e in errs.items() if not (e["rel"] < TOL and e["relmax"] < RELMAX[k])}
    assert not bad, bad
    # padded rows get exactly zero gradient
    _, _, _, _, _, fl, yl, _ = _inputs(name, ragged)
    for b in range(got["df"].shape[0]):
        assert np.all(got["df"][b, int(fl[b]):] == 0) and np.all(got["dg"][b, int(yl[b]) + 1:] == 0)

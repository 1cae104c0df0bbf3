"""CPU oracle for the RNN-T transducer head (joint + RNNTLoss + greedy decode).

TEST INFRASTRUCTURE ONLY.  Nothing under ``myrtlespeech_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker.

PARITY STATUS: **parity unpinned against the reference.**  The reference
snapshot has no RNN-T code at all (``src/myrtlespeech/loss/``
holds only ``ctc_loss.py``; ``protos/speech_to_text.proto:20-34`` lists only CTC
members), so there is no reference output, test or golden vector for this
path.  The oracle is instead pinned by (see ``tests/test_oracle.py``):

* the public RNN-T known-answer vector (B=1,T=2,U=2,V=5 -> cost 4.495666 and
  its gradient),
* brute-force enumeration of every alignment for all T<=4, U<=3,
* ``torchaudio.functional.rnnt_loss`` (torchaudio 2.11, CPU) on ragged batches,
* finite differences in fp64.

What it restates
----------------
* joint: ``z[b,t,u,:] = W . tanh(f[b,t,:] + g[b,u,:]) + bias``  (BASELINE.json
  ``north_star``; composed the way ``model/fully_connected.py:133-166`` composes
  a Linear over ``(x, lens)`` tuples).
* loss: Graves-2012 alpha/beta lattice over ``log_softmax(z)`` (SURVEY.md
  Appendix A).  The loss module owns the log-softmax exactly as the CTC analog
  does (``loss/ctc_loss.py:45,95``).
* reductions follow ``loss/ctc_loss.py:12-23`` naming (none/mean/sum); RNN-T
  ``mean`` is the batch mean (the reference does not define it for RNN-T).
* greedy decode: per frame, emit argmax symbols until blank or
  ``max_symbols_per_step``; output type ``List[List[int]]`` as
  ``post_process/ctc_greedy_decoder.py:17-94``.

Everything is numpy float64 unless ``faithful=True``, which rounds at the same
points as the CUDA path (h to bf16 after an fp32 add, dz to bf16) so that the
1e-3 tolerance measures the kernels and not bf16 itself.
"""
from __future__ import annotations

import itertools
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

NEG = -1.0e30


# --------------------------------------------------------------------------- #
# bf16 helpers
# --------------------------------------------------------------------------- #
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float64."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    bits = a.view(np.uint32).astype(np.uint64)
    rounding = ((bits >> 16) & 1) + 0x7FFF
    bits = ((bits + rounding) >> 16) << 16
    out = bits.astype(np.uint32).view(np.float32)
    out = np.where(np.isfinite(a), out, a)
    return out.astype(np.float64).reshape(np.shape(x))


def logsumexp(z: np.ndarray, axis: int = -1) -> np.ndarray:
    m = np.max(z, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(z - m), axis=axis, keepdims=True))).squeeze(axis)


# --------------------------------------------------------------------------- #
# joint
# --------------------------------------------------------------------------- #
def joint_hidden(f: np.ndarray, g: np.ndarray, faithful: bool = False) -> np.ndarray:
    """h[b,t,u,:] = tanh(f[b,t,:] + g[b,u,:]) -> (B,T,U1,H)."""
    f = np.asarray(f, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    x = f[:, :, None, :] + g[:, None, :, :]
    if faithful:
        x = x.astype(np.float32).astype(np.float64)
    h = np.tanh(x)
    if faithful:
        h = bf16_round(h)
    return h


def joint_logits(f, g, W, bias, faithful: bool = False) -> np.ndarray:
    """(B,T,U1,V) logits of the additive-tanh joint."""
    h = joint_hidden(f, g, faithful)
    W = np.asarray(W, dtype=np.float64)
    z = h @ W.T
    if bias is not None:
        z = z + np.asarray(bias, dtype=np.float64)
    return z


# --------------------------------------------------------------------------- #
# lattice
# --------------------------------------------------------------------------- #
def lattice_alpha_beta(lpb: np.ndarray, lpl: np.ndarray, T: int, U: int):
    """alpha, beta over one utterance.

    lpb[t,u] blank log-prob, lpl[t,u] label log-prob (valid for u<U).
    Lattice is T x (U+1).  Returns (alpha, beta, lnP).
    """
    U1 = U + 1
    alpha = np.full((T, U1), NEG)
    beta = np.full((T, U1), NEG)
    alpha[0, 0] = 0.0
    for t in range(T):
        for u in range(U1):
            if t == 0 and u == 0:
                continue
            a = alpha[t - 1, u] + lpb[t - 1, u] if t > 0 else NEG
            b = alpha[t, u - 1] + lpl[t, u - 1] if u > 0 else NEG
            alpha[t, u] = np.logaddexp(a, b)
    beta[T - 1, U] = lpb[T - 1, U]
    for t in range(T - 1, -1, -1):
        for u in range(U, -1, -1):
            if t == T - 1 and u == U:
                continue
            a = beta[t + 1, u] + lpb[t, u] if t < T - 1 else NEG
            b = beta[t, u + 1] + lpl[t, u] if u < U else NEG
            beta[t, u] = np.logaddexp(a, b)
    lnP = alpha[T - 1, U] + lpb[T - 1, U]
    return alpha, beta, lnP


def lattice_coefs(alpha, beta, lpb, lpl, lnP, T: int, U: int):
    """Occupancy coefficients: c1 (blank arc), c2 (label arc); c0 = c1 + c2.

    dL/dz[t,u,k] = softmax_k * c0 - [k==blank] c1 - [k==y_u] c2.
    """
    U1 = U + 1
    c1 = np.zeros((T, U1))
    c2 = np.zeros((T, U1))
    for t in range(T):
        for u in range(U1):
            if t < T - 1:
                c1[t, u] = np.exp(alpha[t, u] + lpb[t, u] + beta[t + 1, u] - lnP)
            elif u == U:
                c1[t, u] = np.exp(alpha[t, u] + lpb[t, u] - lnP)
            if u < U:
                c2[t, u] = np.exp(alpha[t, u] + lpl[t, u] + beta[t, u + 1] - lnP)
    return c1, c2


def rnnt_loss_from_logits(
    logits: np.ndarray,
    y: np.ndarray,
    f_lens: Sequence[int],
    y_lens: Sequence[int],
    blank: int,
) -> Tuple[np.ndarray, np.ndarray]:
    """Per-utterance loss (B,) and d loss_b / d logits (B,T,U1,V), fp64.

    Rows with t >= f_lens[b] or u > y_lens[b] get exactly zero gradient.
    """
    z = np.asarray(logits, dtype=np.float64)
    B, Tm, U1m, V = z.shape
    loss = np.zeros(B)
    dz = np.zeros_like(z)
    for b in range(B):
        T, U = int(f_lens[b]), int(y_lens[b])
        zb = z[b, :T, : U + 1]
        lse = logsumexp(zb)
        lp = zb - lse[..., None]
        lpb = lp[..., blank]
        lpl = np.full((T, U + 1), NEG)
        for u in range(U):
            lpl[:, u] = lp[:, u, int(y[b, u])]
        alpha, beta, lnP = lattice_alpha_beta(lpb, lpl, T, U)
        loss[b] = -lnP
        c1, c2 = lattice_coefs(alpha, beta, lpb, lpl, lnP, T, U)
        g = np.exp(lp) * (c1 + c2)[..., None]
        g[..., blank] -= c1
        for u in range(U):
            g[:, u, int(y[b, u])] -= c2[:, u]
        dz[b, :T, : U + 1] = g
    return loss, dz


def reduce_loss(loss: np.ndarray, reduction: str) -> np.ndarray:
    if reduction == "none":
        return loss
    if reduction == "sum":
        return loss.sum()
    if reduction == "mean":
        return loss.mean()
    raise ValueError(f"reduction={reduction} not supported")


def rnnt_joint_loss(
    f, g, W, bias, y, f_lens, y_lens, blank: int,
    grad_loss: Optional[np.ndarray] = None,
    faithful: bool = False,
):
    """Joint + loss forward and analytic backward.

    Returns dict(loss (B,), df, dg, dW, db, lse, lp_blank, lp_label).
    ``grad_loss`` is d(total)/d(loss_b), default ones (i.e. reduction="sum").
    """
    f = np.asarray(f, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    B, Tm, H = f.shape
    U1m = g.shape[1]
    V = W.shape[0]
    bias_ = np.zeros(V) if bias is None else np.asarray(bias, dtype=np.float64)
    if grad_loss is None:
        grad_loss = np.ones(B)
    h = joint_hidden(f, g, faithful)
    z = h @ W.T + bias_
    loss, dz = rnnt_loss_from_logits(z, y, f_lens, y_lens, blank)
    dz = dz * np.asarray(grad_loss, dtype=np.float64)[:, None, None, None]
    if faithful:
        dz = bf16_round(dz)
    lse = logsumexp(z)
    lp = z - lse[..., None]
    dh = dz @ W
    dpre = dh * (1.0 - h * h)
    df = dpre.sum(axis=2)
    dg = dpre.sum(axis=1)
    dW = np.einsum("btuv,btuh->vh", dz, h)
    db = dz.sum(axis=(0, 1, 2))
    lp_blank = lp[..., blank]
    lp_label = np.zeros((B, Tm, U1m))
    for b in range(B):
        for u in range(int(y_lens[b])):
            lp_label[b, :, u] = lp[b, :, u, int(y[b, u])]
    return dict(loss=loss, df=df, dg=dg, dW=dW, db=db, lse=lse,
                lp_blank=lp_blank, lp_label=lp_label, dz=dz)


# --------------------------------------------------------------------------- #
# brute force (pins the DP itself)
# --------------------------------------------------------------------------- #
def brute_force_loss(logits: np.ndarray, y: Sequence[int], blank: int) -> float:
    """-ln sum over all alignments, by explicit enumeration.  One utterance.

    logits (T, U+1, V).  An alignment is a lattice path from (0,0) that takes U
    label arcs and T blank arcs, the last arc being the blank out of (T-1,U).
    """
    z = np.asarray(logits, dtype=np.float64)
    T, U1, _ = z.shape
    U = U1 - 1
    lp = z - logsumexp(z)[..., None]
    total = -np.inf
    # choose positions of the U label arcs among the first T-1+U arcs
    for labels_at in itertools.combinations(range(T - 1 + U), U):
        t = u = 0
        s = 0.0
        la = set(labels_at)
        for step in range(T - 1 + U):
            if step in la:
                s += lp[t, u, int(y[u])]
                u += 1
            else:
                s += lp[t, u, blank]
                t += 1
        assert t == T - 1 and u == U
        s += lp[t, u, blank]
        total = np.logaddexp(total, s)
    return -float(total)


# --------------------------------------------------------------------------- #
# greedy decode
# --------------------------------------------------------------------------- #
def greedy_decode(
    f: np.ndarray,
    f_lens: Sequence[int],
    W: np.ndarray,
    bias: Optional[np.ndarray],
    pred_step: Callable[[Optional[int], object], Tuple[np.ndarray, object]],
    blank: int,
    max_symbols_per_step: int,
    faithful: bool = False,
    per_utterance_margin: bool = False,
    tie_margin: Optional[float] = None,
) -> Tuple[List[List[int]], float]:
    """RNN-T greedy search for each utterance.

    ``pred_step(label_or_None, state) -> (g_vec (H,), new_state)`` is the
    prediction network; ``None`` is the start-of-sequence input.  Returns the
    transcripts and the smallest top-2 logit margin seen (so tests can tell a
    genuine mismatch from an argmax tie within accumulation noise); with
    ``per_utterance_margin`` the second value is the list of per-utterance
    minima instead (utterances are independent, so a near-tie only excuses the
    utterance it occurs in).

    With ``tie_margin`` a third value is returned: per utterance, the number of symbols emitted before the first
    decode step whose top-2 margin is <= ``tie_margin`` (``len(hyp)`` if there is none).  A checker that runs in
    different arithmetic can only be held to the transcript up to that point: every decision before it is clear.

    Ties resolve to the lowest index (numpy/torch argmax semantics).
    """
    f = np.asarray(f, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    V = W.shape[0]
    b_ = np.zeros(V) if bias is None else np.asarray(bias, dtype=np.float64)
    out: List[List[int]] = []
    min_margin = np.inf
    margins: List[float] = []
    clear_prefix: List[int] = []
    for b in range(f.shape[0]):
        hyp: List[int] = []
        first_tie = -1
        min_margin = np.inf if per_utterance_margin else min_margin
        g, state = pred_step(None, None)
        for t in range(int(f_lens[b])):
            for _ in range(max_symbols_per_step):
                x = f[b, t] + np.asarray(g, dtype=np.float64)
                if faithful:
                    x = x.astype(np.float32).astype(np.float64)
                h = np.tanh(x)
                if faithful:
                    h = bf16_round(h)
                z = W @ h + b_
                k = int(np.argmax(z))
                top2 = np.partition(z, -2)[-2:]
                min_margin = min(min_margin, float(top2[1] - top2[0]))
                if tie_margin is not None and first_tie < 0 and float(top2[1] - top2[0]) <= tie_margin:
                    first_tie = len(hyp)
                if k == blank:
                    break
                hyp.append(k)
                g, state = pred_step(k, state)
        out.append(hyp)
        margins.append(float(min_margin))
        clear_prefix.append(len(hyp) if first_tie < 0 else first_tie)
    if tie_margin is not None:
        return out, (margins if per_utterance_margin else float(min_margin)), clear_prefix
    if per_utterance_margin:
        return out, margins
    return out, float(min_margin)


def verify_greedy_transcript(
    f_b: np.ndarray,
    n_frames: int,
    W: np.ndarray,
    bias: Optional[np.ndarray],
    pred_step: Callable[[Optional[int], object], Tuple[np.ndarray, object]],
    blank: int,
    max_symbols_per_step: int,
    hyp: Sequence[int],
    eps: float,
    faithful: bool = False,
) -> Tuple[bool, float, int]:
    """Checks a transcript produced by ANOTHER implementation of the greedy search, decision by decision.

    The prediction-network state depends only on the symbols emitted so far, so the logits of every decision are a
    function of (frame t, symbols emitted i): ``z(t, i) = W . tanh(f[t] + g_i) + b`` with ``g_i`` the prediction output
    after ``hyp[:i]``.  The decoder's run is a path through that (t, i) lattice: "blank" moves to ``(t + 1, i)``,
    "emit hyp[i]" to ``(t, i + 1)`` (at most ``max_symbols_per_step`` emissions per frame, then the frame advances).
    The transcript is accepted if such a path from ``(0, 0)`` to ``(n_frames, len(hyp))`` exists on which every action
    taken is an eps-argmax of its ``z`` (within ``eps`` of the best logit) -- i.e. every decision of the other
    implementation is one this oracle would make too, up to a logit difference of ``eps``.  Unlike comparing whole
    transcripts this verifies EVERY symbol, also after a near-tie went the other way.

    Returns ``(accepted, regret, n_ties)``: ``regret`` is the smallest eps that would have been needed (the minimum
    over accepting paths of the largest ``max(z) - z[action]`` along the path; ``inf`` if none exists) and ``n_ties``
    the number of lattice cells visited in which both actions were within ``eps``.
    """
    f_b = np.asarray(f_b, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    b_ = np.zeros(W.shape[0]) if bias is None else np.asarray(bias, dtype=np.float64)
    hyp = [int(k) for k in hyp]
    S = int(max_symbols_per_step)
    # prediction outputs after every prefix of the transcript
    g_list = []
    g, state = pred_step(None, None)
    g_list.append(np.asarray(g, dtype=np.float64))
    for k in hyp:
        g, state = pred_step(k, state)
        g_list.append(np.asarray(g, dtype=np.float64))
    cache = {}

    def gaps(t, i):
        """(max(z) - z[blank], max(z) - z[hyp[i]] or inf)"""
        if (t, i) not in cache:
            x = f_b[t] + g_list[i]
            if faithful:
                x = x.astype(np.float32).astype(np.float64)
            h = np.tanh(x)
            if faithful:
                h = bf16_round(h)
            z = W @ h + b_
            m = float(z.max())
            cache[(t, i)] = (m - float(z[blank]), m - float(z[hyp[i]]) if i < len(hyp) else np.inf)
        return cache[(t, i)]

    # best[(t, i, n)] = smallest achievable "largest regret so far" on a path reaching that state (Dijkstra-like sweep in
    # topological order: t + i is monotone along every transition)
    import heapq
    start = (0, 0, 0)
    best = {start: 0.0}
    heap = [(0, 0.0, start)]          # ordered by t + i, then regret
    n_ties = 0
    seen = set()
    result = np.inf
    while heap:
        _, r, st = heapq.heappop(heap)
        if st in seen or r > best.get(st, np.inf):
            continue
        seen.add(st)
        t, i, n = st
        if t == n_frames:
            if i == len(hyp):
                result = min(result, r)
            continue
        gb, ge = gaps(t, i)
        if gb <= eps and ge <= eps:
            n_ties += 1
        moves = []
        if gb <= eps:
            moves.append(((t + 1, i, 0), max(r, gb)))
        if ge <= eps and i < len(hyp) and hyp[i] != blank:
            nxt = (t + 1, i + 1, 0) if n + 1 == S else (t, i + 1, n + 1)
            moves.append((nxt, max(r, ge)))
        for nxt, rr in moves:
            if rr < best.get(nxt, np.inf):
                best[nxt] = rr
                heapq.heappush(heap, (nxt[0] + nxt[1], rr, nxt))
    return bool(np.isfinite(result)), float(result), n_ties


# --------------------------------------------------------------------------- #
# prediction network step (embedding + single-layer LSTM + projection)
# --------------------------------------------------------------------------- #
def lstm_pred_step(
    emb: np.ndarray,
    w_ih: np.ndarray,
    w_hh: np.ndarray,
    b_ih: Optional[np.ndarray],
    b_hh: Optional[np.ndarray],
    w_proj: np.ndarray,
    b_proj: Optional[np.ndarray],
    faithful: bool = False,
) -> Callable[[Optional[int], object], Tuple[np.ndarray, object]]:
    """Returns the ``pred_step`` callback of :func:`greedy_decode` for an embedding (V+1, E; row V = start of
    sequence) + LSTM cell (torch gate order i, f, g, o; follows the cell of ``torch.nn.LSTM`` that
    ``src/myrtlespeech/model/rnn.py:133-205`` wraps) + linear projection to the joint width.

    ``faithful=True`` rounds where the one-launch CUDA decode rounds: the input half ``W_ih . emb[v] + b`` is an fp32
    table, ``h`` and ``W_hh`` / ``W_proj`` are bf16 operands of an fp32-accumulating product, ``c`` stays fp32 and the
    projected ``g`` is rounded to bf16.
    """
    # a stack of layers is given as lists (one entry per layer); a single layer may be given bare
    if not isinstance(w_ih, (list, tuple)):
        w_ih, w_hh, b_ih, b_hh = [w_ih], [w_hh], [b_ih], [b_hh]
    n_layers = len(w_ih)
    emb = np.asarray(emb, dtype=np.float64)
    w_ih = [np.asarray(w, dtype=np.float64) for w in w_ih]
    w_hh = [np.asarray(w, dtype=np.float64) for w in w_hh]
    w_proj = np.asarray(w_proj, dtype=np.float64)
    hp = w_hh[0].shape[1]
    biases = []
    for l in range(n_layers):
        bias = np.zeros(4 * hp)
        if b_ih[l] is not None:
            bias = bias + np.asarray(b_ih[l], dtype=np.float64)
        if b_hh[l] is not None:
            bias = bias + np.asarray(b_hh[l], dtype=np.float64)
        biases.append(bias)
    table = emb @ w_ih[0].T + biases[0]
    if faithful:
        table = table.astype(np.float32).astype(np.float64)
        biases = [b.astype(np.float32).astype(np.float64) for b in biases]
        w_hh = [bf16_round(w) for w in w_hh]
        w_ih = [w_ih[0]] + [bf16_round(w) for w in w_ih[1:]]   # upper layers consume bf16 h of the layer below
        w_proj = bf16_round(w_proj)
    bp = np.zeros(w_proj.shape[0]) if b_proj is None else np.asarray(b_proj, dtype=np.float64)
    sos = emb.shape[0] - 1

    def sigmoid(x):
        return 1.0 / (1.0 + np.exp(-x))

    def step(label: Optional[int], state):
        if state is None:
            state = [(np.zeros(hp), np.zeros(hp)) for _ in range(n_layers)]
        new_state = []
        x = None
        for l in range(n_layers):
            h, c = state[l]
            if l == 0:
                a = table[sos if label is None else int(label)] + w_hh[0] @ h
            else:
                a = biases[l] + w_ih[l] @ x + w_hh[l] @ h
            i, f, g, o = a[:hp], a[hp:2 * hp], a[2 * hp:3 * hp], a[3 * hp:]
            c = sigmoid(f) * c + sigmoid(i) * np.tanh(g)
            h = sigmoid(o) * np.tanh(c)
            if faithful:
                c = c.astype(np.float32).astype(np.float64)
                h = bf16_round(h)
            new_state.append((h, c))
            x = h
        out = w_proj @ x + bp
        if faithful:
            out = bf16_round(out)
        return out, new_state

    return step


def gru_pred_step(
    emb: np.ndarray,
    w_ih: Sequence[np.ndarray],
    w_hh: Sequence[np.ndarray],
    b_ih: Sequence[Optional[np.ndarray]],
    b_hh: Sequence[Optional[np.ndarray]],
    w_proj: np.ndarray,
    b_proj: Optional[np.ndarray],
    faithful: bool = False,
) -> Callable[[Optional[int], object], Tuple[np.ndarray, object]]:
    """As :func:`lstm_pred_step` for a stack of GRU layers (``torch.nn.GRU`` gate order r, z, n:
    ``n = tanh(W_in x + b_in + r * (W_hn h + b_hn))``, ``h' = (1 - z) n + z h``).  ``faithful=True`` rounds where the
    one-launch CUDA decode rounds: layer 0's input half is an fp32 table, every product sees bf16 ``h`` and bf16
    weights with fp32 accumulation, the blend uses the fp32 ``h``, and the projected ``g`` is rounded to bf16."""
    n_layers = len(w_ih)
    emb = np.asarray(emb, dtype=np.float64)
    w_ih = [np.asarray(w, dtype=np.float64) for w in w_ih]
    w_hh = [np.asarray(w, dtype=np.float64) for w in w_hh]
    w_proj = np.asarray(w_proj, dtype=np.float64)
    hp = w_hh[0].shape[1]
    z3 = np.zeros(3 * hp)
    b_ih = [z3 if b is None else np.asarray(b, dtype=np.float64) for b in b_ih]
    b_hh = [z3 if b is None else np.asarray(b, dtype=np.float64) for b in b_hh]
    table = emb @ w_ih[0].T + b_ih[0]          # input half of layer 0 per label (b_hr, b_hz are added below)
    if faithful:
        r32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
        table[:, :2 * hp] = r32(table[:, :2 * hp] + b_hh[0][:2 * hp])
        table[:, 2 * hp:] = r32(table[:, 2 * hp:])
        b_hh = [np.concatenate([np.zeros(2 * hp), b[2 * hp:]]) if l == 0 else b for l, b in enumerate(b_hh)]
        w_hh = [bf16_round(w) for w in w_hh]
        w_ih = [w_ih[0]] + [bf16_round(w) for w in w_ih[1:]]
        w_proj = bf16_round(w_proj)
    bp = np.zeros(w_proj.shape[0]) if b_proj is None else np.asarray(b_proj, dtype=np.float64)
    sos = emb.shape[0] - 1

    def sigmoid(x):
        return 1.0 / (1.0 + np.exp(-x))

    def step(label: Optional[int], state):
        if state is None:
            state = [np.zeros(hp) for _ in range(n_layers)]
        new_state = []
        x = None
        for l in range(n_layers):
            h = state[l]
            h_op = bf16_round(h) if faithful else h
            gi = table[sos if label is None else int(label)] if l == 0 else w_ih[l] @ x + b_ih[l]
            gh = w_hh[l] @ h_op + b_hh[l]
            r = sigmoid(gi[:hp] + gh[:hp])
            z = sigmoid(gi[hp:2 * hp] + gh[hp:2 * hp])
            n = np.tanh(gi[2 * hp:] + r * gh[2 * hp:])
            h = n + z * (h - n)
            if faithful:
                h = h.astype(np.float32).astype(np.float64)
            new_state.append(h)
            x = bf16_round(h) if faithful else h
        out = w_proj @ x + bp
        if faithful:
            out = bf16_round(out)
        return out, new_state

    return step

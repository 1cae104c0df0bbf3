"""RNN-T greedy decoder with the interface of ``post_process/ctc_greedy_decoder.py:17-97``."""
from typing import List

import torch

from ..functional import greedy_decode_lstm, greedy_decode_lstm_supported, greedy_step
from ..model.rnn_t import JointHandle


class RNNTGreedyDecoder(torch.nn.Module):
    """Decodes RNN-T output using a greedy strategy.

    For every frame, symbols are emitted while the joint's argmax is not blank, up to
    ``max_symbols_per_step`` per frame.  The joint step (tanh + projection + argmax) runs in the
    CUDA library for the whole batch at once; only the emitted ids (``batch`` int32) come back to the
    host per step, not logits.

    Args:
        blank_index: Index of the "blank" symbol.
        model: An :py:class:`myrtlespeech_b200.model.RNNT`.
        max_symbols_per_step: Maximum number of non-blank symbols per encoder frame.
    """

    def __init__(self, blank_index: int, model: torch.nn.Module, max_symbols_per_step: int = 4):
        super().__init__()
        if max_symbols_per_step < 1:
            raise ValueError(f"max_symbols_per_step={max_symbols_per_step} must be >= 1")
        self.blank_index = blank_index
        self.max_symbols_per_step = max_symbols_per_step
        # not registered as a submodule: the decoder does not own the model's parameters
        object.__setattr__(self, "_model", model)

    @property
    def model(self):
        return self._model

    @torch.no_grad()
    def forward(self, x, lengths: torch.Tensor) -> List[List[int]]:
        r"""Decodes using a greedy strategy.

        Args:
            x: what the model produced or consumes, told apart by type and rank (never by comparing sizes):

                * the :py:class:`JointHandle` a training / evaluation forward pass returned -- the reference's
                  report callback calls ``decoder(*last_output)`` (``run/run.py:94``), and the handle carries the
                  encoder output ``f``;
                * a 3-D tensor ``(batch, seq_len, joint_hidden_size)``: encoder output ``f``, already projected to
                  the joint width (see :py:meth:`decode_encoded`);
                * a 4-D tensor ``(batch, channels, features, seq_len)``: model input in the reference's layout
                  (``data/batch.py:45-107``); ``model.encode`` is applied first.

            lengths: 1D integer tensor of valid frames per sequence (of the model input for 4-D ``x``).

        Returns:
            ``List[List[int]]`` of emitted symbol ids per sequence.

        Raises:
            :py:class:`ValueError`: if ``lengths.dtype`` is not an integer type, if the batch sizes
                differ, or if any length exceeds ``seq_len`` (same checks and messages as
                ``post_process/ctc_greedy_decoder.py:49-72``).
        """
        supported_dtypes = [torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64]
        if lengths.dtype not in supported_dtypes:
            raise ValueError(f"lengths.dtype={lengths.dtype} must be in {supported_dtypes}")
        if isinstance(x, JointHandle):
            x = x.f
        if x.dim() not in (3, 4):
            raise ValueError(f"x must be (batch, seq_len, hidden) or (batch, channels, features, seq_len), got {tuple(x.shape)}")
        x_batch = x.size(0)
        seq_len = x.size(1) if x.dim() == 3 else x.size(3)
        l_batch = len(lengths)
        if x_batch != l_batch:
            raise ValueError(f"batch size of x ({x_batch}) and lengths {l_batch} must be equal")
        if not (lengths <= seq_len).all():
            raise ValueError("length values must be less than or equal to x seq_len")

        model = self._model
        was_training = model.training
        model.eval()
        try:
            if x.dim() == 4:
                x, lengths = model.encode(x, lengths)
            return self.decode_encoded(x, lengths)
        finally:
            model.train(was_training)

    @torch.no_grad()
    def decode_encoded(self, f: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
        """Decodes encoder output ``f`` of size ``(batch, seq_len, joint_hidden_size)`` directly."""
        H = self._model.joint.hidden_size
        if f.dim() != 3 or f.size(2) != H:
            raise ValueError(f"encoder output must have size (batch, seq_len, {H}), got {tuple(f.shape)}")
        return self._decode(f, lengths)

    #: run the whole loop (LSTM cell + projection + joint argmax + bookkeeping) as ONE launch when the prediction network
    #: is an LSTM or GRU of up to three layers (``rnnt_greedy_decode_lstm_stack`` / ``_gru_stack``); otherwise one CUDA
    #: graph per step
    USE_FUSED_LOOP = True
    #: decode steps enqueued between two host checks of "is any utterance still active"
    SYNC_EVERY = 32
    #: replay one captured CUDA graph per decode step instead of ~20 eager launches (falls back to eager launches
    #: if the prediction network cannot be captured)
    USE_CUDA_GRAPH = True

    def _decode(self, f: torch.Tensor, lengths: torch.Tensor) -> List[List[int]]:
        """All bookkeeping (current frame, symbols emitted at it, activity mask, emitted ids) lives on the
        device and one decode step is a fixed sequence of in-place updates, so it is captured once as a CUDA graph
        and replayed; the host only polls the activity mask every ``SYNC_EVERY`` steps.  The reference's greedy CTC
        decoder pays one ``.item()`` sync per frame (``post_process/ctc_greedy_decoder.py:77-91``)."""
        model = self._model
        dev = model.joint.fc.weight.device
        B, T, H = f.shape
        S = self.max_symbols_per_step
        blank = self.blank_index
        fb = f.to(dev, torch.bfloat16).contiguous()
        Wb = model.joint.fc.weight.detach().to(torch.bfloat16).contiguous()
        bias = model.joint.fc.bias
        bias = None if bias is None else bias.detach().float().contiguous()
        lens = lengths.to(dev, torch.int32)
        pred = model.prediction
        if self.USE_FUSED_LOOP and T > 0:
            packed = self._packed_prediction(pred, B, Wb.size(0), H)
            if packed is not None:
                table, whh, wproj, bproj, wup, bup, cell = packed
                sym, n_sym = greedy_decode_lstm(fb, lens.contiguous(), Wb, bias, table, whh, wproj, bproj, blank, S,
                                                W_upper=wup, bias_upper=bup, cell=cell)
                sym_h, n_h = sym.cpu(), n_sym.cpu().tolist()
                return [sym_h[b, : n_h[b]].tolist() for b in range(B)]

        g0, hid0 = pred.step(None, None, B, dev)
        g = g0.float().contiguous().clone()
        hid = _clone_hidden(hid0)
        cap = max(1, T * S)                                   # at most S symbols per frame
        i32 = dict(dtype=torch.int32, device=dev)
        sym = torch.zeros(B, cap, **i32)
        n_sym = torch.zeros(B, **i32)
        t_cur = torch.zeros(B, **i32)        # current frame of each utterance
        emitted = torch.zeros(B, **i32)      # symbols emitted at the current frame
        is_sym = torch.zeros(B, **i32)
        label = torch.zeros(B, **i32)
        active = (t_cur < lens).to(torch.int32)

        def step():
            # joint argmax + emit / count / advance for the whole batch in one kernel (state updated in place)
            greedy_step(fb, g, Wb, bias, lens, t_cur, emitted, n_sym, sym, is_sym, label, active, blank, S)
            # advance the prediction network only where a symbol was emitted (computed for all, selected by mask)
            m = is_sym.bool()
            g_new, hid_new = pred.step(label.long(), hid, B, dev)
            g.copy_(torch.where(m[:, None], g_new.float(), g))
            _update_hidden(m, hid_new, hid)

        graph = None
        if self.USE_CUDA_GRAPH:
            graph = _try_capture(step, (g, hid, sym, n_sym, t_cur, emitted, active, is_sym, label))
        n = 0
        while True:
            if n % self.SYNC_EVERY == 0 and not bool(active.any()):
                break
            n += 1
            if graph is not None:
                graph.replay()
            else:
                step()
        sym_h, n_h = sym.cpu(), n_sym.cpu().tolist()
        return [sym_h[b, : n_h[b]].tolist() for b in range(B)]

    def _packed_prediction(self, pred, B: int, V: int, H: int):
        """``_pack_lstm_prediction`` memoised over repeated decodes with unchanged weights (evaluation, serving): they
        skip the table product, the bf16 conversions and their launches.

        The cache is keyed by the identity, storage address and in-place version of every prediction-network
        parameter, and validated by content: an fp64 (sum, sum of magnitudes) per parameter, computed on the device
        and compared on every call (about twenty small reductions and one comparison against a 23 ms decode).  Updates
        that bypass the version counter -- ``p.data.copy_``, ``p.data = ...``, an optimiser or EMA swap writing
        ``.data`` -- therefore repack as well.  :py:meth:`invalidate_cache` drops the cache explicitly."""
        params = list(pred.parameters()) if isinstance(pred, torch.nn.Module) else []
        if not params:
            return _pack_lstm_prediction(pred, B, V, H)
        key = (B, V, H, tuple((id(q), q.data_ptr(), q._version, q.device, tuple(q.shape)) for q in params))
        sums = [q.detach().double().sum() for q in params] + [q.detach().double().abs().sum() for q in params]
        check = torch.stack([t.to(params[0].device) for t in sums])
        cache = self.__dict__.get("_pack_cache")
        if cache is None or cache[0] != key or not torch.equal(cache[1], check):
            cache = (key, check, _pack_lstm_prediction(pred, B, V, H))
            self.__dict__["_pack_cache"] = cache
        return cache[2]

    def invalidate_cache(self) -> None:
        """Forgets the packed prediction-network weights; the next decode repacks them."""
        self.__dict__.pop("_pack_cache", None)

    def extra_repr(self) -> str:
        return f"blank_index={self.blank_index}, max_symbols_per_step={self.max_symbols_per_step}"


def _pack_lstm_prediction(pred, B: int, V: int, H: int):
    """``(gate_table, W_hh, W_proj, b_proj, W_upper, bias_upper, cell)`` for the one-launch decode, or None if ``pred`` is
    not an embedding + unidirectional LSTM / GRU of 1..3 layers + projection (``RNNTPredictionNet`` layout) the fused kernel
    covers.

    LSTM: ``gate_table[v] = W_ih . emb[v] + b_ih + b_hh`` (fp32) folds the embedding lookup and the input half of the cell
    into one row gather per emitted label; row ``vocab_size`` is the start-of-sequence input.  Upper layers are passed as
    ``[W_ih_l | W_hh_l]`` with summed biases.  GRU: four rows per hidden unit -- r, z, the hidden half of n and the input
    half of n -- so that the kernel moves the same data as for an LSTM (``rnnt_greedy_decode_gru_stack``)."""
    emb, rnn, proj = getattr(pred, "embedding", None), getattr(pred, "rnn", None), getattr(pred, "proj", None)
    if not (isinstance(emb, torch.nn.Embedding) and isinstance(rnn, (torch.nn.LSTM, torch.nn.GRU))
            and isinstance(proj, torch.nn.Linear)):
        return None
    if rnn.num_layers > 3 or rnn.bidirectional or getattr(rnn, "proj_size", 0) != 0:
        return None
    if emb.num_embeddings != V + 1 or proj.out_features != H or not emb.weight.is_cuda:
        return None
    Hp, L = rnn.hidden_size, rnn.num_layers
    if not greedy_decode_lstm_supported(B, V, H, Hp, L):
        return None
    dev = emb.weight.device

    def prm(name, l):
        return getattr(rnn, f"{name}_l{l}").detach().float()

    def bias(name, l):
        return prm(name, l) if rnn.bias else torch.zeros(prm("weight_ih", l).size(0), device=dev)

    wproj = proj.weight.detach().to(torch.bfloat16).contiguous()
    bproj = None if proj.bias is None else proj.bias.detach().float().contiguous()
    e = emb.weight.detach().float()
    if isinstance(rnn, torch.nn.LSTM):
        table = e @ prm("weight_ih", 0).t() + (bias("bias_ih", 0) + bias("bias_hh", 0))
        whh = prm("weight_hh", 0)
        wup = [torch.cat([prm("weight_ih", l), prm("weight_hh", l)], 1) for l in range(1, L)]
        bup = [bias("bias_ih", l) + bias("bias_hh", l) for l in range(1, L)]
        cell = "lstm"
    else:
        def split(t):
            return t[:Hp], t[Hp:2 * Hp], t[2 * Hp:]

        w_ir, w_iz, w_in = split(prm("weight_ih", 0)); w_hr, w_hz, w_hn = split(prm("weight_hh", 0))
        b_ir, b_iz, b_in = split(bias("bias_ih", 0)); b_hr, b_hz, b_hn = split(bias("bias_hh", 0))
        table = torch.cat([e @ w_ir.t() + b_ir + b_hr, e @ w_iz.t() + b_iz + b_hz, b_hn.expand(e.size(0), Hp),
                           e @ w_in.t() + b_in], 1)
        whh = torch.cat([w_hr, w_hz, w_hn, torch.zeros_like(w_hn)], 0)
        wup, bup = [], []
        for l in range(1, L):
            w_ir, w_iz, w_in = split(prm("weight_ih", l)); w_hr, w_hz, w_hn = split(prm("weight_hh", l))
            b_ir, b_iz, b_in = split(bias("bias_ih", l)); b_hr, b_hz, b_hn = split(bias("bias_hh", l))
            z = torch.zeros_like(w_hn)
            wup.append(torch.cat([torch.cat([w_ir, w_hr], 1), torch.cat([w_iz, w_hz], 1), torch.cat([z, w_hn], 1),
                                  torch.cat([w_in, z], 1)], 0))
            bup.append(torch.cat([b_ir + b_hr, b_iz + b_hz, b_hn, b_in]))
        cell = "gru"
    wup_t = torch.stack(wup).to(torch.bfloat16).contiguous() if wup else None
    bup_t = torch.stack(bup).contiguous() if bup else None
    return table.contiguous(), whh.to(torch.bfloat16).contiguous(), wproj, bproj, wup_t, bup_t, cell


def _clone_hidden(h):
    if h is None:
        return None
    if isinstance(h, tuple):
        return tuple(_clone_hidden(x) for x in h)
    return h.clone()


def _update_hidden(mask: torch.Tensor, new, old) -> None:
    """In place: old <- where(mask, new, old) for every state tensor (stateless networks carry None)."""
    if new is None or old is None:
        return
    if isinstance(new, tuple):
        for n, o in zip(new, old):
            _update_hidden(mask, n, o)
        return
    old.copy_(torch.where(mask[None, :, None], new, old))


def _try_capture(step, state):
    """Captures ``step`` as a CUDA graph.  The warm-up steps it needs run for real, so the state is snapshotted
    before and restored afterwards.  Returns None if capture is not possible."""
    def flat(x, out):
        if x is None:
            return out
        if isinstance(x, tuple):
            for y in x:
                flat(y, out)
            return out
        out.append(x)
        return out

    tensors = flat(tuple(state), [])
    saved = [t.clone() for t in tensors]
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
    except Exception:
        graph = None
        torch.cuda.synchronize()
    for t, v in zip(tensors, saved):
        t.copy_(v)
    return graph

"""Utterance sharding across the GPUs of one box and the single per-step exchange.

The path shards by utterance: joint, lattice and decode are independent per utterance, so a global
batch is split between ranks with no data-path collective.  What couples the ranks is every *parameter* gradient
-- the joint's ``dW`` / ``db`` directly, and the encoder's and prediction network's through ``df`` / ``dg``, which
back-propagate into parameters that are replicated on every rank -- plus the scalar loss.  One all-reduce (sum)
per step of one flat fp32 buffer ``[grad of every trainable parameter | loss_sum | n_utterances]`` is the only
collective; ``df`` and ``dg`` themselves stay on the GPU that owns the utterance.

:py:class:`GradientReducer` owns that buffer.  Each parameter's ``.grad`` is a *view* into it, so backward passes
accumulate straight into the reduction buffer (no packing copies), and the all-reduce runs on a side stream so that
it can overlap whatever the caller enqueues next.  The reference has no distributed code at all
(SURVEY.md §2.1), so there is no reference interface to mirror here.

The joint-only helpers (``flat_size`` / ``unpack_step`` / ``allreduce_step``) describe the layout ``bench.py`` uses
when it times the hot path alone -- there the joint's ``dW`` / ``db`` are the only parameter gradients that exist.
"""
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_utterances(f_lens: Sequence[int], y_lens: Sequence[int], world_size: int) -> List[List[int]]:
    """Assigns utterances to ranks balancing the lattice size sum_b T_b * (U_b + 1) (longest first, greedy).

    Returns one sorted index list per rank; every utterance appears exactly once.
    """
    if world_size < 1:
        raise ValueError(f"world_size={world_size} must be >= 1")
    cost = [int(t) * (int(u) + 1) for t, u in zip(f_lens, y_lens)]
    order = sorted(range(len(cost)), key=lambda i: (-cost[i], i))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    load = [0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], len(shards[k]), k))
        shards[r].append(i)
        load[r] += cost[i]
    return [sorted(s) for s in shards]


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class GradientReducer:
    """Flat gradient buffer of a replicated model and its once-per-step all-reduce.

    Args:
        params: the trainable parameters of the *whole* model (encoder, prediction network and joint) -- e.g.
            ``seq_to_seq.parameters()``.  All must be fp32 and live on one device.
        group: process group (default: the world).

    Usage per step::

        reducer.zero()                       # instead of optim.zero_grad(): grads are views, keep them
        loss = stt.loss(stt.model(x)[0], y)  # reduction="sum" over this rank's utterances
        loss.backward()                      # accumulates into the flat buffer through the .grad views
        reducer.set_loss(loss, n_utterances)
        reducer.all_reduce()                 # side stream; returns immediately
        reducer.wait()                       # before optim.step() / reading .grad
        loss_mean = reducer.loss_sum / reducer.n_utterances
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("GradientReducer needs fp32 parameters on one device")
        self.group = group
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + 2, dtype=torch.float32, device=dev)   # [grads | loss_sum | n_utterances]
        o = 0
        for p in self.params:
            p.grad = self.flat[o: o + p.numel()].view_as(p)
            o += p.numel()
        self._n_grad = n
        self._stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._done: Optional[torch.cuda.Event] = None

    def zero(self) -> None:
        self.flat.zero_()

    def set_loss(self, loss_sum: torch.Tensor, n_utterances: int) -> None:
        self.flat[self._n_grad] = loss_sum.detach().float()
        self.flat[self._n_grad + 1] = float(n_utterances)

    def all_reduce(self) -> None:
        """Sums the flat buffer over the ranks (NCCL over NVLink on GPUs, gloo in the CPU tests); no-op for one rank.
        On CUDA the collective is issued on a side stream ordered after everything enqueued so far."""
        if _world(self.group) == 1:
            return
        if self._stream is None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self._done = torch.cuda.Event()
            self._done.record(self._stream)
        self.flat.record_stream(self._stream)

    def wait(self) -> None:
        """Orders the current stream after the all-reduce (no host synchronisation)."""
        if self._done is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(self._done)
            self._done = None

    @property
    def loss_sum(self) -> torch.Tensor:
        return self.flat[self._n_grad]

    @property
    def n_utterances(self) -> torch.Tensor:
        return self.flat[self._n_grad + 1]


# ---- joint-only layout (the hot path timed alone: bench.py) ---------------------------------------------------------
def flat_size(V: int, H: int) -> int:
    return V * H + V + 2


def unpack_step(flat: torch.Tensor, V: int, H: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Views of (dW, db, loss_sum, n) inside ``flat``: make these the ``.grad`` of the joint's weight and bias (as
    :py:class:`GradientReducer` does for a whole model) and the backward pass accumulates into the reduction buffer."""
    return flat[: V * H].view(V, H), flat[V * H: V * H + V], flat[V * H + V], flat[V * H + V + 1]


def pack_step(flat: torch.Tensor, dW: torch.Tensor, db: torch.Tensor, loss_sum: torch.Tensor, n: int) -> None:
    """Copies one rank's contribution into ``flat`` (for gradients that were not produced in place)."""
    V, H = dW.shape
    flat[: V * H].copy_(dW.reshape(-1))
    flat[V * H: V * H + V].copy_(db)
    flat[V * H + V] = loss_sum
    flat[V * H + V + 1] = float(n)


def allreduce_step(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sums ``flat`` over the ranks; no-op for one rank."""
    if _world(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat

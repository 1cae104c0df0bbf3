"""Utterance sharding across the GPUs of one box and the single per-step exchange.

The path shards by utterance: joint, lattice and decode are independent per utterance;
only the joint's parameter gradients and the scalar loss couple them.  One all-reduce (sum) of the flat
fp32 buffer ``[dW (V*H) | db (V) | loss_sum | n_utterances]`` per step is the only collective; ``df`` and
``dg`` stay on the GPU that owns the utterance.  The reference has no distributed code at all
(SURVEY.md §2.1), so there is no reference interface to mirror here.
"""
from typing import List, Sequence, Tuple

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


def flat_size(V: int, H: int) -> int:
    return V * H + V + 2


def pack_step(flat: torch.Tensor, dW: torch.Tensor, db: torch.Tensor, loss_sum: torch.Tensor, n: int) -> None:
    """Writes this rank's contribution into ``flat`` (fp32, ``flat_size(V, H)`` elements) in place."""
    V, H = dW.shape
    flat[: V * H].copy_(dW.reshape(-1))
    flat[V * H: V * H + V].copy_(db)
    flat[V * H + V] = loss_sum
    flat[V * H + V + 1] = float(n)


def unpack_step(flat: torch.Tensor, V: int, H: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Views of (dW, db, loss_sum, n) inside ``flat``."""
    return flat[: V * H].view(V, H), flat[V * H: V * H + V], flat[V * H + V], flat[V * H + V + 1]


def allreduce_step(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sums ``flat`` over the ranks (NCCL over NVLink on GPUs, gloo in the CPU tests); no-op for one rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat

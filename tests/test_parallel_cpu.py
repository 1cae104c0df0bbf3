"""World-size-2 gloo test of the N>1 host logic (sharding + the single flat all-reduce) on CPU."""
import os

import hypothesis.strategies as st
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given

from myrtlespeech_b200 import parallel as P


@given(lens=st.lists(st.tuples(st.integers(1, 500), st.integers(0, 150)), min_size=1, max_size=40),
       world=st.integers(1, 8))
def test_shards_partition_the_batch_and_balance(lens, world):
    fl, yl = [a for a, _ in lens], [b for _, b in lens]
    shards = P.shard_utterances(fl, yl, world)
    assert len(shards) == world
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(len(lens)))
    cost = [t * (u + 1) for t, u in lens]
    loads = [sum(cost[i] for i in s) for s in shards]
    # greedy longest-first: no rank exceeds the mean by more than the largest single item
    assert max(loads) <= sum(cost) / world + max(cost)


def test_shard_rejects_bad_world():
    with pytest.raises(ValueError):
        P.shard_utterances([1], [1], 0)


def _worker(rank, world, port, V, H, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(100 + rank)
        dW = torch.randn(V, H, generator=gen); db = torch.randn(V, generator=gen)
        flat = torch.zeros(P.flat_size(V, H))
        P.pack_step(flat, dW, db, torch.tensor(float(rank + 1)), 3 + rank)
        P.allreduce_step(flat)
        if rank == 0:
            torch.save(flat, out)
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_world2(tmp_path):
    V, H, world = 7, 16, 2
    out = str(tmp_path / "flat.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, V, H, out), nprocs=world, join=True)
    flat = torch.load(out)
    want_dW = sum(torch.randn(V, H, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    dW, db, loss_sum, n = P.unpack_step(flat, V, H)
    assert torch.allclose(dW, want_dW, atol=1e-6)
    assert loss_sum.item() == 3.0 and n.item() == 7.0
    assert db.shape == (V,)


def test_allreduce_is_noop_without_process_group():
    flat = torch.arange(P.flat_size(2, 3), dtype=torch.float32)
    assert torch.equal(P.allreduce_step(flat.clone()), flat)

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


# ---- GradientReducer: every trainable gradient of the replicated model is reduced ---------------------------------------
CFG = """
alphabet: "abcdefg_";
input_features: 5;
rnn_t { encoder_hidden_size: 12; encoder_num_layers: 1; pred_embedding_size: 6;
        pred_hidden_size: 12; pred_num_layers: 1; joint_hidden_size: 16; }
rnn_t_loss { blank_index: 7; reduction: SUM; }
rnn_t_greedy_decoder { blank_index: 7; max_symbols_per_step: 2; }
"""


def _model_and_batch():
    from google.protobuf import text_format
    from myrtlespeech_b200.builders import speech_to_text as stt_builder
    from myrtlespeech_b200.protos import speech_to_text_pb2
    torch.manual_seed(5)
    stt = stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))
    g = torch.Generator().manual_seed(6)
    B, T, U = 6, 13, 4
    x = torch.randn(B, 1, 5, T, generator=g)
    y = torch.randint(0, 7, (B, U), generator=g, dtype=torch.int32)
    fl = torch.tensor([13, 13, 11, 9, 8, 6]); yl = torch.tensor([4, 2, 3, 4, 1, 2])
    return stt, x, y, fl, yl


def _loss_sum(stt, x, y, fl, yl, idx):
    """CPU test double for the CUDA loss (sum over the utterances ``idx``): torchaudio on the materialised joint."""
    import torchaudio
    i = torch.tensor(idx)
    tm, um = int(fl[i].max()), int(yl[i].max())       # torchaudio wants the padded sizes to equal the longest lengths
    xs, ys = x[i][..., :tm], y[i][:, :um]
    (out, out_lens), _ = stt.model(((xs, ys), (fl[i], yl[i])))
    return torchaudio.functional.rnnt_loss(out.materialize().float(), ys.int(), out_lens.int(), yl[i].int(), blank=7,
                                           reduction="sum")


def _reducer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stt, x, y, fl, yl = _model_and_batch()             # identical replicas, identical global batch
        shard = P.shard_utterances(fl.tolist(), yl.tolist(), world)[rank]
        red = P.GradientReducer(stt.parameters())
        opt = torch.optim.SGD(stt.parameters(), lr=0.05)
        for _ in range(2):
            red.zero()
            loss = _loss_sum(stt, x, y, fl, yl, shard)
            loss.backward()
            red.set_loss(loss, len(shard))
            red.all_reduce()
            red.wait()
            opt.step()
        torch.save(dict(sd=stt.state_dict(), loss=float(red.loss_sum), n=float(red.n_utterances), shard=shard),
                   f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_gradient_reducer_keeps_replicas_identical_world2(tmp_path):
    """Two ranks, utterances sharded by lattice size, two SGD steps: every parameter -- encoder and prediction network
    included, not only the joint -- is identical on both ranks and equal to a single-process run on the whole batch."""
    pytest.importorskip("torchaudio")
    out = str(tmp_path / "rank")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_reducer_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert sorted(r0["shard"] + r1["shard"]) == list(range(6)) and r0["n"] == 6.0
    assert r0["loss"] == r1["loss"]
    for k in r0["sd"]:
        assert torch.equal(r0["sd"][k], r1["sd"][k]), k
    # single process, whole batch
    stt, x, y, fl, yl = _model_and_batch()
    red = P.GradientReducer(stt.parameters())
    opt = torch.optim.SGD(stt.parameters(), lr=0.05)
    for _ in range(2):
        red.zero()
        loss = _loss_sum(stt, x, y, fl, yl, list(range(6)))
        loss.backward()
        red.set_loss(loss, 6)
        red.all_reduce()
        red.wait()
        opt.step()
    names = [k for k in r0["sd"] if "encoder" in k or "prediction" in k or "joint" in k]
    assert any("encoder" in k for k in names) and any("prediction" in k for k in names)
    for k in names:
        assert torch.allclose(r0["sd"][k], stt.state_dict()[k], rtol=1e-4, atol=1e-6), k
    assert abs(r0["loss"] - float(red.loss_sum)) < 1e-3 * abs(float(red.loss_sum))


def test_gradient_reducer_grads_are_views_of_the_flat_buffer():
    lin = torch.nn.Linear(3, 2)
    red = P.GradientReducer(lin.parameters())
    lin(torch.ones(4, 3)).sum().backward()
    assert red.flat[:6].view(2, 3).data_ptr() == lin.weight.grad.data_ptr()
    assert torch.equal(red.flat[:6], torch.full((6,), 4.0)) and torch.equal(red.flat[6:8], torch.full((2,), 4.0))
    red.zero()
    assert float(lin.weight.grad.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        P.GradientReducer([torch.nn.Parameter(torch.zeros(2, dtype=torch.float64))])

"""Two (or more) ranks over NCCL: a whole RNN-T model trained for a few SGD steps with the CUDA loss, utterances sharded by
lattice size, every parameter gradient all-reduced by parallel.GradientReducer.  Checks that the replicas stay bit-identical
and that they follow a single-process run on the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/ddp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist
from google.protobuf import text_format

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from myrtlespeech_b200 import parallel as P  # noqa: E402
from myrtlespeech_b200.builders import speech_to_text as stt_builder  # noqa: E402
from myrtlespeech_b200.protos import speech_to_text_pb2  # noqa: E402

CFG = """
alphabet: "abcdefghijklmnopqrstuvwxyz '_";
input_features: 40;
rnn_t { encoder_hidden_size: 128; encoder_num_layers: 2; pred_embedding_size: 64;
        pred_hidden_size: 128; pred_num_layers: 1; joint_hidden_size: 256; }
rnn_t_loss { blank_index: 28; reduction: SUM; }
rnn_t_greedy_decoder { blank_index: 28; max_symbols_per_step: 3; }
"""


def build():
    torch.manual_seed(11)
    return stt_builder.build(text_format.Merge(CFG, speech_to_text_pb2.SpeechToText()))


def batch():
    g = torch.Generator().manual_seed(12)
    B, T, U = 16, 120, 30
    x = torch.randn(B, 1, 40, T, generator=g)
    y = torch.randint(0, 28, (B, U), generator=g, dtype=torch.int32)
    fl = torch.randint(T // 2, T + 1, (B,), generator=g).sort(descending=True).values
    yl = torch.randint(U // 2, U + 1, (B,), generator=g)
    fl[0], yl[0] = T, U
    return x, y, fl, yl


def run(stt, x, y, fl, yl, idx, steps, reducer, n_total):
    opt = torch.optim.SGD(stt.parameters(), lr=1e-3)
    i = torch.tensor(idx)
    tm, um = int(fl[i].max()), int(yl[i].max())
    xs, ys = x[i][..., :tm].cuda(), y[i][:, :um].cuda()
    losses = []
    for _ in range(steps):
        reducer.zero()
        out, _ = stt.model(((xs, ys), (fl[i], yl[i])))
        loss = stt.loss(out, (ys, yl[i]))
        loss.backward()
        reducer.set_loss(loss, len(idx))
        reducer.all_reduce()
        reducer.wait()
        opt.step()
        losses.append(float(reducer.loss_sum) / n_total)
    return losses


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    x, y, fl, yl = batch()
    shards = P.shard_utterances(fl.tolist(), yl.tolist(), world)
    stt = build()
    losses = run(stt, x, y, fl, yl, shards[rank], 3, P.GradientReducer(stt.parameters()), len(fl))
    # replicas identical?
    flat = torch.cat([p.detach().reshape(-1) for p in stt.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        # single process on the whole batch (no process group use: world-1 semantics by a private group)
        solo = build()
        solo_group = dist.new_group([0])
        ref_losses = run(solo, x, y, fl, yl, list(range(len(fl))), 3, P.GradientReducer(solo.parameters(), group=solo_group), len(fl))
        ref_flat = torch.cat([p.detach().reshape(-1) for p in solo.parameters()])
        rel = float((flat - ref_flat).norm() / ref_flat.norm())
        print(f"world {world}: replicas bit-identical: {same}; mean loss per step sharded {['%.4f' % v for v in losses]} "
              f"vs single process {['%.4f' % v for v in ref_losses]}; parameters vs single process: rel diff {rel:.2e}; "
              f"shards by lattice rows {[sum(int(fl[i]) * (int(yl[i]) + 1) for i in s) for s in shards]}")
        assert same and rel < 1e-4 and all(abs(a - b) < 1e-3 * abs(b) for a, b in zip(losses, ref_losses))
        print("ddp_check: ok")
    else:
        dist.new_group([0])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

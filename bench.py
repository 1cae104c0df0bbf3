#!/usr/bin/env python
"""Headline benchmark: RNN-T joint + RNNTLoss forward + backward, utterances/s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload target|c2|c3] [--impl ours|reference]

One "step" = one pass of the hot path (joint -> log-softmax -> alpha/beta -> gradients df, dg, dW, db)
over one batch of synthetic utterances.  N > 1 is launched by torchrun, one rank per GPU; utterances are
sharded by rank (weak scaling: B per GPU fixed) and the only exchange is one NCCL all-reduce of the flat
fp32 buffer [dW | db | loss_sum | n] per step.

Printed JSON (rank 0): see the task contract.  `value` times the hot path with inputs resident in HBM;
`e2e` times the public module API (RNNTJoint -> RNNTLoss.forward -> backward) with pinned HOST inputs, so
the host->device copies and the device->host read of the loss are inside the timed region.
`roofline` is for the dominant kernel class, from CUDA events that bracket every kernel launch INSIDE the timed loop
(the per-class sums in `kernels` add up to the step time minus memsets and glue); `cpu_baseline` times the CPU
stand-in for the reference path (oracle/cpu_path.py) on a bounded sample with every host core.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, U, V, H, description)
    "target": (32, 500, 100, 1024, 1024, "north_star target / configs[3] per GPU: B=32 T=500 U=100 V=1024 H=1024 bf16"),
    "c3": (32, 400, 150, 1024, 1024, "configs[2] subword: B=32 T=400 U=150 V=1024 H=1024 bf16"),
    "c2": (32, 500, 100, 29, 512, "configs[1] chars: B=32 T=500 U=100 V=29 H=512 bf16"),
}
METRIC = "rnnt_joint_loss_fwd_bwd_utterances_per_s"
KCLASSES = ["hgen", "joint_fwd", "joint_dz", "joint_dh", "joint_dw", "lattice", "coefs", "misc", "joint_bwd_mega"]
NK = 16


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(tflops_burst=p["bf16_tflops"], tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm_gbs=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu summary
    (profiles/traffic.json, written from an `ncu --set full` capture); None if that kernel was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(workload, {}).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # median over the busier half of the samples (the idle tail before/after is not "under load")
        sm_sorted = sorted(sm, key=lambda x: -x)
        under = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] if power else sm
        return dict(sm_mhz=statistics.median(under) if under else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


def synth(B, T, U, V, H, seed, device):
    """f,g ~ N(0,1), W,bias ~ U(-1/sqrt(H), 1/sqrt(H)), rounded to bf16 once; y uniform non-blank."""
    import torch
    gen = torch.Generator().manual_seed(seed)
    f = torch.randn(B, T, H, generator=gen).bfloat16()
    g = torch.randn(B, U + 1, H, generator=gen).bfloat16()
    W = ((torch.rand(V, H, generator=gen) * 2 - 1) / H ** 0.5).bfloat16()
    bias = (torch.rand(V, generator=gen) * 2 - 1) / H ** 0.5
    y = torch.randint(0, V - 1, (B, U), generator=gen, dtype=torch.int32)
    fl = torch.full((B,), T, dtype=torch.int32)
    yl = torch.full((B,), U, dtype=torch.int32)
    return f, g, W, bias, y, fl, yl


def run_reference(args, B, T, U, V, H, desc, rank):
    """--impl reference: the CPU stand-in for the reference path (oracle/cpu_path.py), all host threads,
    each step a bounded sample (B_cpu utterances of the same T/U/V/H)."""
    if rank != 0:
        return
    import torch
    from oracle import cpu_path
    b_cpu = 1 if V * H >= 1 << 18 else 4
    r = cpu_path.time_steps(b_cpu, T, U, V, H, steps=args.steps, warmup=args.warmup, threads=os.cpu_count())
    out = {
        "impl": "reference", "metric": METRIC, "value": r["utt_per_s"], "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample_batch": b_cpu, "note": "reference snapshot has no RNN-T code; "
                   "torch eager joint + torchaudio.functional.rnnt_loss on host cores stands in (oracle/cpu_path.py)"},
        "cpu_baseline": {"value": r["utt_per_s"], "unit": "utterances/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["utt_per_s"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def decode_bench(dev):
    """BASELINE.json configs[4]: greedy decode, B=128 T=500 V=H=1024, max 4 symbols per frame, through the public
    RNNTGreedyDecoder (host f in pinned memory -> device, transcripts back as List[List[int]]).  Random-init single-layer
    LSTM prediction network (E=256, Hp=512); measured after the training-step timing, not part of `value`."""
    import torch
    from myrtlespeech_b200.model import RNNTJoint
    from myrtlespeech_b200.model.rnn_t import RNNT, RNNTPredictionNet
    from myrtlespeech_b200.post_process import RNNTGreedyDecoder
    B, T, V, H, S = 128, 500, 1024, 1024, 4
    g = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    model = RNNT(torch.nn.Identity(), RNNTPredictionNet(V, 256, 512, 1, H), RNNTJoint(H, V)).to(dev)
    dec = RNNTGreedyDecoder(V - 1, model, max_symbols_per_step=S)
    f_host = torch.randn(B, T, H, generator=g).bfloat16().pin_memory()
    lens = torch.full((B,), T, dtype=torch.int32)
    best, n_sym = None, 0
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = dec(f_host.to(dev, non_blocking=True), lens)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
        n_sym = sum(len(o) for o in out)
    return {"workload": "configs[4]: greedy decode B=128 T=500 V=1024 H=1024 max_symbols_per_step=4, LSTM prediction net "
                        "E=256 Hp=512 (random init)", "ms_per_batch": round(best * 1e3, 2),
            "utterances_per_s": round(B / best, 1), "symbols_emitted": n_sym, "decode_steps_upper_bound": int(n_sym / B) + T,
            "kernel": "greedy_decode_cluster_kernel (one launch per batch)", "timing": "host wall clock around "
            "RNNTGreedyDecoder.forward incl. the H2D copy of f and the D2H copy of the transcripts, best of 3"}


#: measured special-function rate on this pool's B200 (scripts/micro/tanh_flip.cu, 16 warps per SM, boost clock):
#: 30.8 tanh.approx per ns per SM = 16 per clock per SM, the MUFU pipe's width
MUFU_GOPS_PEAK = 30.78 * 148


def draw_batch(args, B, T, U, V, H, rank, world):
    """This rank's utterances.  Primary run: B full-length utterances per rank (weak scaling).  --ragged: ONE global
    batch of B * world utterances (same seed on every rank), T_b ~ U[T/2, T], U_b ~ U[U/2, U], split between the ranks
    by parallel.shard_utterances so that every rank gets the same lattice size sum T_b (U_b + 1), not the same count
    (SURVEY.md 8e); each rank pads to its own longest utterance, as the reference's collate does per batch."""
    import torch
    from myrtlespeech_b200 import parallel as par
    if not args.ragged:
        f, g, W, bias, y, fl, yl = synth(B, T, U, V, H, 1234 + rank, None)
        return f, g, W, bias, y, fl, yl, None
    Bg = B * world
    f, g, W, bias, y, _, _ = synth(Bg, T, U, V, H, 1234, None)
    gen = torch.Generator().manual_seed(4321)
    fl = torch.randint(T // 2, T + 1, (Bg,), generator=gen, dtype=torch.int32)
    yl = torch.randint(U // 2, U + 1, (Bg,), generator=gen, dtype=torch.int32)
    fl[0], yl[0] = T, U
    shards = par.shard_utterances(fl.tolist(), yl.tolist(), world)
    mine = sorted(shards[rank], key=lambda i: -int(fl[i]))    # sorted by length, data/batch.py:97-100
    idx = torch.tensor(mine)
    tm, um = int(fl[idx].max()), int(yl[idx].max())
    loads = [sum(int(fl[i]) * (int(yl[i]) + 1) for i in sh) for sh in shards]
    balance = dict(utterances_per_rank=[len(sh) for sh in shards], lattice_rows_per_rank=loads,
                   imbalance=round(max(loads) / (sum(loads) / world), 4))
    return (f[idx, :tm].contiguous(), g[idx, :um + 1].contiguous(), W, bias, y[idx, :um].contiguous(), fl[idx], yl[idx],
            balance)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="skip the configs[4] greedy-decode measurement")
    ap.add_argument("--recompute", action="store_true",
                    help="recompute the joint's logits and h in the backward pass instead of keeping them from the forward pass "
                         "(7 GB less memory per live graph at the target shape, about 10 %% slower)")
    ap.add_argument("--ragged", action="store_true",
                    help="SURVEY.md 8(d) secondary run: one global ragged batch, sharded by lattice size across the ranks")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B, T, U, V, H, desc = WORKLOADS[args.workload]

    if args.impl == "reference":
        # each CPU step of one target-shape utterance takes ~10 s: bound the run to 1 warm-up + 3 timed steps
        args.steps = 3 if args.steps is None else max(1, min(args.steps, 3))
        args.warmup = 1 if args.warmup is None else max(1, min(args.warmup, 1))
        run_reference(args, B, T, U, V, H, desc, rank)
        return

    args.steps = 50 if args.steps is None else args.steps
    args.warmup = 20 if args.warmup is None else max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import ctypes
    import myrtlespeech_b200 as M
    from myrtlespeech_b200 import _lib
    from myrtlespeech_b200 import parallel as par
    from myrtlespeech_b200.loss import RNNTLoss
    from myrtlespeech_b200.model import RNNTJoint

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a rank-asymmetric collective should fail in minutes, not after NCCL's default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    lib = _lib.load()
    peaks = load_peaks()
    from myrtlespeech_b200 import functional as Fk
    if args.recompute:
        Fk.set_keep_activations(False)
    kept_bytes = int(Fk._kept_bytes(B, T, U, V, H))

    f, g, W, bias, y, fl, yl, balance = draw_batch(args, B, T, U, V, H, rank, world)
    if args.ragged:
        desc += " (ragged: T_b ~ U[T/2, T], U_b ~ U[U/2, U]; one global batch sharded by lattice size)"
    Bl = f.shape[0]                      # utterances on this rank
    blank = V - 1
    fd, gd, yd = f.to(dev), g.to(dev), y.to(dev)
    # the joint's parameters are fp32 master weights (as in a model under bf16 autocast); their .grad are views into
    # the flat reduction buffer [dW | db | loss_sum | n], so the backward pass accumulates straight into it
    Wd = W.float().to(dev).requires_grad_(True)
    bd = bias.to(dev).requires_grad_(True)
    fd.requires_grad_(True); gd.requires_grad_(True)
    reducer = par.GradientReducer([Wd, bd])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def hot_step():
        reducer.wait()                   # the previous step's all-reduce (side stream) owns the buffer until it is done
        reducer.zero()
        loss = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, blank)
        total = loss.sum()
        total.backward()
        if world > 1:
            reducer.set_loss(total, Bl)
            reducer.all_reduce()
        fd.grad = gd.grad = None
        return total

    def barrier():
        reducer.wait()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_times = []   # per-step durations of the last timed() run (median and min are reported too)

    def timed(fn, steps):
        """Per-step CUDA events on the launch stream; L2 flushed between steps outside the events."""
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); reducer.wait(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        per_step = [a.elapsed_time(b) for a, b in evs]
        step_times[:] = per_step
        return sum(per_step)

    for _ in range(args.warmup):
        hot_step()
    barrier()
    lib.rnnt_debug_set(b"reset_launches", 0)
    # every kernel launch of the timed loop is bracketed by a pair of events on the launch stream (two event records per
    # launch, five launches per step): the per-class durations below are measured INSIDE the loop that produces `value`
    lib.rnnt_debug_set(b"time_kernels", 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(hot_step, args.steps)
    rank0_steps = list(step_times)
    launches = int(lib.rnnt_debug_get(b"launches"))
    kms = (ctypes.c_double * NK)(); kn = (ctypes.c_longlong * NK)()
    _lib.check(lib.rnnt_debug_kernel_times(kms, kn, NK))
    lib.rnnt_debug_set(b"time_kernels", 0)
    barrier()

    # ---- end to end through the module API with host buffers -------------------------------------
    joint = RNNTJoint(H, V)
    with torch.no_grad():
        joint.fc.weight.copy_(W.float()); joint.fc.bias.copy_(bias)
    e2e_reducer = par.GradientReducer(joint.parameters())
    loss_mod = RNNTLoss(blank=blank, reduction="sum")
    f_pin, g_pin, y_pin = f.pin_memory(), g.pin_memory(), y.pin_memory()
    # Double-buffered input pipeline, as a data loader with a prefetcher does it: step i's host->device copy is
    # issued on a copy stream while step i-1 computes; every step still copies its own inputs from pinned host
    # memory and reads its loss back to the host, all inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(fd), torch.empty_like(gd), torch.empty_like(yd), torch.cuda.Event()) for _ in range(2)]

    def prefetch(i):
        fb_, gb_, yb_, ev = bufs[i & 1]
        with torch.cuda.stream(copy_stream):
            fb_.copy_(f_pin, non_blocking=True); gb_.copy_(g_pin, non_blocking=True); yb_.copy_(y_pin, non_blocking=True)
            ev.record(copy_stream)

    def e2e_step(i, last):
        fb_, gb_, yb_, ev = bufs[i & 1]
        torch.cuda.current_stream().wait_event(ev)
        if not last:
            prefetch(i + 1)           # buffers (i+1)&1 were last read by step i-1, which has completed (loss.item())
        fx = fb_.detach().requires_grad_(True); gx = gb_.detach().requires_grad_(True)
        e2e_reducer.wait()
        e2e_reducer.zero()
        out = joint((fx, fl), (gx, yl + 1))
        loss = loss_mod(out, (yb_, yl))
        loss.backward()
        if world > 1:
            e2e_reducer.set_loss(loss, Bl)
            e2e_reducer.all_reduce()
        return float(loss.item())  # device -> host read of the step's result

    def e2e_run(steps):
        """One event pair around the whole pipelined loop, first copy included."""
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        copy_stream.wait_event(e0)
        prefetch(0)
        for i in range(steps):
            e2e_step(i, i == steps - 1)
        e2e_reducer.wait()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    e2e_run(3)
    barrier()
    e2e_ms = e2e_run(args.steps)
    barrier()
    if rank == 0:
        clocks = sampler.stop()

    # ---- the same kernels timed ALONE (one step after an idle pause), for comparison with isolated captures ----
    # (round 1 reported the roofline from such a pass; the in-loop figure above is the one that counts)
    iso = {}
    if rank == 0:
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            time.sleep(0.5)
            lib.rnnt_debug_set(b"time_kernels", 1)
            flush.zero_()
            hot_step_isolated = M.rnnt_joint_loss(fd, gd, Wd, bd, yd, fl, yl, blank).sum()
            hot_step_isolated.backward()
            fd.grad = gd.grad = None
            ims = (ctypes.c_double * NK)(); inn = (ctypes.c_longlong * NK)()
            _lib.check(lib.rnnt_debug_kernel_times(ims, inn, NK))
            lib.rnnt_debug_set(b"time_kernels", 0)
            cur = {name: ims[i] for i, name in enumerate(KCLASSES) if inn[i]}
            if best is None or sum(cur.values()) < sum(best.values()):
                best = cur
        iso = {k: round(v, 4) for k, v in best.items()}
    barrier()

    # ---- how many lattice tiles did the backward pass walk?  (tiles whose arc occupancies are all exactly zero contribute
    # exact zeros and are skipped; `prune = -1` walks every tile: timed here for comparison on a short loop) ----
    tiles = None
    no_skip_ms = None
    skip_ab_ms = None
    if rank == 0:
        from myrtlespeech_b200 import functional as Fn
        ws = Fn._ws_pool.get((dev, torch.cuda.current_stream(dev).cuda_stream))
        if ws is not None:
            n2 = (ctypes.c_int * 2)()
            if lib.rnnt_debug_read_active_tiles(ws.data_ptr(), Bl, fd.shape[1], gd.shape[1] - 1, V, H, n2) == 0 and n2[1] > 0:
                tiles = {"walked": int(n2[0]), "total": int(n2[1]), "fraction": round(n2[0] / n2[1], 4)}
    # A/B in alternating blocks (the part's thermal state drifts by several per cent over a run: a single block timed
    # after the main loop compares two different clocks, not two schedules).  EVERY rank runs it: hot_step contains the
    # gradient all-reduce, a collective all ranks must enter (rank 0's figures are the ones reported).
    for _ in range(10):          # back onto the power cap first: the isolated timings above let the part cool down
        hot_step()
    ab = {0: 0.0, -1: 0.0}
    for rnd in range(4):
        for mode in ((-1, 0) if rnd % 2 == 0 else (0, -1)):     # alternate the order: neither mode always follows a pause
            lib.rnnt_debug_set(b"prune", mode)
            for _ in range(2):
                hot_step()
            ab[mode] += timed(hot_step, 5)
    lib.rnnt_debug_set(b"prune", 0)
    no_skip_ms = ab[-1] / 20
    skip_ab_ms = ab[0] / 20
    barrier()

    n_rows_local = int((fl.long() * (yl.long() + 1)).sum())
    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(Bl), float(n_rows_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms = t.tolist()
    n_utt_total, n_rows_total = int(cnt[0].item()), int(cnt[1].item())

    if rank == 0:
        n_rows = n_rows_local                              # rank 0's rows: its kernels are the ones timed per class
        # ALGORITHMIC flops per kernel class (forward 2NHV, backward 4NHV; the logits recompute of the
        # backward pass is extra hardware work and is reported separately as hw_tflops)
        nhv = float(n_rows) * H * V
        flops = {"joint_fwd": 2.0 * nhv, "joint_dz": 0.0, "joint_dh": 2.0 * nhv, "joint_dw": 2.0 * nhv,
                 "joint_bwd_mega": 4.0 * nhv}
        # executed by the hardware: the mega-kernel recomputes the logits (6NHV) unless the forward pass kept them (4NHV)
        hw_flops = {"joint_fwd": 2.0 * nhv, "joint_dz": 2.0 * nhv, "joint_dh": 2.0 * nhv, "joint_dw": 2.0 * nhv,
                    "joint_bwd_mega": (4.0 if kept_bytes else 6.0) * nhv}
        exec_frac = tiles["fraction"] if tiles else 1.0      # share of the lattice the backward pass executes
        kernels = {}
        for i, name in enumerate(KCLASSES):
            if kn[i]:
                per_step = kms[i] / args.steps
                kernels[name] = {"launches_per_step": round(kn[i] / args.steps, 2), "ms_per_step": round(per_step, 4)}
                if name in flops and flops[name] > 0:
                    kernels[name]["tflops"] = round(flops[name] / (per_step * 1e-3) / 1e12, 1)
                    executed = hw_flops[name] * (exec_frac if name == "joint_bwd_mega" else 1.0)
                    kernels[name]["hw_tflops"] = round(executed / (per_step * 1e-3) / 1e12, 1)
        ksum = sum(k["ms_per_step"] for k in kernels.values())
        step_ms = ms_total / args.steps
        tensor_bound = V * H >= 1 << 18
        if tensor_bound:
            dom = max(flops, key=lambda k: kernels.get(k, {}).get("ms_per_step", 0.0))
            achieved = kernels[dom]["tflops"]
            roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tflops_sustained"],
                        "unit": "TFLOP/s", "frac": round(achieved / peaks["tflops_sustained"], 4),
                        "frac_of_burst": round(achieved / peaks["tflops_burst"], 4),
                        "hw_achieved": kernels[dom]["hw_tflops"],
                        "hw_frac": round(kernels[dom]["hw_tflops"] / peaks["tflops_sustained"], 4),
                        "hw_frac_of_burst": round(kernels[dom]["hw_tflops"] / peaks["tflops_burst"], 4),
                        "traffic": ncu_traffic(dom, args.workload + ("" if kept_bytes else "_recompute")),
                        "traffic_source": "profiles/traffic.json (ncu --set full capture of this kernel, per launch)",
                        "peak_source": peaks["source"] + ": the SUSTAINED cuBLAS figure, because the kernel's duration is the mean "
                                       f"over the {args.steps} launches of the timed loop (events around each launch, GPU under the "
                                       "power cap); frac_of_burst divides by the burst figure",
                        "algorithmic_flops_per_launch": flops[dom],
                        "isolated": {"ms": iso.get(dom), "achieved": round(flops[dom] / (iso[dom] * 1e-3) / 1e12, 1) if iso.get(dom) else None,
                                     "frac_of_burst": round(flops[dom] / (iso[dom] * 1e-3) / 1e12 / peaks["tflops_burst"], 4) if iso.get(dom) else None,
                                     "note": "the same kernel timed alone after an idle pause (best of 3), against the BURST "
                                             "peak: comparable with an isolated ncu capture, not with the in-loop figure"},
                        "executed_fraction": exec_frac,
                        "note": "achieved = algorithmic flops of the whole lattice (6NHV basis, logits recompute not counted) / mean "
                                "in-loop CUDA-event duration of the launch.  The backward pass walks only the lattice tiles with "
                                "non-zero arc occupancy (executed_fraction of them; the others contribute exact zeros): hw_achieved "
                                "counts the flops actually executed (with --recompute that includes the logits GEMM; by default "
                                "the forward pass keeps logits and h and the backward pass executes 4NHV)"}
        else:
            # V=29, H=512: 0.03 flop per byte of tanh input on the tensor side -- the path is bound by the special-function
            # work (one tanh per row and column of h in each pass, one exp2 per logit in each pass), not by tensor cores
            # or HBM.  Roofline: MUFU operations per step / time against the measured MUFU rate.
            nc = (V + 31) // 32 * 32
            mufu_ops = 2.0 * n_rows * H + 2.0 * n_rows * nc
            dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
            dom_ops = (n_rows * H + n_rows * nc)
            achieved = dom_ops / (kernels[dom]["ms_per_step"] * 1e-3) / 1e9
            roofline = {"bound": "mufu", "kernel": dom, "achieved": round(achieved, 1), "peak": round(MUFU_GOPS_PEAK, 1),
                        "unit": "Gop/s", "frac": round(achieved / MUFU_GOPS_PEAK, 4), "traffic": ncu_traffic(dom, args.workload),
                        "step_frac": round(mufu_ops / (step_ms * 1e-3) / 1e9 / MUFU_GOPS_PEAK, 4),
                        "peak_source": "measured: scripts/micro/tanh_flip.cu, 30.78 tanh.approx / ns / SM x 148 SMs "
                                       "(16 per clock per SM at the boost clock)",
                        "algorithmic_ops_per_launch": dom_ops,
                        "note": "special-function operations (N*H tanh + N*32 exp2 per pass) / mean in-loop duration; the "
                                "whole step's fraction is step_frac"}
        step_tflops = 6.0 * n_rows_total * H * V / (step_ms * 1e-3) / 1e12

        cpu = None
        if not args.no_cpu_baseline:
            from oracle import cpu_path
            b_cpu = 1 if tensor_bound else 4
            r = cpu_path.time_steps(b_cpu, T, U, V, H, steps=3, warmup=1, threads=os.cpu_count())
            cpu = {"value": round(r["utt_per_s"], 4), "unit": "utterances/s", "cores": r["cores"], "kind": "port",
                   "sample": r["sample"] + ", 1 warm-up step, torch.set_num_threads(os.cpu_count())"}

        decode = None
        if world == 1 and not args.no_decode:
            try:
                decode = decode_bench(dev)
            except Exception as e:  # the decode figure is an extra; never lose the headline line over it
                decode = {"error": repr(e)}

        h2d = f.numel() * 2 + g.numel() * 2 + y.numel() * 4
        out = {
            "metric": METRIC, "value": round(n_utt_total * args.steps / (ms_total * 1e-3), 2), "unit": "utterances/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(step_ms, 4),
            "ms_per_step_median": round(statistics.median(rank0_steps), 4), "ms_per_step_min": round(min(rank0_steps), 4),
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "B_per_gpu": B, "global_batch": n_utt_total, "T": T, "U": U, "V": V, "H": H,
                       "parallelism": f"utterance-sharded dp{world}, one all-reduce of [dW|db|loss|n] on a side stream; the "
                                      "backward pass accumulates dW/db into the reduction buffer through .grad views",
                       "l2": "flushed between timed steps (256 MiB write outside the per-step events)",
                       "timing": "sum of per-step CUDA-event durations on the launch stream, max over ranks"},
            "algorithmic_tflops": round(step_tflops, 1),
            "frac_of_bf16_peak": {"burst": round(step_tflops / world / peaks["tflops_burst"], 4),
                                  "sustained": round(step_tflops / world / peaks["tflops_sustained"], 4),
                                  "basis": "6*N*H*V algorithmic flops per step (recompute not counted)"},
            "e2e": {"value": round(n_utt_total * args.steps / (e2e_ms * 1e-3), 2), "unit": "utterances/s",
                    "ms_per_step": round(e2e_ms / args.steps, 4), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "api": "RNNTJoint -> RNNTLoss.forward(inputs, targets) -> backward; pinned host f/g/y copied every step on a "
                           "copy stream one step ahead (prefetch), loss.item() every step"},
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels": kernels,
            "activations": {"schedule": "kept" if kept_bytes else "recomputed", "kept_bytes_per_live_graph": kept_bytes,
                            "note": "kept: the forward pass also writes the logits (fp16) and h (bf16) of every lattice row and the "
                                    "backward pass streams them back; recomputed (--recompute): nothing is kept, the backward pass "
                                    "rebuilds h and the logits"},
            "backward_tiles": tiles,
            "ms_per_step_every_tile": None if no_skip_ms is None else round(no_skip_ms, 4),
            "value_every_tile": None if no_skip_ms is None else round(Bl * world / (no_skip_ms * 1e-3), 2),
            "ms_per_step_tile_list_ab": None if skip_ab_ms is None else round(skip_ab_ms, 4),
            "tile_skipping_note": "the backward pass skips lattice tiles whose arc occupancies are all exactly zero in fp32 (their "
                                  "gradient contribution is exactly zero, results are identical); *_every_tile is the same step with "
                                  "the skipping switched off and ms_per_step_tile_list_ab the default step, both measured on rank 0 "
                                  "after the main loop (and 10 untimed steps that put the part back on the power cap) in alternating blocks of "
                                  "5 steps, 4 blocks each, the order swapped every round, so that the two share one thermal state",
            "kernels_isolated_ms": iso,
            "kernels_sum_ms": round(ksum, 4),
            "unattributed_ms": round(step_ms - ksum, 4),
            "unattributed_note": "step time minus the kernel classes: four output memsets (df, dg, dW, db), the state "
                                 "prefix zero / save / restore copies, loss.sum and the autograd glue kernels, launch gaps",
            "cpu_baseline": cpu,
            "clocks": clocks,
            "decode": decode,
        }
        if balance is not None:
            out["ragged_balance"] = balance
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""BASELINE.json configs[4]: 1D EDM UNet bf16 training step (forward + backward, NCCL gradient all-reduce, Adam + EMA),
batch 64 per GPU ([64, 6, 4064] synthetic N(0,1) signals), dropout 0.1 active.  One process per GPU:

    python tools/bench_train.py [--batch 64] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_train.py

Prints ONE JSON line (rank 0): samples/s over all ranks (CUDA events, max over ranks), step time, a phase breakdown
(forward / backward / optimiser incl. all-reduce) and the achieved TFLOP/s against 3 x 28.436 GFLOP per sample
(SURVEY 8(d)).  This is a secondary measurement; bench.py stays on the sampling headline.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of  # noqa: E402
from tqdne_b200.config import MovingAverageEnvelopeConfig  # noqa: E402
from tqdne_b200.training import TrainStep1D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--length", type=int, default=4064)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--impl", default="engine", choices=["engine", "reference"],
                help="reference: the reference training step (autograd + torch.optim.Adam + EMA lerp) on the host CPUs")
ap.add_argument("--cpu-batch", type=int, default=2)
args = ap.parse_args()

if args.impl == "reference":
    # LightningEDM.step (edm.py:115-134) + Adam + EMA on the CPU: the unmodified reference module when /root/reference is
    # present (build container), else the oracle port (GPU box); a bounded sample of the batch-64 workload
    if int(os.environ.get("RANK", 0)) != 0:
        sys.exit(0)
    from oracle import reference_loader, torch_ref

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = MovingAverageEnvelopeConfig()
    ucfg = tq.get_1d_unet_config(cfg, 6, 6)
    shell = tq.LightningEDM(ucfg, {}, num_sampling_steps=18)
    sd = seeded_state_dict(shapes_of(shell), 0)
    B, L = args.cpu_batch, args.length
    x, cond = torch.randn(B, 6, L), torch.randn(B, 5)
    if reference_loader.available():
        ref = reference_loader.load()
        mod = ref.edm.LightningEDM(ucfg, {"learning_rate": 1e-4, "max_steps": 100000, "eta_min": 0.0}, num_sampling_steps=18)
        mod.load_state_dict(sd)
        mod.train()
        params = [p for p in mod.parameters() if p.requires_grad]
        loss_fn = lambda: mod.step({"signal": x, "cond": cond}, 0)  # noqa: E731
        kind = "reference"
    else:
        P = {k: v.clone().requires_grad_(not k.endswith("time_embed.W")) for k, v in sd.items()}
        params = [v for v in P.values() if v.requires_grad]

        def loss_fn():
            sigma = (torch.randn(B) * 1.2 - 1.2).exp()
            pred = torch_ref.denoise(P, ucfg, x + torch.randn_like(x) * sigma[:, None, None], sigma, cond)
            return ((pred - x) ** 2 * ((sigma**2 + 0.25) / (sigma * 0.5) ** 2)[:, None, None]).mean()
        kind = "port"
    opt = torch.optim.Adam(params, lr=1e-4)
    ema = [p.detach().clone() for p in params]

    def ref_step():
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        opt.step()
        torch._foreach_lerp_(ema, [p.detach() for p in params], 1e-3)
        return float(loss)

    ref_step()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 5)):
        loss = ref_step()
    dt = (time.perf_counter() - t0) / max(1, args.steps // 5)
    print(json.dumps({"impl": "reference", "metric": "training samples/sec (1D EDM UNet, fwd + bwd + Adam + EMA)", "value": B / dt,
                      "unit": "samples/s", "ms_per_step": dt * 1e3, "dtype": "f32",
                      "cpu_baseline": {"value": B / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": kind,
                                       "sample": f"{B} x [6, {L}] per step"}, "loss": loss}), flush=True)
    sys.exit(0)

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)          # NCCL's version banner goes to stderr, stdout keeps the one JSON line
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)

cfg = MovingAverageEnvelopeConfig()
edm = tq.LightningEDM(tq.get_1d_unet_config(cfg, 6, 6), {"learning_rate": 1e-4, "max_steps": 100000, "eta_min": 0.0},
                      num_sampling_steps=18)
edm.load_state_dict(seeded_state_dict(shapes_of(edm), 0))
edm.to(dev)
B, L = args.batch, args.length
step = TrainStep1D(edm, B, L, lr=1e-4, max_steps=100000)
g = torch.Generator(device=dev).manual_seed(100 + rank)
signal = torch.randn(B, 6, L, device=dev, generator=g)
cond = torch.randn(B, 5, device=dev, generator=g)


def one_step():
    loss = step.forward_backward(signal, cond)
    step.optimizer_step(world)
    return loss


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(max(3, args.warmup)):
    loss = one_step()
assert bool(torch.isfinite(loss)), "non-finite loss"
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(args.steps):
    loss = one_step()
e1.record()
barrier()
wall = time.perf_counter() - t0
ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms_step = float(max(ms[0], ms[1])) / args.steps   # the optimiser's host-side work is inside the step: wall bounds it

# phase breakdown on rank 0 (synchronised wall clock, 3 repetitions)
phases = {}
if rank == 0:
    def timed(fn, reps=3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps * 1e3

    def fwd_only():
        for f in step.fwd:
            f()

    def bwd_only():
        for b in step.bwd:
            b()

    phases = {"forward_ms": timed(fwd_only), "backward_ms": timed(bwd_only)}
if world > 1:
    dist.barrier()
if rank == 0:
    phases["optimizer_ms"] = None
    flop_per_sample = 3 * 28.436e9 * (L / 4064)
    sps = B * world / (ms_step / 1e3)
    print(json.dumps({"metric": "training samples/sec (1D EDM UNet, fwd + bwd + all-reduce + Adam + EMA)", "value": sps,
                      "unit": "samples/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms_step, "dtype": "bf16",
                      "config": {"workload": "BASELINE.json configs[4]: 1D EDM UNet training step", "batch_per_gpu": B,
                                 "length": L, "dropout": step.p_drop, "params": int(step.store.n)},
                      "achieved_tflops_per_gpu": sps / world * flop_per_sample / 1e12, "loss": float(loss), **phases}), flush=True)
if world > 1:
    dist.destroy_process_group()

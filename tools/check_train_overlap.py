"""Data-parallel training step, overlapped vs single gradient all-reduce (torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_train_overlap.py

Each rank trains the 1D EDM UNet for a few steps on its own seeded batches, once with the early all-reduce of the late
gradient bucket (TQ_TRAIN_OVERLAP=1, the default) and twice with one all-reduce after the backward pass (=0), from the same
initial weights.  The all-reduced gradients of the first step must agree between the modes up to the order of the weight
gradient's fp32 atomics; the master parameters after a few steps must agree as well as two runs of ONE mode do (Adam amplifies
that noise on near-zero gradients), and be bit-identical across the ranks in both modes.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from tqdne_b200.config import MovingAverageEnvelopeConfig  # noqa: E402
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
cfg = MovingAverageEnvelopeConfig()
B, L, steps = 8, 1024, 5


def run(overlap: str) -> torch.Tensor:
    os.environ["TQ_TRAIN_OVERLAP"] = overlap
    edm = tq.LightningEDM(tq.get_1d_unet_config(cfg, 6, 6), {"learning_rate": 1e-3, "max_steps": 100})
    edm.load_state_dict(seeded_state_dict(shapes_of(edm), 0))
    edm.cuda()
    torch.manual_seed(1000 + rank)            # sigma / noise draws of this rank
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    g_first = None
    for i in range(steps):
        batch = {"signal": torch.randn(B, 6, L, device="cuda", generator=g), "cond": torch.randn(B, 5, device="cuda", generator=g)}
        loss = edm.training_step(batch)
        if i == 0:
            g_first = edm.__dict__["_tq_train_last"].store.G.clone()   # the summed gradients Adam has just consumed
    ts = edm.__dict__["_tq_train_last"]
    assert bool(torch.isfinite(loss))
    return ts.store.P.clone(), g_first


(p0, g0), (p0b, _), (p1, g1) = run("0"), run("0"), run("1")
grel = float((g1 - g0).norm() / g0.norm())
rel = float((p1 - p0).norm() / p0.norm())
noise = float((p0b - p0).norm() / p0.norm())   # run-to-run spread of ONE mode: the weight gradient's fp32 atomics, amplified by Adam
gathered = [torch.empty_like(p1) for _ in range(world)]
dist.all_gather(gathered, p1)
same1 = all(torch.equal(gathered[0], t) for t in gathered)
dist.all_gather(gathered, p0)
same0 = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    print(f"all-reduced gradients of step 1, overlapped vs single all-reduce: rel-L2 {grel:.2e}", flush=True)
    assert grel < 1e-4
    print(f"after {steps} steps, rel-L2 of the master parameters: overlapped vs single all-reduce {rel:.2e}; single vs single "
          f"(run-to-run) {noise:.2e}; ranks identical: overlapped {same1}, single {same0}", flush=True)
    assert rel < 3 * noise + 1e-6 and same1 and same0
dist.destroy_process_group()

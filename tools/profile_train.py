"""One training step of the 1D UNet (batch 64, L 4064) inside an NVTX range for ncu:
  ncu --nvtx --nvtx-include "train_step/" --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_train.csv python tools/profile_train.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of  # noqa: E402
from tqdne_b200.config import MovingAverageEnvelopeConfig  # noqa: E402
from tqdne_b200.training import TrainStep1D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = MovingAverageEnvelopeConfig()
edm = tq.LightningEDM(tq.get_1d_unet_config(cfg, 6, 6), {}, num_sampling_steps=18)
edm.load_state_dict(seeded_state_dict(shapes_of(edm), 0))
edm.cuda()
step = TrainStep1D(edm, B, 4064)
x, c = torch.randn(B, 6, 4064, device="cuda"), torch.randn(B, 5, device="cuda")
for _ in range(2):
    step.forward_backward(x, c)
    step.optimizer_step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("train_step")
step.forward_backward(x, c)
step.optimizer_step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")

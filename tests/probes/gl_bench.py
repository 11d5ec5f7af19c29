"""Griffin-Lim kernel timing + parity on decoded-spectrogram-like inputs: fused fp64 / fp32 kernels, 768 items, 128 iterations.

    python tests/probes/gl_bench.py [items=768] [n_iter=128]
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import griffinlim_ref  # noqa: E402
from tqdne_b200.representation import LogSpectrogram  # noqa: E402

items = int(sys.argv[1]) if len(sys.argv) > 1 else 768
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = torch.Generator().manual_seed(0)
rep = (torch.tanh(torch.randn(items // 3, 3, 128, 128, generator=g) * 0.3) * 0.6 - 0.2).cuda()
out = {"items": items, "n_iter": n_iter}
for prec in ("fp64", "fp32"):
    for legacy in ("0", "1"):
        os.environ["TQ_GL_LEGACY"] = legacy
        ls = LogSpectrogram(stft_channels=256, hop_size=32, precision=prec)
        ls.n_iter = n_iter
        w = ls.invert_representation_device(rep)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            w = ls.invert_representation_device(rep)
        e1.record()
        torch.cuda.synchronize()
        out[f"{prec}_{'unfused' if legacy == '1' else 'fused'}_ms"] = e0.elapsed_time(e1) / 3
        if prec == "fp64" and legacy == "0":
            ref = griffinlim_ref.logspec_inverse(rep[:1].cpu().numpy(), n_iter=n_iter, precision="fp64")
            got = w[:1].cpu().numpy()
            out["fp64_fused_vs_numpy_fp64"] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
os.environ.pop("TQ_GL_LEGACY")
print(json.dumps(out))

"""Where does a sampler NFE spend time outside the UNet plan replay?  CUDA-event timing of nested variants."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import tqdne_b200 as tq  # noqa: E402
from bench import build_state_dict, cond_grid  # noqa: E402
from tqdne_b200 import _lib  # noqa: E402
from tqdne_b200.config import LatentSpectrogramConfig  # noqa: E402
from tqdne_b200.engine import current_stream_ptr, tq_dtype  # noqa: E402
from tqdne_b200.lowering import get_unet_plan  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = LatentSpectrogramConfig()
enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, 8, 8), {}, autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
edm.load_state_dict(build_state_dict(edm))
edm.eval().cuda().set_engine_precision("bf16")
lib = _lib.lib()
s = torch.cuda.Stream()
ITER = 20


def timed(name, fn):
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        t0 = time.perf_counter()
        e0.record(s)
        for _ in range(ITER):
            fn()
        e1.record(s)
        t_issue = time.perf_counter() - t0
        s.synchronize()
    print(f"{name:45s} {e0.elapsed_time(e1) / ITER:8.3f} ms/iter (CPU issue {t_issue / ITER * 1e3:6.3f} ms/iter)", flush=True)


with torch.cuda.stream(s):
    plan = get_unet_plan(edm.unet, B, (32, 32), uniform_t=True)
    plan.plan.enable_graph(True)
    plan.set_cond(torch.from_numpy(cond_grid(B)).cuda())
    x = torch.randn(B, 1024, 8, device="cuda", dtype=torch.float64)
    x1 = torch.empty_like(x)
    d = torch.empty_like(x)
    tdev = torch.full((64,), 0.3, device="cuda")
    xin, dt_ = plan.xin.t, tq_dtype(plan.act_dtype)
    F, Cf, Cpad = plan.out.t, plan.out.C, plan.cin_pad
    NP = B * 1024


def euler():
    _lib.check(lib.tq_edm_euler(x.data_ptr(), F.data_ptr(), Cf, d.data_ptr(), x1.data_ptr(), xin.data_ptr(), dt_, NP, 8, Cpad,
                                0.5, 0.5, 1.0, -0.1, 1.0, 1, plan.t.data_ptr(), 0.3, current_stream_ptr()), "euler")


def heun():
    _lib.check(lib.tq_edm_heun(x.data_ptr(), x1.data_ptr(), d.data_ptr(), F.data_ptr(), Cf, xin.data_ptr(), dt_, NP, 8, Cpad,
                               0.5, 0.5, 1.0, -0.1, 1.0, 1, plan.t.data_ptr(), 0.3, current_stream_ptr()), "heun")


timed("plan.run (graph)", lambda: plan.run())
timed("t.copy_ + plan.run", lambda: (plan.t.copy_(tdev[3:4]), plan.run()))
timed("euler kernel alone", euler)
timed("heun kernel alone", heun)
timed("t.copy_ + plan.run + euler", lambda: (plan.t.copy_(tdev[3:4]), plan.run(), euler()))
timed("plan.run + euler (t written by the kernel)", lambda: (plan.run(), euler()))
# the plan's kernels + the update kernel captured into ONE graph (tq_plan_run records into a caller-owned capture)
g1 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    s.synchronize()
    with torch.cuda.graph(g1, stream=s):
        plan.run()
        euler()
timed("ONE graph = plan + euler", lambda: g1.replay())
g2 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    s.synchronize()
    with torch.cuda.graph(g2, stream=s):
        plan.run()
        euler()
        plan.run()
        heun()
ITER = 10
timed("ONE graph = 2 x (plan + update), per pair", lambda: g2.replay())
ITER = 20
cond = torch.from_numpy(cond_grid(B)).cuda()
noise = torch.randn(B, 8, 32, 32, device="cuda", dtype=torch.float64)
ae, edm.autoencoder = edm.autoencoder, None
ITER = 3
timed("edm.sample latent (49 NFE)", lambda: edm.sample((B, 8, 32, 32), cond=cond, noise=noise))

"""Which kernels are power-capped?  Runs one op type back to back for ~2 s and reports the median SM clock, the mean
power and the sustained rate (nvidia-smi sampled every 100 ms).  The whole bench step is power-capped on B200
(sw_power_cap, SM clock below max), so time tracks energy, not only idle gaps."""
import math
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from tqdne_b200.engine import Act, Plan, pack_conv  # noqa: E402

dev = torch.device("cuda")


class Smi:
    def __enter__(self):
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100",
                                   "-i", "0"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        return self

    def __exit__(self, *a):
        self.p.terminate()
        out, _ = self.p.communicate(timeout=5)
        rows = [[float(v) for v in ln.split(",")] for ln in out.strip().splitlines() if ln.count(",") == 1]
        rows = rows[len(rows) // 3:]  # drop the ramp
        self.mhz = sorted(r[0] for r in rows)[len(rows) // 2] if rows else float("nan")
        self.watt = sum(r[1] for r in rows) / len(rows) if rows else float("nan")


def conv_plan(N, sp, cin, cout, k, reps=20, **kw):
    H, W = sp
    x = torch.randn(N * H * W * cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev) / math.sqrt(cin * k * k)
    b = torch.randn(cout, device=dev)
    plan = Plan(dev, torch.bfloat16)
    xa = Act(x, N, H, W, cin)
    pc = pack_conv(w, b, [cin], torch.bfloat16)
    for _ in range(reps):
        plan.conv(pc, [xa], dims=2, stats=True, **kw)
    plan.enable_graph(True)
    return plan, 2.0 * N * H * W * cout * cin * k * k * reps


def sustained(name, plan, flop, seconds=2.0):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            plan.run()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0
        with Smi() as smi:
            t0 = time.perf_counter()
            e0.record(s)
            while time.perf_counter() - t0 < seconds:
                for _ in range(10):
                    plan.run()
                n += 10
                s.synchronize()
            e1.record(s)
            s.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:44s} {flop / ms / 1e9:8.1f} TF/s sustained | SM {smi.mhz:6.0f} MHz | {smi.watt:6.0f} W", flush=True)


def gn_plan(N, sp, C, reps=20):
    H, W = sp
    x = torch.randn(N * H * W * C, device=dev).to(torch.bfloat16)
    plan = Plan(dev, torch.bfloat16)
    xa = Act(x, N, H, W, C)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for _ in range(reps):
        plan.groupnorm([xa], g, b, True)   # stand-alone statistics pass + apply (no producing conv here)
    plan.enable_graph(True)
    return plan, N * H * W * C * 2 * reps


def sustained_bytes(name, plan, nbytes, seconds=2.0):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            plan.run()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0
        with Smi() as smi:
            t0 = time.perf_counter()
            e0.record(s)
            while time.perf_counter() - t0 < seconds:
                for _ in range(10):
                    plan.run()
                n += 10
                s.synchronize()
            e1.record(s)
            s.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:44s} {3 * nbytes / ms / 1e6:8.1f} GB/s (stats read + apply read/write) | SM {smi.mhz:6.0f} MHz | {smi.watt:6.0f} W",
          flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "gn":
        for nm, a in [("gn 128ch @32^2 (67 MB)", (256, (32, 32), 128)), ("gn 256ch @32^2 (134 MB)", (256, (32, 32), 256)),
                      ("gn 512ch @8^2 (17 MB)", (256, (8, 8), 512)), ("gn 512ch @4^2 (4 MB)", (256, (4, 4), 512))]:
            plan, nb = gn_plan(*a)
            sustained_bytes(nm, plan, nb)
        sys.exit(0)
    a = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    b = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)

    class MM:
        def run(self):
            for _ in range(4):
                torch.matmul(a, b)
    with torch.cuda.stream(torch.cuda.Stream()):
        pass
    sustained("cuBLAS bf16 8192^3 (torch.matmul)", MM(), 4 * 2.0 * 8192**3)
    for nm, args, kw in [
        ("conv 128->128 3x3 @32^2  BN=128 CG=2", (256, (32, 32), 128, 128, 3), {}),
        ("conv 384->128 3x3 @32^2  BN=128 CG=2", (256, (32, 32), 384, 128, 3), {}),
        ("conv 256->256 3x3 @32^2  BN=256 CG=2", (256, (32, 32), 256, 256, 3), {}),
        ("conv 256->256 3x3 @16^2  BN=256 CG=2", (256, (16, 16), 256, 256, 3), {}),
        ("conv 512->512 3x3 @8^2   BN=256 CG=2", (256, (8, 8), 512, 512, 3), {}),
        ("conv 512->512 3x3 @4^2   BN=128 CG=2", (256, (4, 4), 512, 512, 3), {}),
    ]:
        plan, fl = conv_plan(*args, **kw)
        sustained(nm, plan, fl)

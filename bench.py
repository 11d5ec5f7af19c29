#!/usr/bin/env python
"""bench.py -- waveforms/sec of the HighFEM latent EDM sampling hot path on N B200s (one process per GPU).

A "step" is one pass of the hot path over one batch: 25 Heun steps (49 denoiser calls) of the latent
UNet -> autoencoder decode -> log-spectrogram inverse (128 Griffin-Lim iterations) -> [B, 3, 4064] waveforms,
B = 256 per GPU (BASELINE.json configs[1]); the batch shards over GPUs as independent samples (weak scaling)
with a final gather of the waveforms.  Random-init weights of the named architecture, synthetic conditioning.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 ... bench.py --gpus 8 ...
    python bench.py --impl reference --steps 2 --warmup 1     # the reference algorithm on the host CPUs

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event time, max over ranks.
`e2e`: the same pass through the public API from pinned HOST buffers (cond + noise H2D, waveforms D2H).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NFE_STEPS = 25
FLOP_PER_WAVEFORM = 49 * 16.979e9 + 27.206e9  # SURVEY 8.1 / BASELINE.md section 3 (conv+attn+linear, 2*MAC)
METRIC = "waveforms/sec (3-comp, 25 Heun steps, latent EDM + decode + Griffin-Lim)"


def peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------
# model construction (shared by both arms): named architecture, seeded synthetic weights
# ------------------------------------------------------------------------------------------------------
def build_state_dict(edm, seed=0):
    from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of

    sd = seeded_state_dict(shapes_of(edm), seed)
    # keep the random-init decoder's output inside the normalised log-spectrogram range [-1, 1]
    for k in ("autoencoder.decoder.output_layer.weight", "autoencoder.decoder.output_layer.bias"):
        sd[k] = sd[k] * 0.05
    return sd


def cond_grid(n: int):
    """magnitude x hypocentral distance x vs30 grid, depth 10 km, gap 130 deg, z-scored with the reference's
    dataset statistics (tqdne/generate_waveforms.py:128-159); feature order (dist, mag, vs30, depth, gap)."""
    import numpy as np

    mags = np.linspace(4.5, 7.5, 16)
    dists = np.linspace(10.0, 200.0, 16)
    vs30s = np.linspace(150.0, 900.0, 8)
    g = np.stack(np.meshgrid(dists, mags, vs30s, indexing="ij"), -1).reshape(-1, 3)
    g = np.tile(g, (max(1, -(-n // len(g))), 1))[:n]
    raw = np.concatenate([g, np.full((n, 1), 10.0), np.full((n, 1), 130.0)], axis=1)
    stats = np.array([[101.29891904350877, 40.78415968551517], [4.801697862929673, 0.7146698731358634],
                      [384.7045105848187, 220.11269086015872], [38.359214998072, 22.472499592355014],
                      [129.92139043457396, 89.69479051949207]])
    return ((raw - stats[:, 0]) / stats[:, 1]).astype(np.float32)


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.

    Preferred source: NVML in-process (pynvml) from a background thread, one sample every 100 ms -- each sample is a few
    microsecond-scale driver queries.  An `nvidia-smi -lms 200` child process (the fallback when pynvml is missing) was
    measurably intrusive: the resident region, sampled, came out 3-10 % slower than the unsampled end-to-end region on
    some boxes.  Samples carry time stamps and are filtered to the timed window."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.thread, self.rows, self.stop_flag = gpu_index, None, None, [], False

    # ---- NVML thread
    def _nvml_loop(self, nv, handle):
        reasons = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                    else nv.nvmlClocksThrottleReasonHwSlowdown),
                   ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                   getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0))),
                   ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                   getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0))),
                   ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0)))]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                mask = int(get_reasons(handle))
                self.rows.append((time.time(), sm, mx, pw, [nm for nm, bit in reasons if bit and mask & bit]))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.idx
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "500", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self, t_begin: float | None = None, t_end: float | None = None) -> dict:
        """Median SM clock / throttle reasons of the samples taken inside [t_begin, t_end] (time.time() stamps of the
        timed region).  The sampler is started BEFORE the warm-up so that its start-up is outside the timed region."""
        import datetime

        rows = []   # (inside the timed window, sm, max sm, power, active reasons)
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            for ts, sm, mx, pw, rs in self.rows:
                rows.append((t_begin is None or t_begin - 0.05 <= ts <= t_end + 0.05, sm, mx, pw, rs))
            source = "nvml"
        elif self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
                out = ""
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in out.strip().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    vals = (float(f[1]), float(f[2]), float(f[3]))
                except ValueError:
                    continue
                inside = True
                if t_begin is not None:
                    try:
                        ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        inside = t_begin - 0.05 <= ts <= t_end + 0.05
                    except ValueError:
                        inside = True
                rows.append((inside, *vals, [nm for nm, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            source = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        use = [r for r in rows if r[0]] or rows   # an unparsable / shifted clock must not leave the line without clocks
        sm, mx, pw = [r[1] for r in use], [r[2] for r in use], [r[3] for r in use]
        reasons = {nm for r in use for nm in r[4]}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons), "source": source}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------
def _gl_item(args):
    from oracle import griffinlim_ref

    return griffinlim_ref.logspec_inverse(args, n_iter=128, precision="fp64")


class CpuPipeline:
    """Latent pipeline on the CPU: the unmodified reference modules when /root/reference is present (build
    container), else the oracle port (GPU box).  Griffin-Lim fans out over a process pool like the reference's
    pathos pool (tqdne/representation.py:128-129,156-157); librosa is absent everywhere, so the restated
    griffinlim is used in both cases."""

    def __init__(self, batch: int):
        import torch

        import tqdne_b200 as tq
        from oracle import reference_loader
        self.torch = torch
        self.batch = batch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        from types import SimpleNamespace

        cfg = SimpleNamespace(features_keys=("dist", "mag", "vs30", "depth", "gap"), channels=3, latent_channels=8)
        self.enc_cfg, self.dec_cfg = tq.get_2d_autoencoder_configs(cfg)
        self.unet_cfg = tq.get_2d_unet_config(cfg, 8, 8)
        shell = tq.LightningEDM(self.unet_cfg, {}, autoencoder=tq.LightningAutoencoder(self.enc_cfg, self.dec_cfg, {}))
        self.sd = build_state_dict(shell)
        self.kind = "port"
        self.ref_edm = None
        if reference_loader.available():
            ref = reference_loader.load()
            ae = ref.autoencoder.LightningAutoencoder(self.enc_cfg, self.dec_cfg, {})
            self.ref_edm = ref.edm.LightningEDM(self.unet_cfg, {}, num_sampling_steps=NFE_STEPS, autoencoder=ae)
            self.ref_edm.load_state_dict(self.sd)
            self.ref_edm.eval()
            self.kind = "reference"
        import numpy as np

        self.cond = torch.from_numpy(cond_grid(batch))
        self.np = np
        import multiprocessing as mp

        self.pool = mp.get_context("fork").Pool(min(self.cores, 3 * batch))

    def step(self, seed: int):
        torch = self.torch
        from oracle import torch_ref

        gen = torch.Generator().manual_seed(seed)
        noise = torch.randn((self.batch, 8, 32, 32), generator=gen, dtype=torch.float64)
        with torch.no_grad():
            if self.ref_edm is not None:
                sig = self.ref_edm.edm.sampling_sigmas(NFE_STEPS)
                lat = self.ref_edm.sample_deterministically(noise * sig[0], sig, None, self.cond)
                rep = self.ref_edm.autoencoder.decode(lat.to(torch.float32))
            else:
                sig = torch_ref.sampling_sigmas(NFE_STEPS)
                lat = torch_ref.heun_sample(self.sd, self.unet_cfg, noise * sig[0], sig, self.cond)
                rep = torch_ref.decoder_forward(self.sd, self.dec_cfg, lat.float(), prefix="autoencoder.decoder.")
        items = rep.numpy().reshape(-1, 1, 128, 128)
        waves = self.pool.map(_gl_item, list(items))
        return self.np.concatenate(waves).reshape(self.batch, 3, -1)

    def close(self):
        self.pool.close()


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    batch = args.cpu_batch
    pipe = CpuPipeline(batch)
    for i in range(args.warmup):
        pipe.step(100 + i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        w = pipe.step(i)
    dt = time.perf_counter() - t0
    pipe.close()
    assert w.shape == (batch, 3, 4064)
    val = batch * args.steps / dt
    sample = f"{batch} waveforms per step: full 25-step Heun (49 NFE) + decode + 128-iteration Griffin-Lim"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "waveforms/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HighFEM latent EDM: latent UNet Heun sampling + autoencoder decode + log-spectrogram "
                               "inverse (bounded CPU sample of the batch-256 workload)", "batch_per_step": batch,
                   "heun_steps": NFE_STEPS, "nfe": 2 * NFE_STEPS - 1, "weights": "random-init (seeded)"},
        "cpu_baseline": {"value": val, "unit": "waveforms/s", "cores": pipe.cores, "kind": pipe.kind, "sample": sample},
        "e2e": {"value": val, "unit": "waveforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.configs != "none":
        # BASELINE.md section 4, primary CPU config: the reference's 1D sampler, batch 4, 18 Heun steps, on the host cores
        line["configs"] = {"cfg0": {"workload": "1D EDM UNet (train_1d_edm config), batch 4, 18 Heun steps + envelope inverse",
                                    **cpu_cfg0()}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def kernel_breakdown(edm, batch, iters=3):
    """Per-kernel time of one denoiser call INSIDE the graph that the timed region replays.

    The plan is replayed op by op with a CUDA event after every launch (that gives each kernel's share; the event gaps
    make the op-by-op sum a few per cent longer than the graph), and the captured graph itself is timed over `iters` x 8
    replays; every kind's time is its op-by-op time minus its launches' share of (op-by-op sum - graph time) -- the event gap
    is a per-launch cost -- so the times are those inside the graph that was timed and add up to it.  profiles/ holds the ncu launch list of the same call for a cross-check of the shares."""
    import torch

    from tqdne_b200.lowering import get_unet_plan

    plan = get_unet_plan(edm.unet, batch, (32, 32), uniform_t=True)
    p = plan.plan
    n = p.num_ops
    names = p.op_names()
    meta = p.op_meta
    assert len(meta) == n, (len(meta), n)
    s = torch.cuda.Stream()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(iters)]
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 8 * iters
    with torch.cuda.stream(s):
        for _ in range(2):
            p.run_range(0, n)
        for it in range(iters):
            ev[it][0].record(s)
            for i in range(n):
                p.run_range(i, i + 1)
                ev[it][i + 1].record(s)
        p.enable_graph(True)
        for _ in range(2):
            p.run()
        g0.record(s)
        for _ in range(reps):
            p.run()
        g1.record(s)
    s.synchronize()
    graph_ms = g0.elapsed_time(g1) / reps
    agg = {}
    for i in range(n):
        ms = sum(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters)) / iters
        kind = names[i].split("<")[0].split(" ")[0]
        a = agg.setdefault(kind, {"ms": 0.0, "launches": 0, "flops": 0, "bytes": 0})
        a["ms"] += ms
        a["launches"] += 1
        a["flops"] += meta[i][1]
        a["bytes"] += meta[i][2]
    op_sum = sum(a["ms"] for a in agg.values())
    # the event after every launch costs a roughly constant gap PER LAUNCH (not per microsecond of kernel): remove each
    # kind's launches' share of (op-by-op sum - graph time); the per-kind times then add up to the timed graph
    gap = max(0.0, op_sum - graph_ms) / n
    for a in agg.values():
        a["ms_op_by_op"] = a["ms"]
        a["ms"] = max(a["ms"] - gap * a["launches"], 0.25 * a["ms"])
    scale = graph_ms / sum(a["ms"] for a in agg.values())   # exact closure (the floor above can leave a few microseconds)
    for a in agg.values():
        a["ms"] *= scale
    return agg, graph_ms, op_sum, names


def source_sha16(*rel_paths, extra: str = "") -> str:
    """Hash that ties a committed ncu capture to what it measured: the kernel sources plus `extra`."""
    import hashlib

    h = hashlib.sha256()
    for rp in rel_paths:
        h.update((ROOT / rp).read_bytes())
    h.update(extra.encode())
    return h.hexdigest()[:16]


# The DRAM capture of the igemm launches belongs to (a) the kernel's sources and (b) the launches the lowering asked for:
# every igemm op name carries its tile shape, tile count, K slices and stage kind, so the list of names of the plan is the
# fingerprint of (b) -- an edit of engine.py / lowering.py that does not change a single convolution keeps the capture valid.
IGEMM_SOURCES = ("tqdne_b200/csrc/tq_igemm_sm100.cu", "tqdne_b200/csrc/tq_ptx.cuh")


def igemm_fingerprint(op_names) -> str:
    return source_sha16(*IGEMM_SOURCES, extra="\n".join(n for n in op_names if n.startswith("igemm_sm100")))


# ------------------------------------------------------------------------------------------------------
# the other BASELINE.json configs (secondary measurements inside the same driver-run line)
# ------------------------------------------------------------------------------------------------------
FLOP_CFG0 = 35 * 28.436e9            # 1D EDM UNet, 18 Heun steps = 35 denoiser calls at L = 4064
FLOP_CFG3 = 63 * 271.89e9            # pixel-space 2D UNet, 32 Heun steps = 63 denoiser calls
FLOP_CFG4 = 3 * 28.436e9             # training step, forward + backward, per sample


def _event_time(fn, steps, warmup, torch, barrier=None):
    for _ in range(warmup):
        out = fn()
    (barrier or torch.cuda.synchronize)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    (barrier or torch.cuda.synchronize)()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps, out


def build_1d_edm(steps=18, seed=0):
    import tqdne_b200 as tq
    from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of
    from tqdne_b200.config import MovingAverageEnvelopeConfig

    cfg = MovingAverageEnvelopeConfig()
    ucfg = tq.get_1d_unet_config(cfg, 6, 6)
    edm = tq.LightningEDM(ucfg, {"learning_rate": 1e-4, "max_steps": 100000, "eta_min": 0.0}, num_sampling_steps=steps)
    sd = seeded_state_dict(shapes_of(edm), seed)
    for k in ("unet.out.2.weight", "unet.out.2.bias"):
        sd[k] = sd[k] * 0.05   # keep the random-init output inside the representation's range
    edm.load_state_dict(sd)
    return edm, sd, ucfg, cfg


def measure_cfg0(dev, pk, cpu: bool):
    """configs[0]: 1D EDM UNet (train_1d_edm), batch 4, 18 Heun steps, [4, 6, L] -> envelope inverse -> [4, 3, L];
    L = 4064 (the data set's length) and 4096 (the length BASELINE.json quotes).  With `cpu`, the reference's own CPU
    path on the same config is timed next to it (primary CPU config)."""
    import torch

    edm, sd, ucfg, cfg = build_1d_edm()
    edm.eval().to(dev).set_engine_precision("bf16")
    out = {"workload": "1D EDM UNet (train_1d_edm config), batch 4, 18 Heun steps (35 NFE) + moving-average-envelope inverse",
           "dtype": "bf16", "flop_per_waveform": FLOP_CFG0}
    cond = torch.from_numpy(cond_grid(4)).to(dev)
    for L in (4064, 4096):
        noise = torch.randn((4, 6, L), device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(1))
        noise_h, cond_h = noise.cpu().pin_memory(), cond.cpu().pin_memory()

        def resident():
            return cfg.representation.invert_representation_device(edm.sample((4, 6, L), cond=cond, noise=noise))

        def e2e():
            w = cfg.representation.invert_representation_device(
                edm.sample((4, 6, L), cond=cond_h.to(dev, non_blocking=True), noise=noise_h.to(dev, non_blocking=True)))
            return w.cpu()

        ms, _, w = _event_time(resident, 10, 3, torch)
        assert tuple(w.shape) == (4, 3, L) and bool(torch.isfinite(w).all())
        _, wall, _ = _event_time(e2e, 10, 2, torch)
        fl = FLOP_CFG0 * L / 4064
        out[f"L{L}"] = {"value": 4 / ms * 1e3, "unit": "waveforms/s", "ms_per_step": ms, "e2e_value": 4 / wall * 1e3,
                        "achieved_tflops": 4 / ms * 1e3 * fl / 1e12, "frac_of_bf16_burst_peak": 4 / ms * 1e3 * fl / 1e12 / pk["bf16_tflops"]}
    # the same pipeline at a batch that fills the GPU (the batch-4 case is launch / latency bound by construction)
    B = 64
    noise = torch.randn((B, 6, 4064), device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(2))
    cond = torch.from_numpy(cond_grid(B)).to(dev)
    ms, _, w = _event_time(lambda: cfg.representation.invert_representation_device(edm.sample((B, 6, 4064), cond=cond, noise=noise)),
                           3, 2, torch)
    out["L4064_batch64"] = {"value": B / ms * 1e3, "unit": "waveforms/s", "ms_per_step": ms,
                            "achieved_tflops": B / ms * 1e3 * FLOP_CFG0 / 1e12,
                            "frac_of_bf16_burst_peak": B / ms * 1e3 * FLOP_CFG0 / 1e12 / pk["bf16_tflops"]}
    if cpu:
        out["cpu_baseline"] = cpu_cfg0(sd, ucfg)
    return out


def cpu_cfg0(sd=None, ucfg=None, L=4064):
    """The reference's CPU generate path on configs[0]: LightningEDM.sample_deterministically (unmodified reference when
    /root/reference exists, else the oracle port) + MovingAverageEnvelope inverse, batch 4, 18 Heun steps, all host cores."""
    import torch

    from oracle import reference_loader, torch_ref

    if sd is None:
        _, sd, ucfg, _ = build_1d_edm()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(5)
    noise = torch.randn((4, 6, L), generator=gen, dtype=torch.float64)
    cond = torch.from_numpy(cond_grid(4))
    kind = "port"
    t0 = time.perf_counter()
    with torch.no_grad():
        if reference_loader.available():
            ref = reference_loader.load()
            mod = ref.edm.LightningEDM(ucfg, {}, num_sampling_steps=18)
            mod.load_state_dict(sd)
            mod.eval()
            sig = mod.edm.sampling_sigmas(18)
            t0 = time.perf_counter()
            x = mod.sample_deterministically(noise * sig[0], sig, None, cond)
            kind = "reference"
        else:
            sig = torch_ref.sampling_sigmas(18)
            x = torch_ref.heun_sample(sd, ucfg, noise * sig[0], sig, cond)
        w = torch_ref.mavg_envelope_inverse(x.float().numpy())
    dt = time.perf_counter() - t0
    assert w.shape == (4, 3, L)
    return {"value": 4 / dt, "unit": "waveforms/s", "cores": cores, "kind": kind,
            "sample": f"the whole config: 4 x [6, {L}], 18 Heun steps (35 NFE) + envelope inverse, {dt:.1f} s on the host"}


def measure_cfg3(dev, pk, batch=1024):
    """configs[3]: pixel-space 2D log-spectrogram EDM UNet (train_edm), batch 1024, 32 Heun steps (63 NFE) at
    [3, 128, 128] + Griffin-Lim; sample() micro-batches (max_positions_per_pass).  One timed pass (~20 s)."""
    import torch

    import tqdne_b200 as tq
    from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of
    from tqdne_b200.config import SpectrogramConfig

    cfg = SpectrogramConfig()
    edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, 3, 3), {}, num_sampling_steps=32)
    sd = seeded_state_dict(shapes_of(edm), 0)
    for k in ("unet.out.2.weight", "unet.out.2.bias"):
        sd[k] = sd[k] * 0.05
    edm.load_state_dict(sd)
    edm.eval().to(dev).set_engine_precision("bf16")
    micro = max(1, edm.max_positions_per_pass // (128 * 128))
    cond = torch.from_numpy(cond_grid(batch)).to(dev)
    noise = torch.randn((batch, 3, 128, 128), device=dev, dtype=torch.float64, generator=torch.Generator(device=dev).manual_seed(1))

    def run(n):
        rep = edm.sample((n, 3, 128, 128), cond=cond[:n], noise=noise[:n])
        return cfg.representation.invert_representation_device(torch.tanh(rep))

    run(2 * micro)                      # warm-up: builds and captures the micro-batch plan the full pass replays
    ms, _, w = _event_time(lambda: run(batch), 1, 0, torch)
    assert tuple(w.shape) == (batch, 3, cfg.t) and bool(torch.isfinite(w).all())
    tf = batch / ms * 1e3 * FLOP_CFG3 / 1e12
    return {"workload": "pixel-space 2D log-spectrogram EDM UNet (train_edm config), 32 Heun steps (63 NFE) + Griffin-Lim",
            "batch": batch, "micro_batch": micro, "dtype": "bf16", "value": batch / ms * 1e3, "unit": "waveforms/s",
            "ms_per_step": ms, "flop_per_waveform": FLOP_CFG3, "achieved_tflops": tf, "frac_of_bf16_burst_peak": tf / pk["bf16_tflops"],
            "frac_of_bf16_sustained_peak": tf / pk["bf16_tflops_sustained"]}


def measure_cfg4(dev, pk, world, rank, barrier, batch=64):
    """configs[4]: 1D EDM UNet bf16 training step (forward + backward, NCCL gradient all-reduce, Adam + EMA), batch 64 per
    GPU (512 at 8 GPUs), dropout active, through LightningEDM.training_step."""
    import torch
    import torch.distributed as dist

    edm, _, _, _ = build_1d_edm()
    edm.to(dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    batch_d = {"signal": torch.randn(batch, 6, 4064, device=dev, generator=g), "cond": torch.randn(batch, 5, device=dev, generator=g)}
    ms, wall, loss = _event_time(lambda: edm.training_step(batch_d), 10, 3, torch, barrier)
    t = torch.tensor([max(ms, wall)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    assert bool(torch.isfinite(loss)), "non-finite training loss"
    sps = batch * world / ms * 1e3
    tf = sps / world * FLOP_CFG4 / 1e12
    edm.__dict__.pop("_tq_train", None)
    return {"workload": "1D EDM UNet bf16 training step (fwd + bwd + gradient all-reduce + Adam + EMA), [64, 6, 4064] per GPU",
            "batch_per_gpu": batch, "global_batch": batch * world, "dtype": "bf16", "value": sps, "unit": "samples/s",
            "ms_per_step": ms, "flop_per_sample": FLOP_CFG4, "achieved_tflops_per_gpu": tf,
            "frac_of_bf16_burst_peak": tf / pk["bf16_tflops"], "allreduce": "NCCL, flat fp32 gradient buffer" if world > 1 else None}


def measure_cfg2(args, dev, pk, world, rank, barrier, total=8192):
    """configs[2]: the latent pipeline at a GLOBAL batch of 8192 conditioned on the magnitude / distance / vs30 grid,
    split over the N ranks (strong scaling of a fixed job); the final gather of 8192 x [3, 4064] float32 waveforms
    (400 MB) to rank 0 and its read-back to the host are inside the end-to-end time."""
    import torch
    import torch.distributed as dist

    import tqdne_b200 as tq
    from tqdne_b200 import sharding
    from tqdne_b200.config import LatentSpectrogramConfig

    cfg = LatentSpectrogramConfig()
    enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
    edm = tq.LightningEDM(tq.get_2d_unet_config(cfg, cfg.latent_channels, cfg.latent_channels), {}, num_sampling_steps=NFE_STEPS,
                          autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
    edm.load_state_dict(build_state_dict(edm))
    edm.eval().to(dev).set_engine_precision(args.precision)
    lo, hi = sharding.shard_bounds(total, rank, world)
    cond = torch.from_numpy(cond_grid(total))[lo:hi].contiguous().pin_memory()
    noise = sharding.global_noise((8, 32, 32), lo, hi, seed=1, device="cpu").pin_memory()

    def step():
        rep = edm.sample((hi - lo, 3, 128, 128), cond=cond.to(dev, non_blocking=True), noise=noise.to(dev, non_blocking=True))
        wav = cfg.representation.invert_representation_device(rep).to(torch.float32)
        full = sharding.gather_waveforms(wav, total)
        return sharding.to_host(full) if full is not None else None

    ms, wall, out = _event_time(step, 2, 1, torch, barrier)
    t = torch.tensor([max(ms, wall)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if rank == 0:
        assert tuple(out.shape) == (total, 3, cfg.t)
    val = total / ms * 1e3
    return {"workload": "HighFEM latent EDM, global batch 8192 on the magnitude / distance / vs30 grid, split over the ranks; "
                        "gather + D2H of all waveforms inside the time", "global_batch": total, "batch_per_gpu": hi - lo,
            "scaling": "strong", "value": val, "unit": "waveforms/s", "ms_per_step": ms, "gather_bytes": total * 3 * cfg.t * 4,
            "achieved_tflops_per_gpu": val / world * FLOP_PER_WAVEFORM / 1e12,
            "frac_of_bf16_burst_peak": val / world * FLOP_PER_WAVEFORM / 1e12 / pk["bf16_tflops"]}


def run_engine(args) -> None:
    import torch
    import torch.distributed as dist

    import tqdne_b200 as tq
    from tqdne_b200 import _lib, sharding
    from tqdne_b200.config import LatentSpectrogramConfig

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the first communicator is created: keep stdout for the ONE JSON
        # line by pointing fd 1 at stderr while the process group (eager: device_id given) comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    B = args.batch
    total = B * world
    lo, hi = sharding.shard_bounds(total, rank, world)

    cfg = LatentSpectrogramConfig()
    enc_cfg, dec_cfg = tq.get_2d_autoencoder_configs(cfg)
    unet_cfg = tq.get_2d_unet_config(cfg, cfg.latent_channels, cfg.latent_channels)
    edm = tq.LightningEDM(unet_cfg, {}, num_sampling_steps=NFE_STEPS,
                          autoencoder=tq.LightningAutoencoder(enc_cfg, dec_cfg, {}))
    edm.load_state_dict(build_state_dict(edm))
    edm.eval().to(dev).set_engine_precision(args.precision)
    rep_inv = cfg.representation

    cond_all = torch.from_numpy(cond_grid(total))
    cond_host = cond_all[lo:hi].contiguous().pin_memory()
    noise_host = sharding.global_noise((8, 32, 32), lo, hi, seed=1, device="cpu").pin_memory()
    cond_dev = cond_host.to(dev)
    noise_dev = noise_host.to(dev)
    h2d = cond_host.numel() * 4 + noise_host.numel() * 8
    d2h = total * 3 * cfg.t * 4   # rank 0 reads the gathered waveforms of ALL ranks back

    def step_resident():
        rep = edm.sample((hi - lo, 3, 128, 128), cond=cond_dev, noise=noise_dev)
        return rep_inv.invert_representation_device(rep)

    def step_e2e():
        c = cond_host.to(dev, non_blocking=True)
        z = noise_host.to(dev, non_blocking=True)
        rep = edm.sample((hi - lo, 3, 128, 128), cond=c, noise=z)
        # fp64 Griffin-Lim (parity mode); the waveforms are stored as float32 like the reference's output dataset
        wav = rep_inv.invert_representation_device(rep).to(torch.float32)
        if world > 1:
            full = sharding.gather_waveforms(wav, total)   # the one collective: final gather to rank 0
            return sharding.to_host(full) if full is not None else None
        return sharding.to_host(wav)       # D2H into a pinned buffer, synchronised (what generate() does before writing)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]), out

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(3, args.warmup)):
        w = step_resident()
    assert tuple(w.shape) == (hi - lo, 3, cfg.t) and bool(torch.isfinite(w).all()), "non-finite waveforms"
    _lib.launch_count_reset()
    t_begin = time.time()
    ev_ms, wall_ms, _ = timed(step_resident, args.steps)
    t_end = time.time()
    launches = _lib.launch_count()
    clk = clocks.stop(t_begin, t_end) if rank == 0 else None
    step_e2e()  # warm the pinned-copy path
    e2e_ev_ms, e2e_wall_ms, _ = timed(step_e2e, args.steps)
    e2e_ms = max(e2e_ev_ms, e2e_wall_ms)  # the D2H at the end is host-synchronous: wall clock bounds it

    value = total * args.steps / (ev_ms / 1e3)
    e2e_value = total * args.steps / (e2e_ms / 1e3)
    pk = peaks()
    line = {
        "metric": METRIC, "value": value, "unit": "waveforms/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "HighFEM latent EDM: latent UNet Heun sampling + autoencoder decode + log-spectrogram "
                               "inverse, batch 256 per B200 (BASELINE.json configs[1])",
                   "batch_per_gpu": B, "global_batch": total, "heun_steps": NFE_STEPS, "nfe": 2 * NFE_STEPS - 1,
                   "griffinlim_iters": 128, "griffinlim_precision": "fp64 (the reference's locked NumPy 2 runs complex128)",
                   "weights": "random-init (seeded)", "parallelism": f"batch-sharded x{world}",
                   "cache": "activations of one denoiser call (~1.5 GB) exceed the 126 MB L2; no L2 flush needed",
                   "cuda_graph": bool(edm.use_cuda_graph)},
        "e2e": {"value": e2e_value, "unit": "waveforms/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clk,
        "achieved_tflops_per_gpu": value / world * FLOP_PER_WAVEFORM / 1e12,
        "frac_of_bf16_sustained_peak": value / world * FLOP_PER_WAVEFORM / 1e12 / pk["bf16_tflops_sustained"],
        "wall_ms_per_step": wall_ms / args.steps,
    }
    if rank == 0 and world == 1:
        agg, graph_ms, op_sum_ms, plan_op_names = kernel_breakdown(edm, hi - lo)
        tot_ms = graph_ms
        conv = agg.get("igemm_sm100") or agg.get("igemm_simt")
        ach = conv["flops"] / (conv["ms"] / 1e3) / 1e12
        line["roofline"] = {
            "kernel": "igemm_sm100 (tcgen05 implicit-GEMM conv), all launches of one denoiser call",
            "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
            "peak_source": f"{pk['source']} bf16 burst (MEASURED_PEAKS.json)", "traffic": None,
            "launches_per_call": conv["launches"], "share_of_denoiser_call": conv["ms"] / tot_ms,
            "frac_of_sustained_peak": ach / pk["bf16_tflops_sustained"],
            "timing": "op-by-op CUDA-event time minus the per-launch event gap, so that the kernel times add up to the captured "
                      "graph the step replays",
            "graph_ms_per_denoiser_call": graph_ms, "op_by_op_ms_per_denoiser_call": op_sum_ms,
        }
        # DRAM bytes per launch from the committed ncu capture of the same launches (tools/summarize_dram.py); quoted only
        # while the capture belongs to the kernel sources of this checkout (hash), never a stale number
        tfile = ROOT / "profiles" / "igemm_dram_traffic.json"
        if tfile.exists():
            tj = json.loads(tfile.read_text())
            if tj.get("launches") == conv["launches"] and tj.get("source_sha16") == igemm_fingerprint(plan_op_names):
                line["roofline"]["traffic"] = tj["bytes_per_launch"]
                line["roofline"]["traffic_unit"] = "DRAM bytes per launch (read + write), ncu --set full, " + tj.get("capture", "")
            else:
                line["roofline"]["traffic_note"] = "committed ncu capture belongs to other kernel sources: not quoted"
            line["roofline"]["algorithmic_bytes_per_launch"] = conv.get("bytes", 0) / conv["launches"] or None
        gn = {k: agg[k] for k in agg if k.startswith("gn_")}
        if gn:
            gb = sum(a["bytes"] for a in gn.values())
            gms = sum(a["ms"] for a in gn.values())
            line["roofline_groupnorm"] = {"bound": "hbm", "achieved": gb / (gms / 1e3) / 1e9, "peak": pk["hbm_gbs"],
                                          "unit": "GB/s", "frac": gb / (gms / 1e3) / 1e9 / pk["hbm_gbs"],
                                          "share_of_denoiser_call": gms / tot_ms}
        line["kernel_ms_per_denoiser_call"] = {k: round(a["ms"], 4) for k, a in sorted(agg.items())}
        # the step outside the denoiser calls: decode and the fp64 Griffin-Lim launch
        rep = edm.sample((hi - lo, 3, 128, 128), cond=cond_dev, noise=noise_dev)
        gl_ms, _, _ = _event_time(lambda: rep_inv.invert_representation_device(rep), 3, 1, torch)
        lat = torch.randn((hi - lo, 8, 32, 32), device=dev)
        dec_ms, _, _ = _event_time(lambda: edm.autoencoder.decode(lat), 3, 1, torch)
        line["step_parts_ms"] = {"denoiser_calls_49x": 49 * graph_ms, "decode": dec_ms, "griffinlim_fp64_768_items": gl_ms}
        if not args.no_cpu_baseline:
            pipe = CpuPipeline(args.cpu_batch)
            pipe.step(100)     # warm-up (thread pools, first-touch)
            t0 = time.perf_counter()
            pipe.step(0)
            dt = time.perf_counter() - t0
            pipe.close()
            line["cpu_baseline"] = {"value": args.cpu_batch / dt, "unit": "waveforms/s", "cores": pipe.cores,
                                    "kind": pipe.kind,
                                    "sample": f"{args.cpu_batch} waveforms, full 25-step Heun + decode + 128-iter "
                                              f"Griffin-Lim, {dt:.1f} s on the host (after one warm-up pass)"}
    # ---- the other BASELINE.json configs, in the same line
    which = set(args.configs.split(",")) if args.configs not in ("all", "none") else (
        {"cfg0", "cfg2", "cfg3", "cfg4"} if args.configs == "all" else set())
    configs = {}
    del edm
    torch.cuda.empty_cache()
    if world == 1 and rank == 0:
        if "cfg0" in which:
            configs["cfg0"] = measure_cfg0(dev, pk, cpu=not args.no_cpu_baseline)
        if "cfg3" in which:
            configs["cfg3"] = measure_cfg3(dev, pk, args.cfg3_batch)
    if "cfg4" in which:
        c4 = measure_cfg4(dev, pk, world, rank, barrier)
        if rank == 0:
            configs["cfg4"] = c4
    if "cfg2" in which and world > 1:
        c2 = measure_cfg2(args, dev, pk, world, rank, barrier)
        if rank == 0:
            configs["cfg2"] = c2
    if configs:
        line["configs"] = configs
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="waveforms per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="waveforms per step of the CPU arm / CPU baseline sample "
                                                              "(8 = ~5 s per pass on 16 cores, one warm-up pass + one timed pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="all",
                    help="secondary BASELINE.json configs measured after the headline: all | none | comma list of cfg0,cfg2,cfg3,cfg4")
    ap.add_argument("--cfg3-batch", type=int, default=1024, help="batch of the pixel-space config (configs[3] names 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()

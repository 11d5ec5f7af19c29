"""EDM preconditioning + Heun sampler on the B200 engine.

Drop-in for tqdne/edm.py: `EDM` constants/scalars (edm.py:9-52), `LightningEDM(unet_config, optimizer_params,
num_sampling_steps, deterministic_sampling, edm, autoencoder)` with `.forward`, `.sample`, `.evaluate`
(edm.py:55-238).  The sampler keeps the reference numerics: fp64 state, fp32 sigma schedule, fp32 model I/O,
(sigma_next - sigma) rounded in fp32, final step Euler-only.  Each NFE is: one replay of the UNet kernel
plan + ONE fused element-wise kernel (denoiser post-scaling, Euler/Heun update, next input scaling).
"""

from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib
from .autoencoder import LightningAutoencoder
from .engine import current_stream_ptr, device_guard, nchw_to_nhwc, nhwc_to_nchw, require_cuda, tq_dtype
from .lightning_shim import LightningModule
from .lowering import get_coder_plan, get_unet_plan
from .nn import append_dims
from .unet import UNetModel


class EDM:
    """Hyper-parameters and scalar maps of Karras et al. (reference: tqdne/edm.py:9-52)."""

    sigma_min: float = 0.002
    sigma_max: float = 80.0
    rho: float = 7.0
    sigma_data: float = 0.5
    P_mean: float = -1.2
    P_std: float = 1.2
    S_churn: float = 40
    S_min: float = 0.05
    S_max: float = 50
    S_noise: float = 1.003

    def sigma(self, eps):
        return (eps * self.P_std + self.P_mean).exp()

    def loss_weight(self, sigma):
        return (sigma**2 + self.sigma_data**2) / (sigma * self.sigma_data) ** 2

    def skip_scaling(self, sigma):
        return self.sigma_data**2 / (sigma**2 + self.sigma_data**2)

    def out_scaling(self, sigma):
        return sigma * self.sigma_data / (sigma**2 + self.sigma_data**2) ** 0.5

    def in_scaling(self, sigma):
        return 1 / (sigma**2 + self.sigma_data**2) ** 0.5

    def noise_conditioning(self, sigma):
        return 0.25 * sigma.log()

    def sampling_sigmas(self, num_steps, device=None):
        inv_rho = 1 / self.rho
        idx = torch.arange(num_steps, dtype=torch.float32, device=device)
        lo, hi = self.sigma_min**inv_rho, self.sigma_max**inv_rho
        sigmas = (hi + idx / (num_steps - 1) * (lo - hi)) ** self.rho
        return torch.cat([sigmas, torch.zeros_like(sigmas[:1])])

    def sigma_hat(self, sigma, num_steps):
        gamma = min(self.S_churn / num_steps, 2**0.5 - 1) if self.S_min <= sigma <= self.S_max else 0
        return sigma + gamma * sigma


def _f(x) -> float:
    return float(x)


class _Coeffs:
    """fp32 scalars of one noise level, rounded exactly like the reference's fp32 tensor arithmetic (CPU)."""

    def __init__(self, edm: EDM, sigma: torch.Tensor):
        s = sigma.detach().to("cpu", torch.float32).reshape(())
        self.sigma = _f(s)
        self.c_in = _f(edm.in_scaling(s))
        self.c_out = _f(edm.out_scaling(s))
        self.c_skip = _f(edm.skip_scaling(s))
        self.c_noise = _f(edm.noise_conditioning(s)) if self.sigma > 0 else 0.0


class _SideStream:
    """Run on a non-legacy stream ordered after the caller's stream; order the caller after it on exit."""

    def __init__(self, owner):
        self.owner, self.ctx = owner, None

    def __enter__(self):
        dev = torch.cuda.current_device()   # the module's device: callers enter under device_guard
        self.cur = torch.cuda.current_stream(dev)
        if self.cur.cuda_stream != 0:
            return self
        streams = self.owner.__dict__.setdefault("_tq_streams", {})
        s = streams.get(dev)
        if s is None:
            s = torch.cuda.Stream(device=dev)
            streams[dev] = s
        s.wait_stream(self.cur)
        self.s = s
        self.ctx = torch.cuda.stream(s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            self.cur.wait_stream(self.s)
        return False


class LightningEDM(LightningModule):
    """reference: tqdne/edm.py:55-251 (training step / optimizers are out of this engine's scope)."""

    def __init__(self, unet_config: dict, optimizer_params: dict, num_sampling_steps: int = 25,
                 deterministic_sampling: bool = True, edm: EDM = EDM(), autoencoder: None | LightningAutoencoder = None):
        super().__init__()
        self.unet = UNetModel(**unet_config)
        self.optimizer_params = optimizer_params
        self.num_sampling_steps = num_sampling_steps
        self.deterministic_sampling = deterministic_sampling
        self.edm = edm
        self.autoencoder = autoencoder.eval() if autoencoder else None
        self.config = unet_config
        if self.autoencoder:
            for p in self.autoencoder.parameters():
                p.requires_grad = False
        self.save_hyperparameters(ignore=("autoencoder"))
        # engine knobs (not part of the reference API)
        # micro-batch = this / (H*W) samples per Heun pass: 1024 latent (32 x 32), 64 pixel-space (128 x 128: 843 -> 953 TFLOP/s
        # against micro-batches of 16, tools/pixel_microbatch.py), 258 1-D (L = 4064) samples
        self.max_positions_per_pass = 1024 * 1024
        self.decode_micro_batch = 64
        self.use_cuda_graph = True
        self.compat_rng = True    # reproduce the reference's RNG draw order inside sample()

    # ---- precision ---------------------------------------------------------------------------------
    def set_engine_precision(self, precision: str | None) -> "LightningEDM":
        """'bf16' -> tcgen05 tensor path, 'fp32' -> FFMA parity path, None -> follow the parameter dtype."""
        dt = {None: None, "bf16": torch.bfloat16, "fp32": torch.float32}[precision]
        self.unet.engine_dtype = dt
        if self.autoencoder:
            self.autoencoder.encoder.engine_dtype = dt
            self.autoencoder.decoder.engine_dtype = dt
        return self

    # ---- denoiser ----------------------------------------------------------------------------------
    def forward(self, sample, sigma, cond_sample=None, cond=None):
        """D(x, sigma) = c_out * F(c_in * x, c_noise, cond) + c_skip * x  (reference: edm.py:105-113)."""
        dim = sample.dim()
        sample_in = sample * append_dims(self.edm.in_scaling(sigma), dim)
        inp = sample_in if cond_sample is None else torch.cat((sample_in, cond_sample), dim=1)
        out = self.unet(inp, self.edm.noise_conditioning(sigma), cond=cond)   # device-guarded in UNetModel.forward
        skip = append_dims(self.edm.skip_scaling(sigma), dim) * sample
        return out * append_dims(self.edm.out_scaling(sigma), dim) + skip

    # ---- sampling ------------------------------------------------------------------------------------
    def latent_shape(self, shape) -> tuple:
        """Shape of autoencoder.encode(zeros(shape)) by arithmetic (the reference runs a full dummy encode,
        edm.py:155-157)."""
        enc = self.autoencoder.encoder
        n_down = sum(1 for m in enc.down_blocks if type(m).__name__ == "Downsample")
        ch = enc.output_layer.weight.shape[0] // 2
        return (shape[0], ch, *[s // (2**n_down) for s in shape[2:]])

    @torch.no_grad()
    def sample(self, shape, cond_sample=None, cond=None, noise=None, generator=None):
        """Heun 2nd-order sampler (reference: edm.py:146-169).

        `noise` (optional, engine extension): explicit unit-variance fp64 noise of the (latent) shape; when
        omitted, noise is drawn with the reference's draw order so that equal seeds give equal noise.
        """
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("tqdne_b200: LightningEDM.sample needs the module on a CUDA device (no CPU path)")
        with device_guard(dev):
            return self._sample(tuple(shape), cond, noise, generator, dev, cond_sample)

    def _sample(self, shape, cond, noise, generator, dev, cond_sample=None):
        if cond_sample is not None:
            # signal-conditioned sampling: cond_sample rides along as extra UNet input channels, fixed over the calls
            require_cuda(cond_sample, "cond_sample")
            if self.autoencoder:   # reference: encoded BEFORE the dummy encode (edm.py:151-152), one randn_like draw
                mean, log_std = torch.chunk(self.autoencoder.encoder(cond_sample), 2, dim=1)
                cond_sample = mean + torch.randn(mean.shape, device=dev, dtype=mean.dtype, generator=generator) * torch.exp(log_std)
        if self.autoencoder:
            shape = self.latent_shape(shape)
            if noise is None and self.compat_rng:
                # reference: encode(zeros) draws randn_like(mean) before the sampler noise (autoencoder.py:39)
                torch.randn(shape, device=dev, dtype=torch.float32, generator=generator)
        sigmas = self.edm.sampling_sigmas(self.num_sampling_steps, device="cpu")
        if noise is None:
            noise = torch.randn(shape, device=dev, dtype=torch.float64, generator=generator)
        else:
            assert tuple(noise.shape) == shape, f"noise must have shape {shape}"
            noise = noise.to(dev, torch.float64)
        eps = noise * sigmas[0].to(dev)  # fp64 * 0-dim fp32 -> fp64, like the reference

        N = shape[0]
        spatial = shape[2:]
        P = math.prod(spatial)
        micro = max(1, min(N, self.max_positions_per_pass // P))
        churn = None
        if not self.deterministic_sampling:
            # stochastic sampler: the per-step churn noise (reference: th.randn_like(sample_curr), edm.py:204) is drawn for
            # the WHOLE batch, step by step, from the caller's generator -- reproducible for a given generator and
            # independent of the micro-batching below
            churn = [torch.randn(shape, device=dev, dtype=torch.float64, generator=generator)
                     for _ in range(self.num_sampling_steps)]
        with self._engine_stream():
            outs = []
            for i0 in range(0, N, micro):
                i1 = min(N, i0 + micro)
                c = cond[i0:i1] if cond is not None else None
                nz = [z[i0:i1] for z in churn] if churn is not None else None
                cs = cond_sample[i0:i1] if cond_sample is not None else None
                outs.append(self._sample_chunk(eps[i0:i1].contiguous(), sigmas, c, noises=nz, cond_sample=cs))
            x = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)   # fp64 channels-last [N, P, C]
            C_ = shape[1]
            if not self.autoencoder:
                result = nhwc_to_nchw(x, N, C_, spatial, C_, torch.float32)
            else:
                result = self._decode_latents(x, N, C_, spatial)
        result.record_stream(torch.cuda.current_stream(dev))
        return result

    def _engine_stream(self):
        """CUDA graphs cannot be captured on the legacy default stream: run the sampler on a side stream that
        is ordered after / before the caller's current stream."""
        return _SideStream(self)

    def _sample_chunk(self, eps: torch.Tensor, sigmas: torch.Tensor, cond, noises=None, cond_sample=None) -> torch.Tensor:
        """Deterministic / stochastic Heun loop on one micro-batch; returns the fp64 channels-last state.
        `noises`: per-step unit normal fp64 tensors of eps.shape for the stochastic sampler (drawn here when None)."""
        lib = _lib.lib()
        n, C_ = eps.shape[0], eps.shape[1]
        spatial = tuple(eps.shape[2:])
        NP = n * math.prod(spatial)
        cc = 0
        if cond_sample is not None:
            cc = cond_sample.shape[1]
            assert C_ + cc == self.unet.in_channels and tuple(cond_sample.shape[2:]) == spatial, \
                "cond_sample must supply the UNet input channels beyond the sampled state, at the state's size"
        plan = get_unet_plan(self.unet, n, spatial, uniform_t=True, cond_channels=cc)
        if self.use_cuda_graph:
            plan.plan.enable_graph(True)
        if cond_sample is not None:
            plan.set_cond_sample(cond_sample)
        if cond is not None:
            require_cuda(cond, "cond")
        assert (cond is not None) == (self.unet.cond_features is not None), \
            "must specify cond if and only if the model is conditioned"
        plan.set_cond(cond)
        dev = eps.device
        x = nchw_to_nhwc(eps, torch.float64)              # [n, P, C] fp64
        x1 = torch.empty_like(x)
        d = torch.empty_like(x)
        xin, dt_ = plan.xin.t, tq_dtype(plan.act_dtype)
        F, Cf, Cpad = plan.out.t, plan.out.C, plan.cin_pad
        nsteps = self.num_sampling_steps
        stochastic = not self.deterministic_sampling
        # the scalar schedule (fp32 coefficients, rounded like the reference's 0-dim tensor arithmetic) depends only on
        # the sigma ladder: computed once per ladder and kept with its device copy of the c_noise values
        key = (tuple(float(v) for v in sigmas), stochastic, str(dev))
        sched = self.__dict__.setdefault("_tq_sched", {}).get(key)
        if sched is None:
            co = [_Coeffs(self.edm, s) for s in sigmas]
            hats, co_hat = None, None
            if stochastic:
                hats = [self.edm.sigma_hat(s, nsteps) for s in sigmas[:-1]]
                co_hat = [_Coeffs(self.edm, s) for s in hats]
            tvals = []
            for i in range(nsteps):
                tvals.append(co_hat[i].c_noise if stochastic else co[i].c_noise)
                tvals.append(co[i + 1].c_noise)
            if stochastic:
                dts = [float((sigmas[i + 1] - hats[i]).to(torch.float32)) for i in range(nsteps)]
            else:
                dts = [float(sigmas[i + 1] - sigmas[i]) for i in range(nsteps)]  # fp32 subtraction, like the reference
            sched = (co, hats, co_hat, tvals, dts)
            self.__dict__["_tq_sched"][key] = sched
        co, hats, co_hat, tvals, dts = sched
        st = current_stream_ptr()
        # The time input of denoiser call k (c_noise of its noise level, tvals[k]) is written by the sampler kernel that runs
        # right before that call anyway (tq_edm_* `t_next`): a separate 4-byte copy between two graph replays cost 0.22 ms
        # per call (4.41 -> 4.63 ms, tools/sampler_breakdown.py) -- 5 % of the whole sampler.
        tp = plan.t.data_ptr()

        if not stochastic:
            _lib.check(lib.tq_edm_precondition(x.data_ptr(), xin.data_ptr(), dt_, NP, C_, Cpad, co[0].c_in, tp, tvals[0], st),
                       "precondition")
        for i in range(nsteps):
            cur, nxt = (co_hat[i] if stochastic else co[i]), co[i + 1]
            if stochastic:
                # x_hat = x + S_noise * randn * sqrt(sigma_hat^2 - sigma^2)   (reference: edm.py:203-207)
                s_hat, s_cur = hats[i], sigmas[i]
                scale = float((s_hat**2 - s_cur**2) ** 0.5) * self.edm.S_noise
                if noises is not None:
                    assert tuple(noises[i].shape) == tuple(eps.shape), "per-step noise must have the shape of eps"
                    nz = nchw_to_nhwc(noises[i].to(dev, torch.float64), torch.float64)
                else:
                    nz = nchw_to_nhwc(torch.randn(eps.shape, device=dev, dtype=torch.float64), torch.float64)
                _lib.check(lib.tq_edm_add_noise(x.data_ptr(), nz.data_ptr(), scale, x.numel(), st), "add_noise")
                _lib.check(lib.tq_edm_precondition(x.data_ptr(), xin.data_ptr(), dt_, NP, C_, Cpad, cur.c_in, tp, tvals[2 * i], st),
                           "precondition")
            dt = dts[i]
            last = i == nsteps - 1
            plan.run()                                           # denoiser call 2i
            _lib.check(lib.tq_edm_euler(x.data_ptr(), F.data_ptr(), Cf, d.data_ptr(), x1.data_ptr(), xin.data_ptr(), dt_,
                                        NP, C_, Cpad, cur.c_out, cur.c_skip, cur.sigma, dt, nxt.c_in, 0 if last else 1,
                                        None if last else tp, 0.0 if last else tvals[2 * i + 1], st), "euler")
            if last:
                x, x1 = x1, x
                break
            plan.run()                                           # denoiser call 2i + 1
            nn_cin = co[i + 1].c_in  # the next step starts at sigma_{i+1}
            _lib.check(lib.tq_edm_heun(x.data_ptr(), x1.data_ptr(), d.data_ptr(), F.data_ptr(), Cf, xin.data_ptr(), dt_, NP,
                                       C_, Cpad, nxt.c_out, nxt.c_skip, nxt.sigma, dt, nn_cin, 0 if stochastic else 1,
                                       None if stochastic else tp, 0.0 if stochastic else tvals[2 * i + 2], st), "heun")
        return x

    def _decode_latents(self, x: torch.Tensor, N: int, C_: int, spatial: tuple) -> torch.Tensor:
        """sample.to(fp32) -> autoencoder.decode (reference: edm.py:166-168), micro-batched."""
        lib = _lib.lib()
        dec = self.autoencoder.decoder
        P = math.prod(spatial)
        outs = []
        mb = min(N, self.decode_micro_batch)
        st = current_stream_ptr()
        for i0 in range(0, N, mb):
            n = min(mb, N - i0)
            p = get_coder_plan(dec, "decoder", n, spatial)
            xs = x[i0:i0 + n]
            _lib.check(lib.tq_edm_precondition(xs.data_ptr(), p.xin.t.data_ptr(), tq_dtype(p.act_dtype), n * P, C_,
                                               p.cin_pad, 1.0, None, 0.0, st), "decode input")
            p.run()
            so = (p.out.H, p.out.W) if len(spatial) == 2 else (p.out.W,)
            outs.append(nhwc_to_nchw(p.out.t, n, p.cout, so, p.cout, torch.float32))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    # reference API: the two samplers are also callable on their own with explicit eps (edm.py:171-230)
    def _run_sampler(self, eps, sigmas, cond, deterministic: bool, noises=None, cond_sample=None):
        keep = self.deterministic_sampling
        self.deterministic_sampling = deterministic
        try:
            with device_guard(eps.device), self._engine_stream():
                x = self._sample_chunk(eps.to(torch.float64).contiguous(), sigmas.to("cpu"), cond, noises=noises,
                                       cond_sample=cond_sample)
                out = nhwc_to_nchw(x, eps.shape[0], eps.shape[1], tuple(eps.shape[2:]), eps.shape[1], torch.float64)
        finally:
            self.deterministic_sampling = keep
        out.record_stream(torch.cuda.current_stream(eps.device))
        return out

    @torch.no_grad()
    def sample_deterministically(self, eps, sigmas, cond_sample=None, cond=None):
        return self._run_sampler(eps, sigmas, cond, True, cond_sample=cond_sample)

    @torch.no_grad()
    def sample_stochastically(self, eps, sigmas, cond_sample=None, cond=None, noises=None, generator=None):
        """reference: edm.py:198-230.  Engine extensions: `noises` = the per-step unit normal draws (one fp64 tensor of
        eps.shape per Heun step, what the reference takes from th.randn_like), else drawn from `generator`."""
        if noises is None:
            noises = [torch.randn(eps.shape, device=eps.device, dtype=torch.float64, generator=generator)
                      for _ in range(self.num_sampling_steps)]
        assert len(noises) == self.num_sampling_steps, "one noise tensor per Heun step"
        return self._run_sampler(eps, sigmas, cond, False, noises=noises, cond_sample=cond_sample)

    @torch.no_grad()
    def evaluate(self, batch):
        """reference: edm.py:232-238."""
        cond_sample = batch["cond_signal"] if "cond_signal" in batch else None
        cond = batch["cond"] if "cond" in batch else None
        return self.sample(batch["signal"].shape, cond_sample, cond)

    # ---- training: the 1D UNet (BASELINE.json configs[4]); everything else fails loudly
    def _train_step(self, batch):
        from .training import TrainStep1D

        if self.autoencoder is not None or self.unet.dims != 1 or "cond_signal" in batch:
            raise NotImplementedError("tqdne_b200: the training step is built for the 1D EDM UNet without an autoencoder / "
                                      "cond_signal (train_1d_edm config); train the other models with the reference package")
        x = batch["signal"]
        key = (x.shape[0], x.shape[-1])
        tapes = self.__dict__.setdefault("_tq_train", {})
        ts = tapes.get(key)
        if ts is None:
            # a new (batch, length): a new static tape over the SAME parameter / Adam / EMA state and schedule position
            # (a ragged last batch or an evaluation batch must not restart training)
            op = self.optimizer_params or {}
            shared = next(iter(tapes.values())).store if tapes else None
            ts = TrainStep1D(self, x.shape[0], x.shape[-1], lr=op.get("learning_rate", 1e-4),
                             max_steps=op.get("max_steps", 100000), eta_min=op.get("eta_min", 0.0), store=shared)
            if len(tapes) >= 4:   # bound the memory held by tapes of odd shapes
                tapes.pop(next(k for k in tapes if k != key))
            tapes[key] = ts
        self.__dict__["_tq_train_last"] = ts
        return ts

    def step(self, batch, batch_idx=0):
        """Loss of one batch (reference: edm.py:115-134); the gradients are left in the step's flat buffer."""
        with device_guard(self.device):
            return self._train_step(batch).forward_backward(batch["signal"], batch.get("cond"))

    def training_step(self, batch, batch_idx=0):
        """Loss, gradients, gradient all-reduce over the initialised process group, Adam (cosine schedule) and EMA
        update in one call (reference: training_step + configure_optimizers + the EMA callback, edm.py:136-139,240-251,
        ema.py:24-28).  `sync_trained_weights()` writes the master (or EMA) parameters back into the module."""
        import torch.distributed as dist

        with device_guard(self.device):
            ts = self._train_step(batch)
            ws = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            loss = ts.forward_backward(batch["signal"], batch.get("cond"), world_size=ws)
            ts.optimizer_step(ws)
        return loss

    def sync_trained_weights(self, ema: bool = False):
        ts = self.__dict__.get("_tq_train_last")
        if ts is not None:
            ts.sync_module(ema=ema)

    def validation_step(self, batch, batch_idx=0):  # pragma: no cover
        raise NotImplementedError("tqdne_b200: validation-time sampling is `evaluate(batch)`; there is no validation loss path")

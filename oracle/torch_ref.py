"""CPU restatement of the reference networks and sampler -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Functional torch code driven by a plain state dict + the architecture dict, written from the reference's
behaviour (file:line cited per function), not from its module classes.  fp32 network arithmetic, fp64 sampler
state, exactly like the reference on CPU.  Pinned against the real reference modules by tests/golden/.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _conv(x, w, b, stride=1):
    """conv_nd(..., padding='same') / Downsample's stride-2 k=3 p=1 conv (tqdne/nn.py:16-24, blocks.py:93-101)."""
    k = w.shape[-1]
    fn = F.conv1d if w.dim() == 3 else F.conv2d
    return fn(x, w, b, stride=stride, padding=k // 2)


def _gn(x, w, b):
    """GroupNorm32(32, C) in fp32 (tqdne/nn.py:11-13,90-105)."""
    return F.group_norm(x.float(), 32, w, b, eps=1e-5).to(x.dtype)


def _res_block(sd, p, x, emb):
    """ResBlock._forward (tqdne/unet.py:131-143) and the embedding-free blocks.ResBlock (blocks.py:256-260)."""
    h = _conv(F.silu(_gn(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"])),
              sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"])
    film = None
    if emb is not None:
        e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
        e = e[(...,) + (None,) * (h.dim() - 2)]
        if e.shape[1] == 2 * h.shape[1]:   # use_scale_shift_norm: the projection is 2 * out_channels wide (unet.py:95)
            film = torch.chunk(e, 2, dim=1)
        else:
            h = h + e
    hn = _gn(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"])
    if film is not None:
        hn = hn * (1 + film[0]) + film[1]   # unet.py:135-139
    h = _conv(F.silu(hn), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"])
    if p + "skip_connection.weight" in sd:
        x = _conv(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


def _attention(sd, p, x, heads, causal=False):
    """AttentionBlock._forward + QKVAttention.forward (tqdne/blocks.py:139-145,156-190); `causal` = use_causal_mask
    (:181-186)."""
    b, c = x.shape[:2]
    spatial = x.shape[2:]
    qkv = _conv(_gn(x, sd[p + "norm.weight"], sd[p + "norm.bias"]), sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(b, 3 * c, -1)
    t = qkv.shape[-1]
    d = c // heads
    q, k, v = qkv.chunk(3, dim=1)
    s = 1 / math.sqrt(math.sqrt(d))
    q = (q * s).reshape(b * heads, d, t)
    k = (k * s).reshape(b * heads, d, t)
    w = torch.einsum("bct,bcs->bts", q, k)
    if causal:
        w = w.masked_fill(torch.tril(torch.ones(t, t)).unsqueeze(0) == 0, -torch.inf)
    w = torch.softmax(w.float(), dim=-1).to(q.dtype)
    a = torch.einsum("bts,bcs->bct", w, v.reshape(b * heads, d, t)).reshape(b, c, *spatial)
    return x + _conv(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def _upsample(sd, p, x):
    """Upsample.forward: nearest x2 then conv -- no conv when the layer was built with use_conv=False, i.e. has no
    parameters (tqdne/blocks.py:59-65)."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    if p + "conv.weight" not in sd:
        return x
    return _conv(x, sd[p + "conv.weight"], sd[p + "conv.bias"])


def _downsample(sd, p, x):
    """Downsample.forward: stride-2 conv, or the parameter-free average pool of use_conv=False (tqdne/blocks.py:93-108)."""
    if p + "op.weight" not in sd:
        return (F.avg_pool1d if x.dim() == 3 else F.avg_pool2d)(x, 2, 2)
    return _conv(x, sd[p + "op.weight"], sd[p + "op.bias"], stride=2)


def unet_forward(sd: dict, cfg: dict, x, timesteps, cond=None, prefix: str = ""):
    """UNetModel.forward (tqdne/unet.py:360-398) with the topology of UNetModel.__init__ (unet.py:188-358)."""
    heads = cfg.get("num_heads", 1)
    causal = bool(cfg.get("use_causal_mask", False)) and not cfg.get("flash_attention", True)   # blocks.py:136-139
    mult = cfg.get("channel_mult", (1, 2, 4, 8))
    nres = cfg["num_res_blocks"]
    att = cfg.get("attention_resolutions", (8, 16, 32))
    g = lambda k: sd[prefix + k]  # noqa: E731
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd

    # GaussianFourierProjection (blocks.py:22-26) -> time_mlp; cond_mlp added in place (unet.py:383-388)
    h = timesteps[:, None] * g("time_embed.W")[None, :] * 2 * torch.pi
    emb = torch.cat([torch.sin(h), torch.cos(h)], dim=-1)
    emb = F.linear(F.silu(F.linear(emb, g("time_mlp.0.weight"), g("time_mlp.0.bias"))), g("time_mlp.2.weight"),
                   g("time_mlp.2.bias"))
    if cfg.get("cond_features") is not None:
        if cfg.get("cond_emb_scale") is not None:   # unet.py:386-387 (runs for one conditioning feature only, blocks.py:23)
            hc = cond[:, None] * g("cond_embed.W")[None, :] * 2 * torch.pi
            cond = torch.cat([torch.sin(hc), torch.cos(hc)], dim=-1).view(cond.shape[0], -1)
        emb = emb + F.linear(F.silu(F.linear(cond, g("cond_mlp.0.weight"), g("cond_mlp.0.bias"))),
                             g("cond_mlp.2.weight"), g("cond_mlp.2.bias"))

    hs = []
    h = _conv(x, g("input_blocks.0.0.weight"), g("input_blocks.0.0.bias"))
    hs.append(h)
    idx, ds = 1, 1
    for level in range(len(mult)):
        for _ in range(nres):
            h = _res_block(sub, f"input_blocks.{idx}.0.", h, emb)
            if ds in att:
                h = _attention(sub, f"input_blocks.{idx}.1.", h, heads, causal)
            hs.append(h)
            idx += 1
        if level != len(mult) - 1:
            h = _downsample(sub, f"input_blocks.{idx}.0.", h)
            hs.append(h)
            idx += 1
            ds *= 2
    h = _res_block(sub, "middle_block.0.", h, emb)
    h = _attention(sub, "middle_block.1.", h, heads, causal)
    h = _res_block(sub, "middle_block.2.", h, emb)
    idx = 0
    for level in reversed(range(len(mult))):
        for i in range(nres + 1):
            h = torch.cat([h, hs.pop()], dim=1)  # unet.py:396
            h = _res_block(sub, f"output_blocks.{idx}.0.", h, emb)
            j = 1
            if ds in att:
                h = _attention(sub, f"output_blocks.{idx}.{j}.", h, heads, causal)
                j += 1
            if level and i == nres:
                h = _upsample(sub, f"output_blocks.{idx}.{j}.", h)
                ds //= 2
            idx += 1
    return _conv(F.silu(_gn(h, g("out.0.weight"), g("out.0.bias"))), g("out.2.weight"), g("out.2.bias"))


def _coder_blocks(sd, cfg, x, prefix, kind):
    mult = cfg.get("channel_mult", (1, 2, 4, 8))
    nres = cfg["num_res_blocks"]
    att = cfg.get("attention_resolutions", (8, 16, 32))
    heads = cfg.get("num_heads", 1)
    name = "down_blocks" if kind == "encoder" else "up_blocks"
    idx = 0
    last = len(mult) - 1
    ds = 1 if kind == "encoder" else 2**last
    levels = range(len(mult)) if kind == "encoder" else reversed(range(len(mult)))
    for level in levels:
        if kind == "decoder" and level != last:
            x = _upsample(sd, f"{prefix}{name}.{idx}.", x)
            idx += 1
            ds //= 2
        for _ in range(nres):
            x = _res_block(sd, f"{prefix}{name}.{idx}.", x, None)
            idx += 1
            if ds in att:
                x = _attention(sd, f"{prefix}{name}.{idx}.", x, heads)
                idx += 1
        if kind == "encoder" and level != last:
            x = _conv(x, sd[f"{prefix}{name}.{idx}.op.weight"], sd[f"{prefix}{name}.{idx}.op.bias"], stride=2)
            idx += 1
            ds *= 2
    return x


def decoder_forward(sd: dict, cfg: dict, z, prefix: str = ""):
    """Decoder.forward (tqdne/blocks.py:384-436)."""
    x = _conv(z, sd[prefix + "input_layer.weight"], sd[prefix + "input_layer.bias"])
    x = _coder_blocks(sd, cfg, x, prefix, "decoder")
    return _conv(x, sd[prefix + "output_layer.weight"], sd[prefix + "output_layer.bias"])


def encoder_forward(sd: dict, cfg: dict, x, prefix: str = ""):
    """Encoder.forward (tqdne/blocks.py:313-348)."""
    x = _conv(x, sd[prefix + "input_layer.weight"], sd[prefix + "input_layer.bias"])
    x = _coder_blocks(sd, cfg, x, prefix, "encoder")
    return _conv(x, sd[prefix + "output_layer.weight"], sd[prefix + "output_layer.bias"])


def classifier_embed(sd: dict, enc_cfg: dict, x):
    """LithningClassifier.embed (tqdne/classifier.py:51-55): Encoder -> mean over the spatial dims -> SiLU, Linear,
    SiLU, Linear (output_MLP, classifier.py:39-44)."""
    h = encoder_forward(sd, enc_cfg, x, prefix="encoder.")
    h = h.mean(dim=list(range(2, h.dim())))
    h = F.linear(F.silu(h), sd["output_MLP.1.weight"], sd["output_MLP.1.bias"])
    return F.linear(F.silu(h), sd["output_MLP.3.weight"], sd["output_MLP.3.bias"])


def classifier_forward(sd: dict, enc_cfg: dict, x):
    """LithningClassifier.forward (tqdne/classifier.py:57-59)."""
    return F.linear(classifier_embed(sd, enc_cfg, x), sd["output_layer.weight"], sd["output_layer.bias"])


def frechet_distance(x, y, eps: float = 1e-6):
    """tqdne.metric.frechet_distance (tqdne/metric.py:13-44), non-isotropic branch, NumPy / SciPy like the reference."""
    import numpy as np
    from scipy import linalg

    mu_x, mu_y = x.mean(0), y.mean(0)
    cov_x, cov_y = np.cov(x, rowvar=False), np.cov(y, rowvar=False)
    covmean = linalg.sqrtm(cov_x @ cov_y)
    if not np.isfinite(covmean).all():
        off = np.eye(cov_x.shape[0]) * eps
        covmean = linalg.sqrtm((cov_x + off) @ (cov_y + off))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return float(np.sum((mu_x - mu_y) ** 2) + np.trace(cov_x) + np.trace(cov_y) - 2 * np.trace(covmean))


# ---- EDM (tqdne/edm.py) ----------------------------------------------------------------------------
SIGMA_MIN, SIGMA_MAX, RHO, SIGMA_DATA = 0.002, 80.0, 7.0, 0.5
S_CHURN, S_MIN, S_MAX, S_NOISE = 40, 0.05, 50, 1.003


def sampling_sigmas(num_steps: int):
    """EDM.sampling_sigmas (edm.py:39-46): fp32, trailing 0."""
    i = torch.arange(num_steps, dtype=torch.float32)
    s = (SIGMA_MAX ** (1 / RHO) + i / (num_steps - 1) * (SIGMA_MIN ** (1 / RHO) - SIGMA_MAX ** (1 / RHO))) ** RHO
    return torch.cat([s, torch.zeros_like(s[:1])])


def denoise(sd, cfg, x, sigma, cond, prefix="unet.", cond_sample=None):
    """LightningEDM.forward (edm.py:105-113): D = c_out * F(c_in x [cat cond_sample], 0.25 ln sigma, cond) + c_skip x."""
    ex = (...,) + (None,) * (x.dim() - 1)
    c_in = 1 / (sigma**2 + SIGMA_DATA**2) ** 0.5
    c_out = sigma * SIGMA_DATA / (sigma**2 + SIGMA_DATA**2) ** 0.5
    c_skip = SIGMA_DATA**2 / (sigma**2 + SIGMA_DATA**2)
    inp = x * c_in[ex]
    if cond_sample is not None:
        inp = torch.cat((inp, cond_sample), dim=1)   # edm.py:109
    out = unet_forward(sd, cfg, inp, 0.25 * sigma.log(), cond, prefix)
    return out * c_out[ex] + c_skip[ex] * x


def heun_sample(sd, cfg, eps, sigmas, cond, prefix="unet.", trace=None, cond_sample=None):
    """LightningEDM.sample_deterministically (edm.py:171-196): fp64 state, fp32 denoiser, NFE = 2N-1."""
    n = len(sigmas) - 1
    x_next = eps
    for i in range(n):
        s, s_next = sigmas[i], sigmas[i + 1]
        x = x_next
        pred = denoise(sd, cfg, x.float(), s.repeat(len(x)), cond, prefix, cond_sample).double()
        if trace is not None:
            trace.append(pred)
        d = (x - pred) / s
        x_next = x + d * (s_next - s)
        if i < n - 1:
            pred2 = denoise(sd, cfg, x_next.float(), s_next.repeat(len(x)), cond, prefix, cond_sample).double()
            if trace is not None:
                trace.append(pred2)
            d2 = (x_next - pred2) / s_next
            x_next = x + (s_next - s) * (0.5 * d + 0.5 * d2)
    return x_next


def sigma_hat(sigma, num_steps: int):
    """EDM.sigma_hat (edm.py:48-52): fp32 0-dim arithmetic."""
    gamma = min(S_CHURN / num_steps, 2**0.5 - 1) if S_MIN <= sigma <= S_MAX else 0
    return sigma + gamma * sigma


def heun_sample_stochastic(sd, cfg, eps, sigmas, cond, noises, prefix="unet."):
    """LightningEDM.sample_stochastically (edm.py:198-230) with the th.randn_like draws given as `noises[i]`."""
    n = len(sigmas) - 1
    x_next = eps
    for i in range(n):
        s, s_next = sigmas[i], sigmas[i + 1]
        x = x_next
        s_hat = sigma_hat(s, n)
        x_hat = x + (noises[i] * S_NOISE) * (s_hat**2 - s**2) ** 0.5
        pred = denoise(sd, cfg, x_hat.float(), s_hat.repeat(len(x)), cond, prefix).double()
        d = (x_hat - pred) / s_hat
        x_next = x_hat + d * (s_next - s_hat)
        if i < n - 1:
            pred2 = denoise(sd, cfg, x_next.float(), s_next.repeat(len(x)), cond, prefix).double()
            d2 = (x_next - pred2) / s_next
            x_next = x_hat + (s_next - s_hat) * (0.5 * d + 0.5 * d2)
    return x_next


# ---- representations (tqdne/representation.py) ---------------------------------------------------------
def mavg_envelope_inverse(rep, log_eps=1e-6, eps=1e-6):
    """MovingAverageEnvelope.invert_representation (representation.py:57-60), NumPy semantics."""
    import numpy as np

    rep = np.asarray(rep)
    scaled, log_env = np.split(rep, 2, axis=-2)
    return scaled * (np.exp(log_env + np.log(log_eps) / 2) + eps)

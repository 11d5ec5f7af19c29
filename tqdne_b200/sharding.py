"""Multi-GPU sampling: the batch shards as independent samples, one process per GPU.

No collective runs on the data path; the only exchange is the final gather of the finished waveforms
(`gather_waveforms`), 49 KB per waveform over NVLink / NVSwitch via NCCL (gloo in the CPU tests).
Noise is a function of the GLOBAL sample index, so results do not depend on the number of GPUs.
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun environment."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_bounds(n_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of the global batch owned by `rank` (ragged tails allowed)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_noise(shape: tuple, lo: int, hi: int, seed: int, device, dtype=torch.float64, block: int = 64) -> torch.Tensor:
    """Unit normal noise for global samples [lo, hi): sample i is drawn from a generator seeded by
    (seed, i // block) so any partition of the batch reproduces the same per-sample noise."""
    out = torch.empty((hi - lo, *shape), dtype=dtype, device=device)
    b0, b1 = lo // block, (hi - 1) // block if hi > lo else lo // block - 1
    for b in range(b0, b1 + 1):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 1_000_003 + b)
        blk = torch.randn((block, *shape), generator=g, dtype=dtype)
        s, e = max(lo, b * block), min(hi, (b + 1) * block)
        out[s - lo:e - lo] = blk[s - b * block:e - b * block].to(device)
    return out


def gather_waveforms(local: torch.Tensor, n_total: int, dst: int = 0) -> torch.Tensor | None:
    """Final gather of per-rank results [n_local, ...] to rank `dst` in global order.  Returns the
    [n_total, ...] tensor on `dst`, None elsewhere.  Single-process: returns `local`."""
    rank, ws = world()
    if ws == 1 or not (dist.is_available() and dist.is_initialized()):
        return local
    sizes = [shard_bounds(n_total, r, ws) for r in range(ws)]
    max_n = max(hi - lo for lo, hi in sizes)
    even = all(hi - lo == max_n for lo, hi in sizes)
    if even:
        pad = local.contiguous()
    else:
        pad = torch.zeros((max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    # the receive buffers are the rows of ONE [ws, max_n, ...] tensor: equal shards need no concatenation afterwards
    stacked = torch.empty((ws, max_n, *local.shape[1:]), dtype=local.dtype, device=local.device) if rank == dst else None
    dist.gather(pad, list(stacked.unbind(0)) if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    if even:
        return stacked.view(ws * max_n, *local.shape[1:])
    return torch.cat([stacked[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def gather_ragged(local: torch.Tensor, counts: list[int], dst: int = 0) -> list[torch.Tensor] | None:
    """Gather per-rank tensors [counts[r], ...] (counts may differ and may be 0) to rank `dst`; returns the list of
    per-rank tensors on `dst`, None elsewhere.  Single-process: [local]."""
    rank, ws = world()
    if ws == 1 or not (dist.is_available() and dist.is_initialized()):
        return [local]
    max_n = max(counts)
    if max_n == 0:
        return [local[:0] for _ in counts] if rank == dst else None
    pad = local
    if local.shape[0] != max_n:
        pad = torch.zeros((max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    stacked = torch.empty((ws, max_n, *local.shape[1:]), dtype=local.dtype, device=local.device) if rank == dst else None
    dist.gather(pad.contiguous(), list(stacked.unbind(0)) if rank == dst else None, dst=dst)
    if rank != dst:
        return None
    return [stacked[r, : counts[r]] for r in range(ws)]


_PINNED: dict = {}


def to_host(t: torch.Tensor) -> torch.Tensor:
    """Device -> host copy into a cached PINNED buffer (a fresh pageable `t.cpu()` of the gathered waveforms costs
    first-touch page faults on ~50 KB per waveform every call).  The returned tensor is reused by the next call with the
    same shape: consume (or copy) it before calling again."""
    if t.device.type != "cuda":
        return t
    key = (tuple(t.shape), t.dtype)
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) > 8:
            _PINNED.clear()
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        _PINNED[key] = buf
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf

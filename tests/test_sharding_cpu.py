"""CPU: the multi-GPU host logic (batch sharding, partition-independent noise, final gather) with
world_size = 2 over gloo -- the N > 1 path of bench.py / generate_waveforms without GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tqdne_b200 import sharding


def test_shard_bounds_cover_the_batch_without_overlap():
    for n in (0, 1, 5, 8, 8192, 8191):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_noise_is_a_function_of_the_global_sample_index():
    full = sharding.global_noise((2, 3), 0, 200, seed=7, device="cpu")
    for ws in (2, 3, 8):
        parts = [sharding.global_noise((2, 3), *sharding.shard_bounds(200, r, ws), seed=7, device="cpu") for r in range(ws)]
        assert torch.equal(torch.cat(parts), full)
    assert not torch.equal(full, sharding.global_noise((2, 3), 0, 200, seed=8, device="cpu"))
    assert full.dtype == torch.float64 and abs(float(full.std()) - 1) < 0.1


def _worker(rank, ws, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = sharding.shard_bounds(n_total, rank, ws)
        # each "waveform" carries its global sample index so the gathered order can be checked
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None].expand(hi - lo, 3, 16).contiguous()
        full = sharding.gather_waveforms(local, n_total, dst=0)
        # the round-by-round gather of generate(): batches of 2 rows per rank, ragged last round, an empty contribution
        batch, got = 2, torch.full((n_total,), -1.0)
        shards = [sharding.shard_bounds(n_total, r, ws) for r in range(ws)]
        rounds = max(-(-(h - l) // batch) for l, h in shards)
        for k in range(rounds):
            spans = [(min(h, l + k * batch), min(h, l + (k + 1) * batch)) for l, h in shards]
            a, b = spans[rank]
            parts = sharding.gather_ragged(local[a - lo:b - lo], [y - x for x, y in spans])
            if rank == 0:
                for (x, y), part in zip(spans, parts):
                    got[x:y] = part[:, 0, 0]
            else:
                assert parts is None
        if rank == 0:
            ok = full.shape == (n_total, 3, 16) and torch.equal(full[:, 0, 0], torch.arange(n_total, dtype=torch.float32))
            ok = ok and torch.equal(got, torch.arange(n_total, dtype=torch.float32))
            ret.put(bool(ok))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_final_gather_world_size_2_gloo_ragged_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True
